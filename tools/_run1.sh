set -x
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cluster_shapes or feature_pair or chunked or estep or golden or drift" > gpurun_out/r02t_pytest.log 2>&1; tail -5 gpurun_out/r02t_pytest.log
python tools/sweep_plans.py --rows 5000000 --features 100 --k 64 --engine 5 x,x,x > gpurun_out/r02t_sweep.txt 2>&1
python tools/sweep_plans.py --rows 5000000 --features 100 --k 64 --engine 4 x,x,x >> gpurun_out/r02t_sweep.txt 2>&1
python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 5 x,x,x >> gpurun_out/r02t_sweep.txt 2>&1
MIXDIR_B200_SEG_GLOBAL_STATS=1 python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 5 x,x,x >> gpurun_out/r02t_sweep.txt 2>&1
cat gpurun_out/r02t_sweep.txt
