python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02z_pytest_multi8.log 2>&1; tail -3 gpurun_out/r02z_pytest_multi8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err
echo bench rc=$?
tail -c 600 gpurun_out/r02z_bench_n8.json
