timeout 120 python tools/sweep_plans.py --rows 480000 --features 60 --k 32 --engine 5 --steps 4 x,x,x 2>&1 | tail -1 | cut -c1-80
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 5 x,x,x 2>&1 | tail -1 | cut -c1-80
timeout 120 python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 5 x,x,x 2>&1 | tail -1 | cut -c1-80
