timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/config_times.py --no-cpu 2>&1 | grep -E "^M1|^M2|^S3"
timeout 200 python tools/sweep_plans.py --rows 5000000 --features 100 --k 64 --engine 4 x,x,x 2>&1 | tail -1 | cut -c1-100
timeout 120 python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 5 x,x,x 2>&1 | tail -1 | cut -c1-80
timeout 120 python tools/sweep_plans.py --rows 4000000 --features 60 --k 32 --engine 4 x,x,x 2>&1 | tail -1 | cut -c1-80
