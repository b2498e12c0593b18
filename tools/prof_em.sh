# usage (under gpurun): ENGINE=3 KREGEX=k_em_bucket TAG=r1_bucket bash tools/prof_em.sh
ENGINE=${ENGINE:-0}; KREGEX=${KREGEX:-k_em_}; TAG=${TAG:-r1}
CMD="python bench.py --steps 3 --warmup 3 --rows 2000000 --e2e-steps 1 --cpu-rows 2000 --engine $ENGINE"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/prof_em_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
