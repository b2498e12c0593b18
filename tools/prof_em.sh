# usage (under gpurun): ENGINE=5 KREGEX=k_em_seg TAG=r02_seg bash tools/prof_em.sh
# launch list of the library's kernels, then one `ncu --set full` capture of the E+M kernel, each after
# the identical command has exited 0 without ncu (B200_PROFILING.md); summaries: tools/ncu_summary.py
ENGINE=${ENGINE:-0}; KREGEX=${KREGEX:-k_em_}; TAG=${TAG:-r02}
CMD="python bench.py --steps 3 --warmup 3 --rows 2000000 --e2e-steps 1 --cpu-rows 2000 --engine $ENGINE --no-records --no-class-prob"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_\|mxd -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/prof_em_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
