python bench.py > gpurun_out/r02z2_bench.json 2> gpurun_out/r02z2_bench.err
echo bench rc=$?
ENGINE=0 KREGEX=k_em_seg TAG=r02z2 bash tools/prof_em.sh
