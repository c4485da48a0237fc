export TILE=4 RUNS=40000 BINS=256
python scripts/prof_case.py > gpurun_out/p2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -s 2 -c 1 -o gpurun_out/prof_r1_l4 python scripts/prof_case.py > gpurun_out/p2_ncu.log 2>&1
cat gpurun_out/p2_plain.log; tail -3 gpurun_out/p2_ncu.log
