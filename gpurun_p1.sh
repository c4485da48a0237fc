python scripts/prof_case.py > gpurun_out/p1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -s 1 -c 1 -o gpurun_out/prof_r1_l32 python scripts/prof_case.py > gpurun_out/p1_ncu.log 2>&1
cat gpurun_out/p1_plain.log; tail -5 gpurun_out/p1_ncu.log
