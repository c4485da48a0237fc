python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
export TILE=4 RUNS=40000 BINS=256
python scripts/prof_case.py > gpurun_out/p3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -s 2 -c 1 -o gpurun_out/prof_r1_l4b python scripts/prof_case.py > gpurun_out/p3_ncu.log 2>&1
cat gpurun_out/p3_plain.log; tail -2 gpurun_out/p3_ncu.log
