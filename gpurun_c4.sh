python scripts/abc_full.py
python bench.py --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e > gpurun_out/t_plain.log 2>&1 && ncu --set full --clock-control none -k regex:ssa_kernelILi4 -s 1 -c 1 -o gpurun_out/prof_r1_c2_full python bench.py --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e > gpurun_out/t_ncu.log 2>&1
tail -2 gpurun_out/t_ncu.log | cut -c1-200
