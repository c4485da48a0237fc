# launch list of the bench command (short form of the same command line), after a plain run exited 0
python bench.py --steps 2 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --replicates 2000 > gpurun_out/launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --replicates 2000 > gpurun_out/launch_ncu.log 2>&1
tail -3 gpurun_out/launch_plain.log | cut -c1-300
export TILE=4 RUNS=40000 BINS=256
python scripts/prof_case.py > gpurun_out/p4_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -s 2 -c 1 -o gpurun_out/prof_r1_l4d python scripts/prof_case.py > gpurun_out/p4_ncu.log 2>&1
cat gpurun_out/p4_plain.log
