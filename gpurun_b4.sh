python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
for tw in 32 8 4; do
TILE=$tw REPS=2 RUNS=40000 python scripts/prof_case.py
done
TILE=4 REPS=2 RUNS=40000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=160000 BINS=256 python scripts/prof_case.py
