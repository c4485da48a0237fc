python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
for tw in 32 16 8 4; do
TILE=$tw REPS=2 python scripts/prof_case.py
done
TILE=8 REPS=2 RUNS=40000 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=40000 python scripts/prof_case.py
