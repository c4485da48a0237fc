python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
TILE=4 REPS=2 RUNS=40000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=160000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=10000 BINS=256 python scripts/prof_case.py
TILE=32 REPS=2 RUNS=1000 python scripts/prof_case.py
