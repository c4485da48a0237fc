python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for v in mb1 mb5 mb6; do
cp variants/$v.so ecdna-evo_b200/libecdna_b200.so
echo "== $v"
TILE=4 REPS=2 RUNS=40000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=160000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=10000 BINS=256 python scripts/prof_case.py
TILE=32 REPS=2 RUNS=1000 python scripts/prof_case.py
done
