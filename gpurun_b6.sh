python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
TILE=4 REPS=2 RUNS=40000 BINS=256 python scripts/prof_case.py
TILE=4 REPS=2 RUNS=10000 BINS=256 python scripts/prof_case.py
TILE=8 REPS=2 RUNS=10000 BINS=256 python scripts/prof_case.py
TILE=32 REPS=2 RUNS=1000 BINS=512 python scripts/prof_case.py
python bench.py --workload C2 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width 4 --smem-bins 256 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 256bins', d['value'], d['ms_per_step'], d['config']['tile_width'], d['config']['blocks_per_sm'], d['config']['kmax'], d['config']['spilled'])"
