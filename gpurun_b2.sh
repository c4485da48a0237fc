python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload C1 --steps 2 --warmup 1 --no-abc --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C1', d['value'], d['ms_per_step'], d['config']['tile_width'], d['config']['blocks_per_sm'])"
for tw in 32 16 8; do
python bench.py --workload C2 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width $tw | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['value'], d['ms_per_step'], d['config']['tile_width'], d['config']['blocks_per_sm'])"
done
