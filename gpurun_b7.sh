for tw in 32 8 4; do
python bench.py --workload C5 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width $tw --replicates 1000 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C5', d['value'], d['ms_per_step'], d['config']['tile_width'], d['config']['blocks_per_sm'], d['config']['kmax'], d['config']['spilled'])"
done
for tw in 32 8 4; do
python bench.py --workload C1 --steps 2 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width $tw | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C1', d['value'], d['ms_per_step'], d['config']['tile_width'])"
done
for tw in 8 4; do
python bench.py --workload C3 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width $tw --smem-bins 256 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', d['value'], d['ms_per_step'], d['config']['tile_width'], d['config']['blocks_per_sm'], d['config']['kmax'], d['config']['spilled'])"
done
