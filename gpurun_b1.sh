python __graft_entry__.py smoke
for ws in C1 C2; do
python bench.py --workload $ws --steps 2 --warmup 1 --no-abc --cpu-seconds 5
done
python bench.py --workload C2 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width 16
python bench.py --workload C2 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --tile-width 8
python bench.py --workload C2 --steps 1 --warmup 1 --no-abc --no-cpu-baseline --no-e2e --replicates 40000 --tile-width 8
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
