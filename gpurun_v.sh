timeout 600 python -m pytest tests -x -q -m gpu --timeout 60 --timeout-method=thread 2>&1 | tail -4
timeout 100 python __graft_entry__.py smoke
timeout 300 python bench.py --steps 2 --warmup 3 --no-abc --cpu-seconds 5 | cut -c1-400
