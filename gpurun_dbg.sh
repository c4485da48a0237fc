timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 60 -k distribution_matches 2>&1 | grep -E "^E|assert|line" | head -20
