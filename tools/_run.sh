timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -s 2>&1 | grep -E "^E|FAILED|passed|failed|worst" | head -20
