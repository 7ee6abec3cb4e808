timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -q -x -s -k mirror 2>&1 | grep -E "mirror off|passed|failed|^E" | head -8
