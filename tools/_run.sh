timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_modules_gpu.py -m gpu -q 2>&1 | grep -E "^E|FAILED|passed|failed" | head -12
