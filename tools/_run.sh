timeout 600 python -m pytest tests/test_umma_selftest.py tests/test_forward_gpu.py tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -2
