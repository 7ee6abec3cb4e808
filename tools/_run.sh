timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -k "40x56" 2>&1 | grep -E "^E|assert" | head -12
