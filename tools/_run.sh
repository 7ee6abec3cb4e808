timeout 300 python -m pytest tests/test_forward_gpu.py tests/test_shapes_gpu.py -m gpu -q -x 2>&1 | grep -E "^E|FAILED|passed|failed" | head -12
timeout 300 python tools/per_launch.py dense_fwd 2>&1 | tail -1
