timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/per_launch.py dense_fwd 2>&1 | tail -1
