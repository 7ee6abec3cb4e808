timeout 300 python -m pytest tests/test_forward_gpu.py tests/test_shapes_gpu.py -m gpu -q -x 2>&1 | tail -3
for v in 0; do echo "dbg $v"; RL_FWD_DBG=$v timeout 120 python tools/per_launch.py dense_fwd 2>&1 | tail -1; done
