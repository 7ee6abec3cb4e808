"""python tools/block_probe.py B H W level up — runs the fused DenseBlock backward parity case of tests/test_ops_gpu.py for
arbitrary shapes and prints every error instead of asserting"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_ops_gpu as T
B, H, W, level, up = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] == "1"
net = T.Net(B, H, W, seed=300 + level + (7 if up else 0), options={"bwd_fused": 1})
try:
    T._block_case(net, level, up, True)
    print("PASS")
except AssertionError as e:
    print("FAIL", str(e)[:200])
