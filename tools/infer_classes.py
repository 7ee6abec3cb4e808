"""per-kernel-class CUDA-event times of one 32-image 480x640 inference chunk (evaluate_batch with predictions)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sim2real_lane_segment_b200 import _lib
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
torch.manual_seed(0)
m = SimpleTrainModule(num_cls=4).cuda().eval()
x = torch.randn(32, 3, 480, 640, device="cuda"); y = torch.randint(0, 4, (32, 480, 640), device="cuda")
for _ in range(2): m.evaluate_batch((x, y), return_predictions=True)
torch.cuda.synchronize()
L = _lib.lib()
L.rl_profile_begin(); m.evaluate_batch((x, y), return_predictions=True)
msb = (ctypes.c_double * 13)(); lnb = (ctypes.c_long * 13)()
L.rl_profile_end(msb, lnb)
tot = sum(msb)
for n, t, c in zip(bench.KCLASS, msb, lnb):
    if c: print(f"{n:14s} {t:8.3f} ms {100*t/tot:5.1f}%  {c} launches")
print("total", round(tot, 2), "ms for 32 images ->", round(32 / tot * 1e3), "img/s (kernel time only)")
