"""per-launch CUDA-event times of one training step (B=64, 120x160) — cheaper than an ncu launch list"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sim2real_lane_segment_b200 import _lib
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
from oracle import steps as osteps
torch.manual_seed(0)
m = SimpleTrainModule(num_cls=4).cuda().train()
(opt,), _ = m.configure_optimizers()
x, y = osteps.synthetic_batch(64, 120, 160, 4, seed=1)
x, y = x.cuda(), y.cuda()
def step(i):
    opt.zero_grad(); l = m.training_step((x, y), i); l.backward(); opt.step()
for i in range(3): step(i)
L = _lib.lib()
for eng in m.runtime().engines.values():
    eng.set_option("overlap", 0)
    if os.environ.get("BWD_FUSED") == "0":
        eng.set_option("bwd_fused", 0)
step(0)
L.rl_profile_begin(); step(0)
msb = (ctypes.c_double * bench.NK)(); lnb = (ctypes.c_long * bench.NK)()
L.rl_profile_end(msb, lnb)
ms = (ctypes.c_float * 4096)(); cl = (ctypes.c_int * 4096)()
n = L.rl_profile_launches(ms, cl, 4096)
names = bench.KCLASS
want = sys.argv[1:] or names
for k in want:
    ci = names.index(k)
    vals = [round(ms[i] * 1e3) for i in range(n) if cl[i] == ci]
    print(k, "n=%d total=%.3f ms" % (len(vals), sum(vals) / 1e3), vals)
