import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sim2real_lane_segment_b200 import _lib
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
from oracle import steps as osteps
torch.manual_seed(0)
m = SimpleTrainModule(num_cls=4).cuda().train()
(opt,), _ = m.configure_optimizers()
x, y = osteps.synthetic_batch(64, 120, 160, 4, seed=1); x, y = x.cuda(), y.cuda()
for i in range(3):
    opt.zero_grad(); l = m.training_step((x, y), i); l.backward(); opt.step()
torch.cuda.synchronize()
L = _lib.lib()
buf = (ctypes.c_longlong * 16)()
L.rl_debug_ts(buf)
t = list(buf)
print("last dgrad_dense launch (level-0 down layer 0), block 0 thread 0, cycles since kernel start:")
print([t[i] - t[0] for i in range(9)])
