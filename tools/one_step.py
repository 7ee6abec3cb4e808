"""two training steps (B=64, 120x160) for ncu captures: the first warms up / allocates, the second is the one profiled
(NCU_SKIP = launches of the wanted kernel in the first step)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
from oracle import steps as osteps
torch.manual_seed(0)
m = SimpleTrainModule(num_cls=4).cuda().train()
(opt,), _ = m.configure_optimizers()
x, y = osteps.synthetic_batch(64, 120, 160, 4, seed=1)
x, y = x.cuda(), y.cuda()
for i in range(int(os.environ.get("STEPS", "2"))):
    opt.zero_grad(); l = m.training_step((x, y), i); l.backward(); opt.step()
torch.cuda.synchronize()
print("ok", float(l))
