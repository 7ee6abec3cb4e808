"""one tiny training step + eval forward, for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
torch.manual_seed(0)
B, H, W = 2, 64, 96
m = SimpleTrainModule(num_cls=4).cuda().train()
(opt,), _ = m.configure_optimizers()
x = torch.randn(B, 3, H, W, device="cuda"); y = torch.randint(0, 4, (B, H, W), device="cuda")
opt.zero_grad(); loss = m.training_step((x, y), 0); loss.backward(); opt.step()
torch.cuda.synchronize()
m.eval()
with torch.no_grad():
    p = m(x)
torch.cuda.synchronize()
print("sanitize step ok", float(loss), float(p.sum()))
