"""single-GPU reproduction helper: one training step at a given shape, reports the first NaN/Inf in activations and
gradients (python tools/nan_probe.py B H W)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fcdensenet as net
from oracle import steps
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 64, 96)
C = 4
m = SimpleTrainModule(num_cls=C)
m.load_state_dict(net.init_state_dict(num_cls=C, seed=7))
m = m.cuda().train()
rt = m.runtime()
rt.dropout_p = 0.0
(opt,), _ = m.configure_optimizers()
OPT = sys.argv[4] if len(sys.argv) > 4 else ""
for rep in range(2):
    x, y = steps.synthetic_batch(B, H, W, C, seed=50 + rep)
    x, y = x.cuda(), y.cuda()
    opt.zero_grad()
    if OPT and rep == 0:
        with torch.no_grad():
            m(x)                                    # creates the plan
        for eng in rt.engines.values():
            for o in OPT.split(","):
                eng.set_option(o.split("=")[0], int(o.split("=")[1]))
    loss = m.training_step((x, y), rep)
    loss.backward()
    torch.cuda.synchronize()
    eng = next(iter(rt.engines.values()))
    print("rep", rep, "loss", float(loss))
    for lvl in range(6):
        a = eng.level_view(lvl).float()
        g = eng.level_view(lvl, grad=True).float()
        badc = torch.nonzero(~torch.isfinite(g).all(dim=(0, 1, 2))).flatten().tolist()
        rng = []
        for c in badc:
            if rng and rng[-1][1] == c - 1: rng[-1][1] = c
            else: rng.append([c, c])
        nb = int((~torch.isfinite(g)).sum())
        acc = eng.level_acc(lvl)
        print(f"  level {lvl} {tuple(g.shape)}: act finite {bool(torch.isfinite(a).all())} dact bad elements {nb} channel ranges {rng[:8]} acc finite {bool(torch.isfinite(acc).all())}")
        if nb:
            bp = torch.nonzero(~torch.isfinite(g[..., badc[0]]))
            print("     first bad channel", badc[0], "bad pixels", bp.shape[0], bp[:6].tolist())
    bad = [n for n, p in m.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print("  non-finite grads:", len(bad), bad[:12])
