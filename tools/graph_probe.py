import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
from sim2real_lane_segment_b200.graphs import GraphedTrainStep
from oracle import steps as osteps
for overlap in (0, 1):
    torch.manual_seed(0)
    m = SimpleTrainModule(num_cls=4).cuda().train()
    (opt,), _ = m.configure_optimizers()
    x, y = osteps.synthetic_batch(4, 64, 96, 4, seed=1)
    x, y = x.cuda(), y.cuda()
    opt.zero_grad(); m.training_step((x, y), 0).backward(); opt.step()
    for eng in m.runtime().engines.values():
        eng.set_option("overlap", overlap)
    try:
        g = GraphedTrainStep(m, opt, x, y, warmup=2)
        l = [float(g(x, y)) for _ in range(3)]
        print("overlap", overlap, "graph ok", l)
    except Exception as e:
        print("overlap", overlap, "FAILED", type(e).__name__, str(e)[:300])
        traceback.print_exc(limit=3)
    torch.cuda.synchronize()
