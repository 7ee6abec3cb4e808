"""does running the main chain on a high-priority stream (side-stream weight gradients at default priority) help?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
torch.manual_seed(0)
m = SimpleTrainModule(num_cls=4).cuda().train()
(opt,), _ = m.configure_optimizers()
x, y = bench.synthetic_batch(64, 120, 160, 4, seed=1)
x, y = x.cuda(), y.cuda()
def step(i):
    opt.zero_grad(); l = m.training_step((x, y), i); l.backward(); opt.step()
def run(stream):
    with torch.cuda.stream(stream):
        for i in range(4): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(15): step(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 15
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")
print("default stream   ms/step", round(run(torch.cuda.current_stream()), 3))
print("prio 0 stream    ms/step", round(run(torch.cuda.Stream(priority=0)), 3))
print("prio -1 stream   ms/step", round(run(torch.cuda.Stream(priority=-1)), 3))
print("default again    ms/step", round(run(torch.cuda.current_stream()), 3))
