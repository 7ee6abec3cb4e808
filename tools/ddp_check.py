"""torchrun --nproc-per-node=2 tools/ddp_check.py — the overlapped all-reduce (ddp.GradSync fired from inside the backward
pass) must produce exactly what a plain sum of the ranks' local gradients gives."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from oracle import fcdensenet as net
from oracle import steps
from sim2real_lane_segment_b200 import ddp
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, H, W, C = 4, 64, 96, 4
m = SimpleTrainModule(num_cls=C)
m.load_state_dict(net.init_state_dict(num_cls=C, seed=7))
m = m.cuda().train()
m.runtime().dropout_p = 0.0
(opt,), _ = m.configure_optimizers()
rt = m.runtime()
x, y = steps.synthetic_batch(B, H, W, C, seed=50 + rank)
x, y = x.cuda(), y.cuda()
opt.zero_grad()
m.training_step((x, y), 0).backward()
local_g = rt.grads.clone()
want = local_g.clone()
dist.all_reduce(want)                                   # reference exchange: one plain all-reduce after backward
rt.grad_sync = ddp.GradSync()
opt.zero_grad()
m.training_step((x, y), 1).backward()                   # buckets are all-reduced while backward is still running
scale = rt.grad_sync.finish()
torch.cuda.synchronize()
err = float((rt.grads - want).norm() / want.norm())
loc = float((rt.grads - local_g).norm() / local_g.norm())
print(f"rank {rank}: overlapped vs plain all-reduce rel diff {err:.2e} (vs local gradient {loc:.2e}), scale {scale}")
assert scale == 1.0 / world and err < 1e-4 and loc > 1e-2
dist.barrier()
if rank == 0:
    print("ddp_check ok")
dist.destroy_process_group()
