"""GPU diagnostic: per-tensor gradient error of (a) the CUDA path and (b) the oracle under torch bf16 autocast,
both against the fp32 oracle — separates kernel bugs from the bf16 noise floor of this network."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fcdensenet as net, steps
from tests.test_backward_gpu import _run_cuda
from tests.util import rel_err

path = sys.argv[1]
fx = torch.load(path)
sd, eng, g, loss = _run_cuda(fx, None)
C = fx["num_cls"]
x, y = fx["x"].cuda(), fx["y"].long().cuda()

def oracle_grads(autocast):
    s = {k: v.cuda() for k, v in sd.items()}
    tr = steps.Trainer(s, num_cls=C)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        tr.step(x, y)
    return {k: v.float().cpu() for k, v in tr.grads().items()}

g32 = oracle_grads(False)
g16 = oracle_grads(True)
gmax = max(float(v.norm()) for v in g32.values())
rows = []
for k, v in g32.items():
    n = float(v.norm())
    if n > 1e-4 * gmax:
        rows.append((rel_err(g[k], v), rel_err(g16[k], v), n, k))
rows.sort(reverse=True)
print(f"{'ours':>8} {'autocast':>8} {'|g|':>10}  tensor")
for r in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{r[0]:8.4f} {r[1]:8.4f} {r[2]:10.3e}  {r[3]}")
import statistics
print("median ours", statistics.median(r[0] for r in rows), "median autocast", statistics.median(r[1] for r in rows), "n", len(rows))
