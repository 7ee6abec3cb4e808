"""As grad_noise_probe.py but on a synthetic batch of arbitrary size (no fixture): python tools/grad_noise_probe2.py B H W"""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fcdensenet as net, steps
from tests.test_backward_gpu import _run_cuda
from tests.util import rel_err

B, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, y = steps.synthetic_batch(B, H, W, 4, seed=5)
fx = {"B": B, "H": H, "W": W, "num_cls": 4, "seed": 3, "x": x, "y": y.to(torch.uint8)}
sd, eng, g, loss = _run_cuda(fx, None)
xc, yc = x.cuda(), y.cuda()

def oracle_grads(autocast):
    s = {k: v.cuda() for k, v in sd.items()}
    tr = steps.Trainer(s, num_cls=4)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        l, _, _ = tr.step(xc, yc)
    return {k: v.float().cpu() for k, v in tr.grads().items()}, l

g32, l32 = oracle_grads(False)
g16, l16 = oracle_grads(True)
print("loss ours", loss, "fp32", l32, "autocast", l16)
gmax = max(float(v.norm()) for v in g32.values())
rows = []
for k, v in g32.items():
    n = float(v.norm())
    if n > 1e-4 * gmax:
        rows.append((rel_err(g[k], v), rel_err(g16[k], v), n, k))
rows.sort(reverse=True)
print(f"{'ours':>8} {'autocast':>8} {'|g|':>10}  tensor")
for r in rows[:25]:
    print(f"{r[0]:8.4f} {r[1]:8.4f} {r[2]:10.3e}  {r[3]}")
print("...")
for r in rows[-8:]:
    print(f"{r[0]:8.4f} {r[1]:8.4f} {r[2]:10.3e}  {r[3]}")
print("median ours", statistics.median(r[0] for r in rows), "median autocast", statistics.median(r[1] for r in rows), "n", len(rows))
tot = torch.cat([g[k].reshape(-1) for k in g32]); tot32 = torch.cat([g32[k].reshape(-1) for k in g32]); tot16 = torch.cat([g16[k].reshape(-1) for k in g32])
print("whole-gradient rel err ours", rel_err(tot, tot32), "autocast", rel_err(tot16, tot32), "cos ours", float(torch.nn.functional.cosine_similarity(tot, tot32, dim=0)))
