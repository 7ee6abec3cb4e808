"""Context numbers on one B200 (not the bench contract):
  * inference, BASELINE.json configs[3]: 256 images at 3x480x640 through TrainingBase.evaluate_batch in chunks of 32
    (masks + 4x4 confusion counts), this library vs stock PyTorch (oracle graph, cuDNN) fp32 and bf16 autocast
  * training, configs[1]: B=64, 3x120x160, this library vs stock PyTorch eager (oracle graph + torch.optim.AdamW)
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fcdensenet as net, steps
from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule

dev = torch.device("cuda")
out = {}

def timed(fn, n, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---------------------------------------------------------------- inference 480x640
C, H, W, CH, N = 4, 480, 640, 32, 256
sd = net.init_state_dict(num_cls=C, seed=42)
m = SimpleTrainModule(num_cls=C); m.load_state_dict(sd); m = m.cuda().eval()
x, y = steps.synthetic_batch(CH, H, W, C, seed=1)
xs = [x.cuda() for _ in range(2)]; ys = [y.cuda() for _ in range(2)]
cm_total = torch.zeros(C, C, dtype=torch.long, device=dev)
def ours_chunk():
    global cm_total
    o = m.evaluate_batch((xs[0], ys[0]), return_predictions=True)
    cm_total += o["confusion_matrix"]
ms = timed(ours_chunk, N // CH)
out["inference_480x640_ours_img_s"] = CH / (ms / 1e3)
del m; torch.cuda.empty_cache()
sdc = {k: v.cuda() for k, v in sd.items()}
def torch_chunk(autocast):
    def f():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            p, _ = net.forward(sdc, xs[0][:8])
            p.argmax(1)
    return f
out["inference_480x640_torch_fp32_img_s"] = 8 / (timed(torch_chunk(False), 4) / 1e3)
out["inference_480x640_torch_bf16_autocast_img_s"] = 8 / (timed(torch_chunk(True), 4) / 1e3)
torch.cuda.empty_cache()

# ---------------------------------------------------------------- training 120x160, B=64
B, H, W = 64, 120, 160
x, y = steps.synthetic_batch(B, H, W, C, seed=2); x, y = x.cuda(), y.cuda()
m = SimpleTrainModule(num_cls=C); m.load_state_dict(sd); m = m.cuda().train()
(opt,), _ = m.configure_optimizers()
def ours_step():
    opt.zero_grad(); l = m.training_step((x, y), 0); l.backward(); opt.step()
out["train_b64_ours_img_s"] = B / (timed(ours_step, 10, 3) / 1e3)
del m, opt; torch.cuda.empty_cache()
for name, ac in (("fp32", False), ("bf16_autocast", True)):
    tr = steps.Trainer({k: v.cuda() for k, v in sd.items()}, num_cls=C)
    def torch_step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            tr.opt.zero_grad(set_to_none=True)
            loss, acc, lh, nr = steps.training_step(tr.sd, x, y, C)
        loss.backward(); tr.opt.step()
    out[f"train_b64_torch_eager_{name}_img_s"] = B / (timed(torch_step, 5, 2) / 1e3)
    del tr; torch.cuda.empty_cache()
print(json.dumps({k: round(v, 1) for k, v in out.items()}, indent=1))
