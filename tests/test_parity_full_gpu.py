"""Whole-network parity at the sizes BASELINE.json quotes: a B=64 120x160 training step, a 4x480x640 evaluation with a
committed mask-mismatch count, and a loss trajectory over fresh batches with Dropout2d active (masks injected so that the
CUDA path and the fp32 oracle drop the same channels)."""
import pytest
import torch

from oracle import fcdensenet as net
from oracle import steps
from oracle.make_golden import make_drop_masks
from tests.util import install_drop_masks, rel_err, unflatten_params

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    """the oracle runs on the GPU here (sizes the CPU cannot finish in seconds): true fp32, no TF32"""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _module(sd, C, train):
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
    m = SimpleTrainModule(num_cls=C)
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m = m.cuda()
    return m.train() if train else m.eval()


def test_training_step_at_batch_64():
    """configs[1] at full size: loss, accuracy and gradients of ONE step (dropout masks injected) vs the fp32 oracle and vs
    torch's own bf16 autocast of the oracle"""
    C, B, H, W = 4, 64, 120, 160
    sd = net.init_state_dict(num_cls=C, seed=11)
    x, y = steps.synthetic_batch(B, H, W, C, seed=12)
    xc, yc = x.cuda(), y.cuda()
    masks = make_drop_masks(B, 77, num_cls=C)
    m = _module(sd, C, True)
    rt = m.runtime()
    rt.dropout_mask_hook = lambda eng: install_drop_masks(eng, masks)
    m.prepare()
    rt.grads.zero_()
    loss = m.training_step((xc, yc), 0)
    loss.backward()
    torch.cuda.synchronize()
    mine_named = unflatten_params(rt.grads.clone(), C)
    res = {}
    for name, ac in (("fp32", False), ("bf16", True)):
        tr = steps.Trainer({k: v.cuda() for k, v in sd.items()}, num_cls=C)
        mk = {k: v.cuda() for k, v in masks.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            l, acc, _, _ = steps.training_step(tr.sd, xc, yc, C, drop_masks=mk)
        l.backward()
        g = torch.cat([tr.sd[n].grad.reshape(-1).float() for n in tr.param_names])
        res[name] = (float(l), float(acc), g)
        names = list(tr.param_names)
        del tr
    mine = torch.cat([mine_named[n].reshape(-1) for n in names])
    ref, ac = res["fp32"][2], res["bf16"][2]
    e_mine, e_ac = rel_err(mine, ref), rel_err(ac, ref)
    cos = float(torch.nn.functional.cosine_similarity(mine, ref, dim=0))
    print(f"B=64 loss: ours {float(loss):.5f} fp32 {res['fp32'][0]:.5f} autocast {res['bf16'][0]:.5f}; tr_acc ours "
          f"{float(m.logged['tr_acc']):.3f} fp32 {res['fp32'][1]:.3f}")
    print(f"whole gradient rel err: ours {e_mine:.4f}, torch bf16 autocast {e_ac:.4f}; cosine {cos:.5f}")
    assert abs(float(loss) - res["fp32"][0]) < 2e-3
    assert abs(float(m.logged["tr_acc"]) - res["fp32"][1]) < 0.5          # percent of pixels
    assert cos >= 0.99 and e_mine <= 1.10 * e_ac


# (seed, committed mismatch count of the argmax masks vs the fp32 oracle out of 4*480*640 = 1 228 800 pixels, measured on
# B200; the eval path has no atomics, so the count is reproducible)
EVAL_FIXTURES = [(31, 2400), (32, 2300)]   # measured: 2257, 2185 (0.18 %: a random-init net is near-tied almost everywhere)


@pytest.mark.parametrize("seed,committed", EVAL_FIXTURES)
def test_eval_480x640_mask_mismatch_count(seed, committed):
    """configs[3] geometry: bf16 activations move a few probabilities across a tie; every mismatching pixel must be a near
    tie of the fp32 probabilities, and their number must not exceed the committed count"""
    C, B, H, W = 4, 4, 480, 640
    sd = net.init_state_dict(num_cls=C, seed=seed)
    for k in sd:                          # running statistics as after some training: not the (0, 1) initial values
        if k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(sd[k].shape, generator=torch.Generator().manual_seed(seed))
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=torch.Generator().manual_seed(seed + 1))
    x, y = steps.synthetic_batch(B, H, W, C, seed=seed + 2)
    m = _module(sd, C, False)
    out = m.evaluate_batch((x.cuda(), y.cuda()), return_predictions=True)
    with torch.no_grad():
        ref, _ = net.forward({k: v.cuda() for k, v in sd.items()}, x.cuda())
    mask, rmask = out["labels_hat"].long(), ref.argmax(1)
    mism = mask != rmask
    n = int(mism.sum())
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[mism]
    cm_ref = steps.confusion_matrix(rmask.cpu(), y, C)
    iou_ours, iou_ref = float(out["iou"]) / B, float(steps.iou(rmask.cpu(), y))
    print(f"seed {seed}: {n} of {mask.numel()} mask pixels differ from the fp32 oracle ({100.0 * n / mask.numel():.4f} %), "
          f"largest fp32 margin among them {float(margin.max()) if n else 0:.4f}; IoU ours {iou_ours:.5f} fp32 {iou_ref:.5f}; "
          f"max |dp| {float((out['probs'] - ref).abs().max()):.4f}")
    assert n <= committed, (n, committed)
    assert n == 0 or float(margin.max()) < 0.02          # only near ties flip (measured: 0.0067)
    assert abs(iou_ours - iou_ref) < 2e-3
    assert int((out["confusion_matrix"].cpu() - cm_ref).abs().sum()) <= 2 * n


def test_loss_trajectory_fresh_batches_with_dropout():
    """60 optimizer steps, a different batch every step (12 distinct ones), Dropout2d(0.2) active with the same keep masks
    in both runs: the bf16 CUDA path follows the fp32 oracle's loss curve while the loss is still falling"""
    C, B, H, W = 4, 16, 120, 160
    sd = net.init_state_dict(num_cls=C, seed=44)
    batches = [steps.synthetic_batch(B, H, W, C, seed=100 + i) for i in range(12)]
    batches = [(x.cuda(), y.cuda()) for x, y in batches]
    m = _module(sd, C, True)
    rt = m.runtime()
    (opt,), _ = m.configure_optimizers()
    tr = steps.Trainer({k: v.cuda() for k, v in sd.items()}, num_cls=C, lr=1e-3, decay=1e-4)
    state = {}
    rt.dropout_mask_hook = lambda eng: install_drop_masks(eng, state["masks"])
    mine, ref = [], []
    for i in range(60):
        x, y = batches[i % 12]
        state["masks"] = make_drop_masks(B, 1000 + i, num_cls=C)
        opt.zero_grad()
        loss = m.training_step((x, y), i)
        loss.backward()
        opt.step()
        mine.append(float(loss))
        ref.append(tr.step(x, y, drop_masks={k: v.cuda() for k, v in state["masks"].items()})[0])
    dev = [abs(a - b) for a, b in zip(mine, ref)]
    print("step   ours    oracle")
    for i in (0, 1, 2, 5, 10, 20, 30, 40, 50, 59):
        print(f"{i:4d} {mine[i]:.5f} {ref[i]:.5f}")
    drop = ref[0] - min(ref)
    tail = abs(sum(mine[-10:]) - sum(ref[-10:])) / 10
    print(f"oracle loss drop {drop:.4f}; max |ours - oracle| {max(dev):.4f}, mean {sum(dev) / len(dev):.4f}, first 10 steps "
          f"{max(dev[:10]):.2e}, first 20 {max(dev[:20]):.2e}, mean of the last 10 steps differs by {tail:.4f}")
    assert ref[-1] > 0.76 + 0.02, "the trajectory must not have saturated at the 0.744 lower bound"
    assert drop > 0.08
    # Both runs are chaotic and neither is reproducible (fp32 atomics here, cuDNN / atomics in the oracle).  Over 18 runs on
    # B200: the ORACLE's own drop 0.104 .. 0.135; largest single-step deviation 0.029 .. 0.064; mean deviation 0.0053 ..
    # 0.0079; first 10 steps 2.6e-4 .. 3.8e-4; first 20 steps 1.0e-3 .. 1.9e-3 (a 2e-3 bound on those failed about once in
    # twelve runs); mean of the last ten steps 0.002 .. 0.016 apart.  The parity statement is the agreement before the
    # trajectories decorrelate (first 10 steps) and of the averaged curves afterwards; single-step maxima are not asserted.
    assert max(dev[:10]) < 3e-3
    assert sum(dev) / len(dev) < 0.02
    assert tail < 0.04
