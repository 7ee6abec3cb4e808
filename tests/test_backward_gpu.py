"""Backward parity of the CUDA path (bf16 activation storage, fp32 accumulation) through the C ABI.

Stated bf16 bound (north_star: "a stated looser bound for bf16").  Two fp32 evaluations of this graph already differ
by 2e-3 (tests/test_oracle_golden.py).  In bf16 the deep layers are intrinsically noisy: the head divides by T=0.05 and
applies softmax twice, and ~60 BatchNorm layers renormalise with batch statistics, so PyTorch's own bf16 autocast of
the reference lands at a whole-gradient relative error of ~0.14 against fp32 (measured on B200, B=8, 120x160;
profiles/r01_gradient_noise.txt).  We therefore require
  (1) tensors near the loss (classifier, last dense block), where little noise has accumulated: <= 2e-2 norm-wise
      (3e-2 on the tiny reference fixtures, whose bottleneck BatchNorm sees only 2..15 samples);
  (2) the whole gradient: cosine similarity with the fp32 oracle >= 0.985 and a relative error no larger than 1.10x
      that of torch.autocast(bf16) running the oracle on the same inputs — i.e. at least as accurate as stock bf16;
  (3) per-tensor: no tensor may be more than 1.5x worse than autocast + 0.15 (catches a broken kernel, whose error
      would be O(1) while autocast stays at its noise level).
"""
import glob
import os

import pytest
import torch

from oracle import fcdensenet as net
from oracle import steps
from oracle.make_golden import make_drop_masks
from sim2real_lane_segment_b200.engine import Engine
from tests.util import flatten_state_dict, install_drop_masks, rel_err, unflatten_params

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.pt")))
NEAR_LOSS = ["classifier.finalConv.weight", "classifier.finalConv.bias",
             "featureExtractor.denseBlocksUp.4.layers.4.conv.weight"]


def _run_cuda(fx, masks, options=None):
    B, H, W, C = fx["B"], fx["H"], fx["W"], fx["num_cls"]
    sd = net.init_state_dict(num_cls=C, seed=fx["seed"])
    params, bn = flatten_state_dict(sd, C, "cuda")
    eng = Engine(B, H, W, C, training=True).bind()
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    eng.pack_weights(params)
    if masks is not None:
        install_drop_masks(eng, masks)
    x = fx["x"].cuda()
    y = fx["y"].cuda()
    grads = torch.zeros_like(params)
    eng.forward(params, bn, x, train=True)
    eng.head(params, y=y, weighted=True, backward=1)
    eng.backward(params, grads)
    torch.cuda.synchronize()
    hist, cm, loss, bad = eng.scalars()
    return sd, eng, unflatten_params(grads.cpu(), C), float(loss[0] / loss[1])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_loss_and_near_loss_gradients_match_reference_fixture(path):
    fx = torch.load(path)
    masks = make_drop_masks(fx["B"], fx["drop_seed"], num_cls=fx["num_cls"]) if "drop_seed" in fx else None
    sd, eng, g, loss = _run_cuda(fx, masks)
    assert abs(loss - float(fx["train_loss"])) < 0.01, (loss, float(fx["train_loss"]))
    for k in NEAR_LOSS:
        e = rel_err(g[k], fx["grads"][k])
        print(f"{e:8.4f} {k}")
        assert e < 3e-2, (k, e)
    # every gradient is finite and of a sane magnitude.  (On these tiny fixtures the deep layers normalise over 2..15
    # samples, where bf16 gradients are dominated by rounding noise — torch's own bf16 autocast is off by >100% there,
    # see profiles/r01_gradient_noise.txt — so deep-layer directions are checked on a real-sized batch below.)
    gmax = max(fx["grad_norms"].values())
    for k, n in fx["grad_norms"].items():
        assert torch.isfinite(g[k]).all(), k
        assert float(g[k].norm()) < 10.0 * n + 0.1 * gmax, (k, float(g[k].norm()), n)


def _oracle_grads(sd, x, y, C, autocast, masks=None):
    s = {k: v.cuda() for k, v in sd.items()}
    tr = steps.Trainer(s, num_cls=C)
    m = {k: v.cuda() for k, v in masks.items()} if masks else None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        tr.step(x, y, drop_masks=m)
    return {k: v.float().cpu() for k, v in tr.grads().items()}


@pytest.mark.gpu
@pytest.mark.parametrize("with_drop", [False, True])
def test_whole_gradient_vs_fp32_oracle_and_autocast_envelope(with_drop):
    B, H, W, C = 8, 120, 160, 4
    x, y = steps.synthetic_batch(B, H, W, C, seed=5)
    fx = {"B": B, "H": H, "W": W, "num_cls": C, "seed": 3, "x": x, "y": y.to(torch.uint8)}
    masks = make_drop_masks(B, 99, num_cls=C) if with_drop else None
    sd, eng, g, loss = _run_cuda(fx, masks)
    g32 = _oracle_grads(sd, x.cuda(), y.cuda(), C, False, masks)
    g16 = _oracle_grads(sd, x.cuda(), y.cuda(), C, True, masks)
    names = list(g32.keys())
    cat = lambda d: torch.cat([d[k].reshape(-1) for k in names])
    mine, ref, ac = cat(g), cat(g32), cat(g16)
    e_mine, e_ac = rel_err(mine, ref), rel_err(ac, ref)
    cos = float(torch.nn.functional.cosine_similarity(mine, ref, dim=0))
    print(f"whole-gradient rel err: ours {e_mine:.4f}  torch bf16 autocast {e_ac:.4f}  cosine {cos:.5f}")
    assert cos >= 0.985
    assert e_mine <= 1.10 * e_ac, (e_mine, e_ac)
    for k in NEAR_LOSS:
        assert rel_err(g[k], g32[k]) < 2e-2, (k, rel_err(g[k], g32[k]))
    gmax = max(float(v.norm()) for v in g32.values())
    worst = []
    for k in names:
        if float(g32[k].norm()) > 1e-4 * gmax:
            em, ea = rel_err(g[k], g32[k]), rel_err(g16[k], g32[k])
            worst.append((em - 1.5 * ea, em, ea, k))
    worst.sort(reverse=True)
    print("largest excess over the autocast envelope:", worst[:5])
    assert worst[0][0] < 0.15, worst[:5]


def _spread(fx, opt_a, opt_b):
    """whole-gradient distance between two configurations, next to the run-to-run spread of the second one"""
    out = [_run_cuda(fx, None, o) for o in (opt_a, opt_b, opt_b)]
    assert abs(out[0][3] - out[1][3]) < 1e-6
    cat = lambda d: torch.cat([d[k].reshape(-1) for k in out[0][2]])
    a, b, b2 = cat(out[0][2]), cat(out[1][2]), cat(out[2][2])
    return rel_err(a, b), rel_err(b2, b)


@pytest.mark.gpu
def test_planar_activation_mirror_gives_the_same_gradients():
    """the per-layer input-gradient kernels read x either from the NHWC arenas or from the planar mirror the forward pass
    writes beside them (plan option "xp"): same values, so the gradients agree up to the order of the fp32 atomics"""
    B, H, W, C = 4, 64, 96, 4
    x, y = steps.synthetic_batch(B, H, W, C, seed=9)
    fx = {"B": B, "H": H, "W": W, "num_cls": C, "seed": 4, "x": x, "y": y.to(torch.uint8)}
    d_ab, d_bb = _spread(fx, {"xp": 0}, {"xp": 1, "bwd_fused": 0})
    print(f"whole gradient, mirror off vs on: {d_ab:.2e}; on vs on (run-to-run): {d_bb:.2e}")
    assert d_ab <= 3.0 * d_bb + 1e-3


@pytest.mark.gpu
def test_block_fused_backward_matches_the_per_layer_backward():
    """plan option "bwd_fused": DenseBlocks through dgrad_bulk / wgrad_dense2 vs five per-layer launches.  Same
    arithmetic up to (i) one bf16 rounding of the block-input gradient per block instead of one per layer and (ii) the
    m*xhat operand split of the weight gradient, so the two whole gradients differ by bf16-rounding noise only."""
    B, H, W, C = 4, 64, 96, 4
    x, y = steps.synthetic_batch(B, H, W, C, seed=9)
    fx = {"B": B, "H": H, "W": W, "num_cls": C, "seed": 4, "x": x, "y": y.to(torch.uint8)}
    d_ab, d_bb = _spread(fx, {"bwd_fused": 0}, {"bwd_fused": 1})
    print(f"whole gradient, per-layer vs block-fused: {d_ab:.2e}; fused vs fused (run-to-run): {d_bb:.2e}")
    assert d_ab <= 0.06
