"""Edge geometry and the remaining reference call paths on the GPU: column strips (W > 160), odd sizes, camera
resolution inference (BASELINE.json configs[3] shape at a small batch), the MME entropy step with GradReverse."""
import pytest
import torch

from oracle import fcdensenet as net
from oracle import steps
from sim2real_lane_segment_b200.engine import Engine
from tests.util import flatten_state_dict, rel_err, unflatten_params


def _eval_vs_oracle(B, H, W, C, seed):
    sd = net.init_state_dict(num_cls=C, seed=seed)
    x, y = steps.synthetic_batch(B, H, W, C, seed=seed)
    params, bn = flatten_state_dict(sd, C, "cuda")
    eng = Engine(B, H, W, C, training=False).bind()
    eng.pack_weights(params)
    eng.forward(params, bn, x.cuda(), train=False)
    probs = torch.empty(B, C, H, W, device="cuda")
    mask = torch.empty(B, H, W, dtype=torch.uint8, device="cuda")
    eng.head(params, y=y.cuda(), weighted=False, probs=probs, mask=mask)
    torch.cuda.synchronize()
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref, _ = net.forward(sdc, x.cuda())
    assert float((probs - ref).abs().mean()) < 0.01
    agree = float((mask == ref.argmax(1).to(torch.uint8)).float().mean())
    assert agree > 0.97, agree
    hist, cm, loss, bad = eng.scalars()
    assert torch.equal(cm[:C * C].cpu().view(C, C).long(), steps.confusion_matrix(mask.cpu().long(), y, C))
    return agree


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 64, 200), (2, 35, 330), (1, 480, 640), (3, 33, 47)])
def test_eval_forward_strips_and_odd_sizes(shape):
    B, H, W = shape
    print(shape, _eval_vs_oracle(B, H, W, 4, seed=H + W))


@pytest.mark.gpu
def test_training_step_with_column_strips():
    """W = 200 -> two column strips in the 3x3 forward and input-gradient kernels"""
    B, H, W, C = 2, 96, 200, 4
    sd = net.init_state_dict(num_cls=C, seed=12)
    x, y = steps.synthetic_batch(B, H, W, C, seed=12)
    params, bn = flatten_state_dict(sd, C, "cuda")
    eng = Engine(B, H, W, C, training=True).bind()
    eng.pack_weights(params)
    grads = torch.zeros_like(params)
    eng.forward(params, bn, x.cuda(), train=True)
    eng.head(params, y=y.cuda(), weighted=True, backward=1)
    eng.backward(params, grads)
    torch.cuda.synchronize()
    g = unflatten_params(grads.cpu(), C)
    tr = steps.Trainer({k: v.cuda() for k, v in sd.items()}, num_cls=C)
    loss_ref, _, _ = tr.step(x.cuda(), y.cuda())
    hist, cm, loss, bad = eng.scalars()
    assert abs(float(loss[0] / loss[1]) - loss_ref) < 5e-3
    go = {k: v.float().cpu() for k, v in tr.grads().items()}
    for k, tol in (("classifier.finalConv.weight", 3e-2), ("featureExtractor.denseBlocksUp.4.layers.4.conv.weight", 3e-2),
                   ("featureExtractor.denseBlocksUp.4.layers.0.conv.weight", 8e-2),
                   ("featureExtractor.denseBlocksUp.4.layers.0.norm.weight", 8e-2)):
        assert rel_err(g[k], go[k]) < tol, (k, rel_err(g[k], go[k]))
    cat = lambda d: torch.cat([d[k].reshape(-1) for k in go])
    assert float(torch.nn.functional.cosine_similarity(cat(g), cat(go), dim=0)) > 0.97


@pytest.mark.gpu
def test_mme_unlabelled_step_gradient_reversal():
    from sim2real_lane_segment_b200.trainingModules import MMETrainingModule
    C, B, H, W = 4, 4, 96, 128
    sd = net.init_state_dict(num_cls=C, seed=17)
    xl, yl = steps.synthetic_batch(B, H, W, C, seed=3)
    xu, _ = steps.synthetic_batch(B, H, W, C, seed=4)
    m = MMETrainingModule(num_cls=C)
    m.load_state_dict(sd)
    m = m.cuda().train()
    rt = m.runtime()
    rt.dropout_p = 0.0
    (optG, optF), _ = m.configure_optimizers()
    batch = (xl.cuda(), xu.cuda(), yl.cuda(), None)
    optG.zero_grad()
    loss = m.training_step(batch, 0, optimizer_idx=0)
    loss.backward()
    torch.cuda.synchronize()
    g = unflatten_params(rt.grads.cpu().clone(), C)
    # oracle on the GPU in fp32
    sdc = {k: v.cuda().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    ref_loss, _ = steps.mme_unlabelled_loss(sdc, xu.cuda())
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 2e-3, (float(loss), float(ref_loss))
    for k in ("classifier.finalConv.weight", "classifier.finalConv.bias"):
        assert rel_err(g[k], sdc[k].grad.cpu()) < 3e-2, (k, rel_err(g[k], sdc[k].grad.cpu()))
    k = "featureExtractor.denseBlocksUp.4.layers.4.conv.weight"
    e = rel_err(g[k], sdc[k].grad.cpu())
    assert e < 4e-2, (k, e)   # the oracle's extractor gradient is already the reversed one
    # SGD-nesterov step over two lr groups, then the labelled AdamW step, run end to end
    optG.step()
    optF.zero_grad()
    loss2 = m.training_step(batch, 0, optimizer_idx=1)
    loss2.backward()
    optF.step()
    torch.cuda.synchronize()
    assert torch.isfinite(rt.params).all()
