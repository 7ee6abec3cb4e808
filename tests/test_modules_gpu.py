"""The reference-facing Python surface (trainingModules / models mirrors) on the GPU: state_dict interchange,
forward / evaluate_batch / training_step + FusedAdamW against the oracle, loss trajectory over 100 steps."""
import pytest
import torch

from oracle import fcdensenet as net
from oracle import steps


def _module(cls, sd, C):
    m = cls(num_cls=C)
    missing = m.load_state_dict({k: v.clone() for k, v in sd.items()})
    return m.cuda()


@pytest.mark.gpu
def test_forward_and_evaluate_batch_match_oracle():
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
    C = 4
    sd = net.init_state_dict(num_cls=C, seed=21)
    x, y = steps.synthetic_batch(4, 64, 96, C, seed=8)
    m = _module(SimpleTrainModule, sd, C).eval()
    with torch.no_grad():
        probs = m(x.cuda())
    ref, _ = net.forward(sd, x)
    assert probs.shape == ref.shape and probs.dtype == torch.float32
    assert float((probs.cpu() - ref).abs().mean()) < 0.01
    out = m.evaluate_batch((x.cuda(), y.cuda()), return_predictions=True)
    # integer metrics are bit-exact functions of our own predictions (PL-1.2.1 semantics restated in the oracle)
    lh = out["labels_hat"].long().cpu()
    assert torch.equal(lh, out["probs"].cpu().argmax(1))
    cm = steps.confusion_matrix(lh, y, C)
    assert torch.equal(out["confusion_matrix"].cpu(), cm)
    w = 4
    assert abs(float(out["acc"]) / w - float(steps.accuracy(lh, y))) < 1e-6
    assert abs(float(out["iou"]) / w - float(steps.iou(lh, y))) < 1e-6
    assert abs(float(out["dice"]) / w - float(steps.dice_from_cm(cm))) < 1e-6
    # and close to the fp32 oracle's evaluation of the same batch
    o = steps.evaluate_batch(sd, x, y)
    assert abs(float(out["loss"]) - float(o["loss"])) / w < 0.01
    assert abs(float(out["iou"]) - float(o["iou"])) / w < 0.02
    # state_dict round trip keeps the reference's keys and values
    sd2 = m.state_dict()
    assert list(sd2.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(sd2[k].cpu(), sd[k]), k


@pytest.mark.gpu
def test_loss_trajectory_100_steps():
    """fwd + weighted CE + bwd + AdamW, 100 steps on one fixed batch: CUDA path (bf16) vs oracle (fp32, on the GPU)"""
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
    C, B, H, W = 4, 8, 64, 96
    sd = net.init_state_dict(num_cls=C, seed=33)
    x, y = steps.synthetic_batch(B, H, W, C, seed=9)
    xc, yc = x.cuda(), y.cuda()
    m = _module(SimpleTrainModule, sd, C).train()
    m.runtime().dropout_p = 0.0
    (opt,), _ = m.configure_optimizers()
    tr = steps.Trainer({k: v.cuda() for k, v in sd.items()}, num_cls=C, lr=1e-3, decay=1e-4)
    mine, ref = [], []
    for i in range(100):
        opt.zero_grad()
        loss = m.training_step((xc, yc), i)
        loss.backward()
        opt.step()
        mine.append(float(loss))
        ref.append(tr.step(xc, yc)[0])
    print("step   ours    oracle")
    for i in (0, 1, 2, 5, 10, 20, 50, 99):
        print(f"{i:4d} {mine[i]:.5f} {ref[i]:.5f}")
    assert abs(mine[0] - ref[0]) < 5e-3
    dev = max(abs(a - b) for a, b in zip(mine, ref))
    drop = ref[0] - min(ref)
    assert mine[-1] < mine[0] - 0.5 * drop          # it trains
    assert dev < 0.25 * drop + 0.01, (dev, drop)     # and follows the fp32 trajectory
    # BN running statistics and counters advanced like torch's
    sdm = m.state_dict()
    k = "featureExtractor.denseBlocksDown.0.layers.0.norm.num_batches_tracked"
    assert int(sdm[k]) == 100
    k = "featureExtractor.denseBlocksDown.0.layers.0.norm.running_mean"
    torch.testing.assert_close(sdm[k].cpu(), tr.sd[k].cpu(), rtol=0.05, atol=0.02)


@pytest.mark.gpu
def test_generic_autograd_path_matches_fused_loss():
    """probs = model(x); F.cross_entropy(probs, y, weight) through torch autograd  ==  fused training_step"""
    import torch.nn.functional as F
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule, getClassWeight
    C = 4
    sd = net.init_state_dict(num_cls=C, seed=4)
    x, y = steps.synthetic_batch(2, 64, 64, C, seed=2)
    xc, yc = x.cuda(), y.cuda()
    m = _module(SimpleTrainModule, sd, C).train()
    rt = m.runtime()
    rt.dropout_p = 0.0
    m.prepare()
    rt.grads.zero_()
    loss_a = m.training_step((xc, yc), 0)
    loss_a.backward()
    ga = rt.grads.clone()
    # restore BN running stats side effects do not matter for gradients; rerun through generic autograd
    rt.grads.zero_()
    probs = m(xc)
    loss_b = F.cross_entropy(probs, yc, weight=getClassWeight(y, C).cuda())
    loss_b.backward()
    gb = rt.grads.clone()
    assert abs(float(loss_a) - float(loss_b)) < 1e-4
    assert float((ga - gb).norm() / gb.norm()) < 2e-2


@pytest.mark.gpu
def test_device_prefetcher_delivers_every_batch_in_order():
    from sim2real_lane_segment_b200.prefetch import DevicePrefetcher
    host = [(torch.randn(4, 3, 32, 32).pin_memory(), torch.randint(0, 4, (4, 32, 32)).pin_memory()) for _ in range(7)]
    got = []
    for x, y in DevicePrefetcher(host):
        assert x.is_cuda and y.is_cuda
        # consume on the current stream (a slow op keeps the buffer busy while the next copy is staged)
        got.append(((x * 1.0).clone(), y.clone(), float((x @ x.transpose(-1, -2)).sum())))
    torch.cuda.synchronize()
    assert len(got) == 7
    for (hx, hy), (dx, dy, _) in zip(host, got):
        assert torch.equal(dx.cpu(), hx) and torch.equal(dy.cpu(), hy)
    assert list(DevicePrefetcher([])) == []
