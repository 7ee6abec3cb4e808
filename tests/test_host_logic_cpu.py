"""Host-side logic that needs no GPU: metric formulas, class weights, module surface, loud failure without CUDA,
bucketed gradient exchange on gloo (world_size 2)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import steps
from sim2real_lane_segment_b200 import _lib, ddp, spec
from sim2real_lane_segment_b200.trainingModules import MMETrainingModule, SimpleTrainModule, getClassWeight
from sim2real_lane_segment_b200.trainingModules.TrainingBase import metrics_from_confusion


def test_metrics_from_confusion_match_pl_semantics():
    g = torch.Generator().manual_seed(0)
    for trial in range(200):
        C = 4
        n = int(torch.randint(1, 400, (1,), generator=g))
        kmax = int(torch.randint(1, C + 1, (1,), generator=g))       # absent classes, data-dependent K for iou
        y = torch.randint(0, kmax, (n,), generator=g)
        p = torch.randint(0, kmax if trial % 2 else C, (n,), generator=g)
        cm = steps.confusion_matrix(p, y, C)
        acc, dice, iou = metrics_from_confusion(cm)
        assert abs(float(acc) - float(steps.accuracy(p, y))) < 1e-6
        assert abs(float(iou) - float(steps.iou(p, y))) < 1e-6
        assert abs(float(dice) - float(steps.dice_from_cm(cm))) < 1e-6


def test_get_class_weight_matches_reference_semantics():
    y = torch.tensor([[0, 0, 1], [1, 1, 3]])
    assert torch.equal(getClassWeight(y, 4), steps.get_class_weight(y, 4))
    assert torch.isinf(getClassWeight(y, 4)[2])
    with pytest.raises(AssertionError):
        getClassWeight(y, 3)


def test_module_surface_and_state_dict_keys():
    m = SimpleTrainModule(lr=2e-3, decay=1e-5, lrRatio=100, num_cls=4)
    assert (m.lr, m.decay, m.lrRatio, m.num_cls) == (2e-3, 1e-5, 100, 4)
    assert m.featureExtractor.getFeatureChannels() == 288
    keys = list(m.state_dict().keys())
    assert keys == [n for n, _, _ in spec.entries(4)] and len(keys) == 434
    assert sum(p.numel() for p in m.parameters()) == 3461220
    import argparse
    p = SimpleTrainModule.add_model_specific_args(argparse.ArgumentParser())
    a = p.parse_args(["-lr", "0.01", "--decay", "0.1"])
    assert a.learningRate == 0.01 and a.lrRatio == 1000
    assert MMETrainingModule(num_cls=3).classifier.finalConv.weight.shape == (3, 288, 1, 1)


def test_no_cpu_fallback():
    m = SimpleTrainModule(num_cls=4)
    with pytest.raises(_lib.RightLaneError):
        m(torch.zeros(1, 3, 64, 64))


def test_bucket_ranges_cover_exactly():
    for n in (1, 255, 256, 3461220, 1000003):
        for nb in (1, 2, 3, 7):
            r = ddp.bucket_ranges(n, nb)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo % 256 == 0 for lo, _ in r)


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    sync = ddp.GradSync(n_buckets=3)
    sync.start(g)
    scale = sync.finish()
    want = torch.arange(n, dtype=torch.float32) * sum(range(1, world + 1))
    assert scale == 1.0 / world
    assert torch.equal(g, want)
    # the ranges the backward pass fires while it is still running (Engine.backward_segments): async, awaited by finish()
    g2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
    sync.reduce_ranges(g2, [(n // 2, n)])
    sync.reduce_ranges(g2, [(0, n // 4), (n // 4, n // 2)])
    assert sync.finish() == 1.0 / world
    assert torch.equal(g2, want)
    buf = torch.full((10,), float(rank))
    cnt = [torch.tensor(rank * 5)]
    ddp.broadcast_buffers(buf, cnt)
    assert float(buf.sum()) == 0.0 and int(cnt[0]) == 0
    dist.destroy_process_group()


def test_gradient_exchange_world_size_2_gloo():
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, 100003), nprocs=2, join=True)


def test_backward_segments_partition_the_parameter_arena():
    """the three all-reduce buckets (up path + classifier, deep down path, shallow down path) cover every parameter
    exactly once and the segments walk the layer list from the classifier down to the first conv"""
    from sim2real_lane_segment_b200.engine import Engine
    eng = Engine(2, 120, 160, 4, training=True)      # host-side plan only: no device work before bind()
    segs = eng.backward_segments()
    n = eng.query(0)
    assert segs[0][0] == n and segs[-1][1] == 0
    for (h0, l0, _), (h1, l1, _) in zip(segs, segs[1:]):
        assert h1 == l0 - 1
    ranges = sorted(r for _, _, rs in segs for r in rs)
    assert ranges[0][0] == 0 and ranges[-1][1] == eng.n_params
    for (a0, b0), (a1, b1) in zip(ranges, ranges[1:]):
        assert b0 == a1 and b0 > a0
    pl, _, _, npar, _ = spec.layout(4)
    up_start = pl["featureExtractor.transUpBlocks.0.convTrans.weight"][0]
    assert segs[0][2] == [(up_start, npar)]
    assert segs[2][2][0] == (0, pl["featureExtractor.denseBlocksDown.2.layers.0.norm.weight"][0])
    eng.close()


def test_checkpoint_roundtrip_has_lightning_121_layout(tmp_path):
    """Trainer.save_checkpoint writes what `load_from_checkpoint(checkpoint_path=..., num_cls=4)` (test.py:28-30) reads:
    state_dict under the reference's 434 names, hyper_parameters, epoch / global_step"""
    import torch
    from oracle import fcdensenet as net
    from sim2real_lane_segment_b200.pl_shim import Trainer
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
    m = SimpleTrainModule(lr=2e-3, decay=1e-4, lrRatio=100, num_cls=4)
    path = Trainer().save_checkpoint(str(tmp_path / "m.ckpt"), model=m)
    ck = torch.load(path)
    assert {"state_dict", "hyper_parameters", "epoch", "global_step", "pytorch-lightning_version"} <= set(ck)
    assert list(ck["state_dict"].keys()) == list(net.init_state_dict(num_cls=4, seed=0).keys())
    # the reference saves lr / decay / lrRatio only (TrainingBase.py:29): num_cls travels as a load_from_checkpoint kwarg
    assert ck["hyper_parameters"] == {"lr": 2e-3, "decay": 1e-4, "lrRatio": 100}
    m2 = SimpleTrainModule.load_from_checkpoint(checkpoint_path=path, num_cls=4)
    assert m2.lr == 2e-3 and m2.lrRatio == 100
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
