"""Input pipeline: oracle vs cv2-generated fixtures on CPU; rl_preprocess vs oracle/fixtures on GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as pp

PREP = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prep_*.npz")))


@pytest.mark.parametrize("path", PREP, ids=os.path.basename)
def test_oracle_restates_cv2_bit_exactly(path):
    fx = np.load(path)
    H, W = int(fx["H"]), int(fx["W"])
    assert np.array_equal(pp.resize_linear_u8(fx["img"], H, W), fx["linear"])
    assert np.array_equal(pp.resize_nearest_u8(fx["label"], H, W), fx["nearest"])
    assert np.array_equal(pp.to_gray_u8(fx["linear"]), fx["gray"])
    x, y = pp.my_transform(fx["img"], fx["label"], width=W, height=H)
    assert x.dtype == np.float32 and x.shape == (3, H, W) and y.dtype == np.int64
    assert np.array_equal(x, fx["norm"].transpose(2, 0, 1))
    assert np.array_equal(y, fx["nearest"].astype(np.int64))


def test_oracle_edge_cases():
    rng = np.random.default_rng(3)
    one = rng.integers(0, 256, (1, 1, 3), dtype=np.uint8)           # 1x1 source: every output pixel is that pixel
    assert np.all(pp.resize_linear_u8(one, 5, 7) == one[0, 0])
    img = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    assert np.array_equal(pp.resize_linear_u8(img, 7, 9), img)      # same size: copy
    x, y = pp.my_transform(img, None, width=9, height=7)
    assert y is None
    flat = np.full((16, 16, 3), 200, dtype=np.uint8)                # constant image stays constant at any scale
    assert np.all(pp.resize_linear_u8(flat, 5, 31) == 200)


@pytest.mark.gpu
@pytest.mark.parametrize("path", PREP, ids=os.path.basename)
@pytest.mark.parametrize("gray", [False, True])
def test_rl_preprocess_matches_cv2_fixture(path, gray):
    from sim2real_lane_segment_b200.dataManagement.myTransforms import MyTransform
    fx = np.load(path)
    H, W = int(fx["H"]), int(fx["W"])
    x, y = MyTransform(width=W, height=H, gray=gray)(fx["img"], fx["label"])
    want = pp.normalize(fx["gray"] if gray else fx["linear"]).transpose(2, 0, 1)
    assert torch.equal(x.cpu(), torch.from_numpy(np.ascontiguousarray(want)))          # bit-exact, float32
    assert torch.equal(y.cpu(), torch.from_numpy(fx["nearest"].astype(np.int64)))


@pytest.mark.gpu
def test_rl_preprocess_batch_and_ragged_sizes():
    from sim2real_lane_segment_b200.dataManagement.myTransforms import MyTransform
    rng = np.random.default_rng(11)
    for (B, Hs, Ws, H, W) in [(5, 48, 64, 120, 160), (3, 101, 77, 33, 45), (2, 1, 1, 4, 4), (64, 120, 160, 120, 160)]:
        img = rng.integers(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
        lab = rng.integers(0, 4, (B, Hs, Ws), dtype=np.uint8)
        x, y = MyTransform(width=W, height=H)(img, lab)
        for b in range(B):
            wx, wy = pp.my_transform(img[b], lab[b], width=W, height=H)
            assert torch.equal(x[b].cpu(), torch.from_numpy(wx)), (B, Hs, Ws, H, W, b)
            assert torch.equal(y[b].cpu(), torch.from_numpy(wy))
    x, y = MyTransform()(img[0])            # no label: returned unchanged (None)
    assert y is None and x.shape == (3, 120, 160)


@pytest.mark.gpu
def test_overlay_and_demo_frames_match_the_oracle():
    """makeDemoVideo body: frame -> MyTransform -> forward -> argmax -> painted, resized frame.  The overlay is checked
    bit-exactly against the oracle fed with OUR mask; the mask against the argmax of our own probabilities."""
    import ctypes
    from sim2real_lane_segment_b200 import _lib
    from sim2real_lane_segment_b200.dataManagement.myTransforms import MyTransform
    from sim2real_lane_segment_b200.makeDemoVideo import predictFrames
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule
    rng = np.random.default_rng(5)
    # kernel alone, both mask dtypes
    L = _lib.lib()
    frames = rng.integers(0, 256, (3, 90, 130, 3), dtype=np.uint8)
    pred = rng.integers(0, 4, (3, 60, 80), dtype=np.uint8)
    want = np.stack([pp.overlay(frames[b], pred[b], width=80, height=60) for b in range(3)])
    fr = torch.from_numpy(frames).cuda()
    for dt, flag in ((torch.uint8, 0), (torch.int64, 1)):
        mk = torch.from_numpy(pred).to(dt).cuda()
        out = torch.empty(3, 60, 80, 3, dtype=torch.uint8, device="cuda")
        rc = L.rl_overlay(ctypes.c_void_p(fr.data_ptr()), 3, 90, 130, 60, 80, ctypes.c_void_p(mk.data_ptr()), flag,
                          ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        assert np.array_equal(out.cpu().numpy(), want)
    # whole consumer path
    torch.manual_seed(0)
    model = SimpleTrainModule(num_cls=4).cuda().eval()
    frames = rng.integers(0, 256, (2, 240, 320, 3), dtype=np.uint8)
    tf = MyTransform(width=160, height=120)
    out, mask = predictFrames(frames, model, tf)
    x, _ = tf(frames)
    with torch.no_grad():
        _, ref_mask = torch.max(model.forward(x), 1)
    assert torch.equal(mask.long(), ref_mask)
    want = np.stack([pp.overlay(frames[b], mask[b].cpu().numpy(), width=160, height=120) for b in range(2)])
    assert np.array_equal(out.cpu().numpy(), want)
