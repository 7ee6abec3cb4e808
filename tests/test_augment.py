"""Random augmentations of the input pipeline (reference myTransforms.py:8-20).
CPU: the oracle restates cv2 bit-exactly (committed cv2-generated fixtures; exhaustive colour-conversion and line checks
against cv2 itself when it is importable); the host mirror draws the same parameters as the oracle from the same `random`
stream.  GPU: rl_preprocess_augment (through MyTransform(augment=True, params=...)) against fixtures and oracle, bit-exact."""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import augment as A
from oracle import preprocess as pp

AUG = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "aug_*.npz")))
H, W = 120, 160


def fixture_params(fx, cls=A.AugmentParams):
    k = fx["kernel"]
    g = fx["gauss"]
    return cls(fx["lut_h"], fx["lut_s"], fx["lut_v"], int(fx["hsv"]), tuple(int(c) for c in fx["crop"]),
               k if k.size else None, g if g.size else None)


@pytest.mark.parametrize("path", AUG, ids=os.path.basename)
def test_oracle_restates_cv2_augmentations_bit_exactly(path):
    fx = np.load(path)
    P = fixture_params(fx)
    x, y = A.augment_u8(fx["img"], fx["label"], P, H, W)
    assert np.array_equal(x, fx["out_u8"])
    assert np.array_equal(y, fx["out_label"])
    xf, yl = A.my_transform_augment(fx["img"], fx["label"], P, width=W, height=H)
    assert xf.dtype == np.float32 and np.array_equal(xf, fx["norm"].transpose(2, 0, 1))
    assert np.array_equal(yl, fx["out_label"].astype(np.int64))
    xg, _ = A.my_transform_augment(fx["img"], fx["label"], P, width=W, height=H, gray=True)
    assert np.array_equal(xg, pp.normalize(fx["out_gray"]).transpose(2, 0, 1))


def test_colour_conversions_against_cv2_exhaustively():
    cv2 = pytest.importorskip("cv2")
    v = np.arange(256, dtype=np.uint8)
    for r0 in range(0, 256, 64):  # all 2^24 colours in four slabs
        r, g, b = np.meshgrid(v[r0:r0 + 64], v, v, indexing="ij")
        img = np.stack([r, g, b], -1).reshape(2048, 2048, 3)
        assert np.array_equal(A.rgb2hsv_u8(img), cv2.cvtColor(img, cv2.COLOR_RGB2HSV))
    for h0 in range(0, 180, 45):  # every (h < 180, s, v); widths are multiples of 64 so cv2 stays on its vector path
        h, s, vv = np.meshgrid(np.arange(h0, h0 + 45, dtype=np.uint8), v, v, indexing="ij")
        hsv = np.stack([h, s, vv], -1).reshape(45 * 32, 2048, 3)
        assert np.array_equal(A.hsv2rgb_u8(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB))


def test_motion_blur_line_and_filter_against_cv2():
    cv2 = pytest.importorskip("cv2")
    import itertools
    from sim2real_lane_segment_b200.dataManagement.myTransforms import line_kernel
    n6 = []
    for ks in (3, 5, 7):
        for xs, ys, xe, ye in itertools.product(range(ks), repeat=4):
            if (xs, ys) == (xe, ye):
                continue
            k = np.zeros((ks, ks), np.uint8)
            cv2.line(k, (xs, ys), (xe, ye), 1, thickness=1)
            ref = (k.astype(np.float32) / np.float32(k.sum())).astype(np.float32)
            assert np.array_equal(A.motion_blur_kernel(ks, xs, ys, xe, ye), ref), (ks, xs, ys, xe, ye)
            assert np.array_equal(line_kernel(ks, xs, ys, xe, ye), ref), (ks, xs, ys, xe, ye)
            if int(k.sum()) == 6 and len(n6) < 4:
                n6.append(ref)
    rng = np.random.default_rng(0)
    # 6-tap kernels (weight 1/6 is inexact in float32) on images made of exact .5 cases are where summation order shows
    for k in n6 + [A.motion_blur_kernel(3, 0, 0, 2, 2), A.motion_blur_kernel(5, 4, 0, 0, 3), A.motion_blur_kernel(7, 0, 6, 6, 1)]:
        for img in (rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (rng.integers(0, 4, (H, W, 3)) * 6 + 3).astype(np.uint8)):
            assert np.array_equal(A.filter2d_u8(img, k), cv2.filter2D(img, ddepth=-1, kernel=k))


def test_host_mirror_draws_what_the_oracle_draws():
    from sim2real_lane_segment_b200.dataManagement import myTransforms as T
    for seed in range(40):
        a = A.draw_params(480, 640, H, W, rng=random.Random(seed))
        b = T.draw_params(480, 640, H, W, rng=random.Random(seed))
        assert a.crop == b.crop and a.hsv == b.hsv
        for n in ("lut_h", "lut_s", "lut_v"):
            assert np.array_equal(getattr(a, n), getattr(b, n))
        assert (a.kernel is None) == (b.kernel is None) and (a.gauss is None) == (b.gauss is None)
        if a.kernel is not None:
            assert np.array_equal(a.kernel, b.kernel)
        else:
            assert np.array_equal(a.gauss, b.gauss)
    # parameter ranges of the reference's configuration (myTransforms.py:9-13)
    rng = random.Random(7)
    ps = [A.draw_params(480, 640, H, W, rng=rng) for _ in range(300)]
    hs = np.array([p.crop[3] for p in ps])
    assert hs.min() >= H // 2 and hs.max() <= 4 * H and all(p.crop[2] == int(p.crop[3] * (W / H)) for p in ps)
    assert all(p.crop[0] + p.crop[2] <= 640 and p.crop[1] + p.crop[3] <= 480 for p in ps)
    nb = sum(p.kernel is not None for p in ps)
    assert 100 < nb < 200                                              # OneOf: equal odds
    assert all(p.kernel.shape[0] in (3, 5, 7) and abs(float(p.kernel.sum()) - 1) < 1e-6 for p in ps if p.kernel is not None)
    sig = np.array([p.gauss.std() for p in ps if p.gauss is not None])
    assert sig.min() > 10 ** 0.5 * 0.95 and sig.max() < 50 ** 0.5 * 1.05
    with pytest.raises(ValueError):                                    # crop taller than the frame: albumentations raises
        for _ in range(50):
            A.draw_params(100, 640, H, W, rng=rng)
    raw, noise = T.pack_params(ps[:5], H, W)
    assert raw.shape == (5, T.PARAM_BYTES) and T.PARAM_BYTES == 768 + 4 * 6 + 49 * 4 + 4


@pytest.mark.gpu
@pytest.mark.parametrize("path", AUG, ids=os.path.basename)
@pytest.mark.parametrize("gray", [False, True])
def test_rl_preprocess_augment_matches_cv2_fixture(path, gray):
    from sim2real_lane_segment_b200.dataManagement import myTransforms as T
    fx = np.load(path)
    t = T.MyTransform(width=W, height=H, gray=gray, augment=True)
    x, y = t(fx["img"], fx["label"], params=[fixture_params(fx, T.AugmentParams)])
    want = pp.normalize(fx["out_gray"] if gray else fx["out_u8"]).transpose(2, 0, 1)
    assert torch.equal(x.cpu(), torch.from_numpy(np.ascontiguousarray(want)))
    assert torch.equal(y.cpu(), torch.from_numpy(fx["out_label"].astype(np.int64)))


@pytest.mark.gpu
def test_rl_preprocess_augment_batch_of_drawn_parameters():
    """a batch of 24 frames with independently drawn parameters (both OneOf branches, every kernel size) vs the oracle"""
    from sim2real_lane_segment_b200.dataManagement import myTransforms as T
    rng = np.random.default_rng(21)
    B, Hs, Ws = 24, 480, 640
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    base = np.stack([xx * 255 // (Ws - 1), yy * 255 // (Hs - 1), (xx + 2 * yy) % 256], -1)
    img = np.clip(base[None] + rng.integers(-60, 61, (B, Hs, Ws, 3)), 0, 255).astype(np.uint8)
    img[:, :40] = img[:, :40, :, :1]                                                             # grey band: s == 0
    lab = rng.integers(0, 4, (B, Hs, Ws), dtype=np.uint8)
    pr = random.Random(5)
    plist = [T.draw_params(Hs, Ws, H, W, rng=pr) for _ in range(B)]
    plist[0].lut_h, plist[0].lut_s, plist[0].lut_v, plist[0].hsv = np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), False
    assert any(p.kernel is not None for p in plist) and any(p.gauss is not None for p in plist)
    t = T.MyTransform(width=W, height=H, augment=True)
    x, y = t(img, lab, params=plist)
    for b in range(B):
        P = A.AugmentParams(plist[b].lut_h, plist[b].lut_s, plist[b].lut_v, plist[b].hsv, plist[b].crop, plist[b].kernel, plist[b].gauss)
        wx, wy = A.my_transform_augment(img[b], lab[b], P, width=W, height=H)
        assert torch.equal(x[b].cpu(), torch.from_numpy(wx)), b
        assert torch.equal(y[b].cpu(), torch.from_numpy(wy)), b
    # the un-injected call draws from the global `random` stream: reproducible under random.seed, shapes as the reference's
    random.seed(3)
    x1, y1 = t(img[:4], lab[:4])
    random.seed(3)
    x2, y2 = t(img[:4], lab[:4])
    assert torch.equal(x1, x2) and torch.equal(y1, y2) and x1.shape == (4, 3, H, W) and y1.dtype == torch.int64
    xs, ys = t(img[0], lab[0])
    assert xs.shape == (3, H, W) and ys.shape == (H, W)
    with pytest.raises(ValueError):
        t(img[:1], lab[:1], params=[T.AugmentParams(plist[1].lut_h, plist[1].lut_s, plist[1].lut_v, True, (600, 0, 160, 120))])
