"""Generates tests/golden/aug_*.npz with cv2 itself: every pixel operation of MyTransform(augment=True)
(myTransforms.py:8-20) executed by the library the reference's albumentations calls end up in — cv2.cvtColor
RGB2HSV / cv2.LUT / HSV2RGB, crop + cv2.resize INTER_LINEAR / INTER_NEAREST, cv2.filter2D, and albumentations 0.5.2's
gauss_noise formula (numpy) — for fixed, recorded parameters.  Run in the build container (cv2 4.13 installed):
    python oracle/make_golden_augment.py
The parameters are drawn by oracle.augment.draw_params from a seeded `random.Random` (kernel lines drawn by cv2.line)."""
import os
import random
import sys
import zlib

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import augment as A  # noqa: E402
from oracle import preprocess as pp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
H, W = 120, 160


def cv2_line_kernel(ksize, xs, ys, xe, ye):
    k = np.zeros((ksize, ksize), dtype=np.uint8)
    cv2.line(k, (xs, ys), (xe, ye), 1, thickness=1)
    return (k.astype(np.float32) / np.float32(np.sum(k))).astype(np.float32)


def cv2_pipeline(img, label, P):
    x = img
    if P.hsv:
        hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
        h, s, v = cv2.split(hsv)
        hsv = cv2.merge((cv2.LUT(h, P.lut_h), cv2.LUT(s, P.lut_s), cv2.LUT(v, P.lut_v)))
        x = cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)
    cx, cy, cw, ch = P.crop
    x = x[cy:cy + ch, cx:cx + cw]
    x = cv2.resize(x, (W, H), interpolation=cv2.INTER_LINEAR) if (ch, cw) != (H, W) else x
    y = label[cy:cy + ch, cx:cx + cw]
    y = cv2.resize(y, (W, H), interpolation=cv2.INTER_NEAREST) if (ch, cw) != (H, W) else y
    if P.kernel is not None:
        x = cv2.filter2D(x, ddepth=-1, kernel=P.kernel)
    if P.gauss is not None:
        x = np.clip(x.astype("float32") + P.gauss, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


# name, source size, forced crop height (None: drawn), forced branch (0 blur / 1 noise / None: drawn), forced blur line
CASES = [("480x640_drawn_a", 480, 640, None, None, None), ("480x640_drawn_b", 480, 640, None, None, None),
         ("480x640_half_blur6", 480, 640, 240, 0, (7, 0, 1, 5, 6)), ("480x640_full_noise", 480, 640, 480, 1, None),
         ("240x320_up_blur3", 240, 320, 60, 0, (3, 0, 2, 2, 0)), ("240x320_identity_noise", 240, 320, 120, 1, None)]
for name, Hs, Ws, force_h, force_branch, force_line in CASES:
    seed = zlib.crc32(name.encode())
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    base = np.stack([xx * 255 // (Ws - 1), yy * 255 // (Hs - 1), (xx * 3 + yy * 2) % 256], -1).astype(np.int64)
    img = np.clip(base + rng.integers(-25, 26, (Hs, Ws, 3)), 0, 255).astype(np.uint8)
    img[: Hs // 6] = img[: Hs // 6, :, :1]            # a grey band (s == 0 pixels)
    label = rng.integers(0, 4, (Hs // 8 + 1, Ws // 8 + 1), dtype=np.uint8).repeat(8, 0).repeat(8, 1)[:Hs, :Ws].copy()
    pr = random.Random(seed)
    while True:
        try:
            P = A.draw_params(Hs, Ws, H, W, rng=pr, line=cv2_line_kernel)
        except ValueError:   # crop taller than this (small) source frame: albumentations raises too; draw again
            continue
        if force_branch is None or (P.kernel is not None) == (force_branch == 0):
            break
    if force_h is not None:
        cw = int(force_h * (W / H))
        P.crop = ((Ws - cw) // 3, (Hs - force_h) // 2, cw, force_h)
    if force_line is not None:
        P.kernel = cv2_line_kernel(*force_line)
    x_u8, y_u8 = cv2_pipeline(img, label, P)
    gray = cv2.cvtColor(cv2.cvtColor(x_u8, cv2.COLOR_RGB2GRAY), cv2.COLOR_GRAY2RGB)
    np.savez_compressed(os.path.join(OUT, f"aug_{name}.npz"), img=img, label=label, lut_h=P.lut_h, lut_s=P.lut_s, lut_v=P.lut_v,
                        hsv=np.int32(P.hsv), crop=np.array(P.crop, np.int32),
                        kernel=P.kernel if P.kernel is not None else np.zeros((0, 0), np.float32),
                        gauss=P.gauss if P.gauss is not None else np.zeros((0,), np.float64),
                        out_u8=x_u8, out_label=y_u8, out_gray=gray, norm=pp.normalize(x_u8), cv2_version=cv2.__version__)
    # the restatement must reproduce cv2 bit for bit
    mine, mine_y = A.augment_u8(img, label, P, H, W)
    assert np.array_equal(mine, x_u8) and np.array_equal(mine_y, y_u8), name
    print(name, "crop", P.crop, "blur" if P.kernel is not None else "noise", "taps", int((P.kernel > 0).sum()) if P.kernel is not None else 0, "ok")
