"""Generates tests/golden/prep_*.npz with cv2 itself (the third-party library the reference's input pipeline calls,
myTransforms.py:8-31 through albumentations): run in the build container, where cv2 is installed.
    python oracle/make_golden_preprocess.py
Each fixture: uint8 BGR image, uint8 label, and cv2.resize INTER_LINEAR / INTER_NEAREST / RGB2GRAY outputs, plus the
float32 Normalize() result computed with albumentations' published formula on cv2's resized image."""
import os
import sys
import zlib

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import preprocess as pp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("480x640_to_120x160", 480, 640, 120, 160), ("97x131_to_120x160", 97, 131, 120, 160),
         ("240x320_to_120x160", 240, 320, 120, 160), ("120x160_identity", 120, 160, 120, 160),
         ("333x517_to_61x77", 333, 517, 61, 77)]
for name, Hs, Ws, H, W in CASES:
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    # smooth-ish synthetic road image plus noise so both flat and busy regions occur
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    base = (np.stack([xx * 255 // max(Ws - 1, 1), yy * 255 // max(Hs - 1, 1), (xx + yy) % 256], -1)).astype(np.int64)
    img = np.clip(base + rng.integers(-40, 41, (Hs, Ws, 3)), 0, 255).astype(np.uint8)
    label = (rng.integers(0, 4, (Hs // 8 + 1, Ws // 8 + 1), dtype=np.uint8)).repeat(8, 0).repeat(8, 1)[:Hs, :Ws].copy()
    lin = cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
    near = cv2.resize(label, (W, H), interpolation=cv2.INTER_NEAREST)
    gray = cv2.cvtColor(cv2.cvtColor(lin, cv2.COLOR_RGB2GRAY), cv2.COLOR_GRAY2RGB)
    np.savez_compressed(os.path.join(OUT, f"prep_{name}.npz"), img=img, label=label, H=H, W=W, linear=lin, nearest=near,
                        gray=gray, norm=pp.normalize(lin), cv2_version=cv2.__version__)
    print(name, "ok", cv2.__version__)
