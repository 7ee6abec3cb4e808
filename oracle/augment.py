"""TEST INFRASTRUCTURE — CPU restatement of the reference's random augmentation pipeline (MyTransform(augment=True)).

Follows rightLaneNetwork/dataManagement/myTransforms.py:8-20:
    Compose([HueSaturationValue(always_apply=True),
             RandomSizedCrop(min_max_height=(height // 2, height * 4), height, width, w2h_ratio=width / height, always_apply=True),
             OneOf([MotionBlur(p=0.5), GaussNoise(p=0.5)], p=1)])
followed by ToGray / NoOp, Normalize, ToTensorV2 (oracle/preprocess.py).

The arithmetic lives in two third-party packages that are not vendored in the reference: albumentations
(requirements.txt: `albumentations~=0.5.2`, NOT installed in the build container) and opencv-python (`~=4.5`; cv2 4.13 is
installed).  What is restated, and how each part is pinned:
  * the PIXEL ARITHMETIC for given parameters — cv2.cvtColor RGB2HSV / HSV2RGB on uint8, cv2.LUT, the crop +
    cv2.resize(INTER_LINEAR / INTER_NEAREST), cv2.filter2D with a float32 kernel (BORDER_REFLECT_101), and
    albumentations' gauss_noise (float32 image + float64 noise, clip, truncate) — is restated in numpy and pinned
    bit-exactly against cv2 4.13 itself by oracle/make_golden_augment.py (fixtures tests/golden/aug_*.npz) and, for the
    colour conversions, exhaustively over all 2^24 colours / all (h < 180, s, v) triples by tests/test_augment_cpu.py when
    cv2 is importable;
  * the PARAMETER SAMPLING (which `random` calls albumentations 0.5.2 makes, in which order) is restated from the
    package's published source (albumentations/core/composition.py, core/transforms_interface.py,
    augmentations/transforms.py at tag 0.5.2).  Stream equality with the real package is UNPINNED (the package is absent);
    the distributions are those of the source.  The CUDA path takes the drawn parameters as an input, so its parity
    does not depend on this part.

cv2 specifics found while pinning (cv2 4.13, x86-64 vector paths):
  * RGB2HSV uint8 is integer arithmetic with 12-bit reciprocal tables (color_hsv: RGB2HSV_b);
  * HSV2RGB uint8 goes through float32: s, v scaled by 1/255, tab1 = v*(1-s), tab2 = v*fma(-s, h, 1),
    tab3 = v*fma(-s, 1-h, 1), result *255 and TRUNCATED (not rounded);
  * filter2D (8u, float32 kernel, ksize < 11) accumulates acc = fma(src, k, acc) in float32 over the kernel's non-zero
    taps in row-major order, then rounds half to even.
albumentations hands cv2 the frame exactly as cv2.imread produced it (BGR) but asks for COLOR_RGB2HSV: the restatement
keeps that (channel 0 is treated as "R")."""
import random as _random

import numpy as np

from . import preprocess as pp

F32 = np.float32
HSCALE = F32(6.0 / 180.0)
INV255 = F32(1.0 / 255.0)
_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])  # (b, g, r) table index per sector


def _div_tables():
    idx = np.arange(256, dtype=np.float64)
    with np.errstate(divide="ignore"):
        sdiv = np.rint((255 << 12) / idx)
        hdiv = np.rint((180 << 12) / (6.0 * idx))
    sdiv[0] = 0
    hdiv[0] = 0
    return sdiv.astype(np.int64), hdiv.astype(np.int64)


_SDIV, _HDIV = _div_tables()


def rgb2hsv_u8(img):
    """cv2.cvtColor(img, COLOR_RGB2HSV) for uint8 (H in [0,180)): RGB2HSV_b, hsv_shift = 12"""
    i = img.astype(np.int64)
    r, g, b = i[..., 0], i[..., 1], i[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = (diff * _SDIV[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def _fma32(a, b, c):
    """fl32(a*b + c) with one rounding (products of two float32 are exact in float64; the float64 sum is rounded once more,
    harmless here because |c| = 1 and |a*b| <= 1 keep the float64 sum exact to 2^-52)"""
    return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(F32)


def hsv2rgb_u8(hsv):
    """cv2.cvtColor(hsv, COLOR_HSV2RGB) for uint8 with h < 180 (vector path of HSV2RGB_b, see the module docstring)"""
    h = hsv[..., 0].astype(F32) * HSCALE
    sec = np.floor(h).astype(np.int64)
    hf = h - sec.astype(F32)
    bad = (sec < 0) | (sec >= 6)
    sec = np.where(bad, 0, sec)
    hf = np.where(bad, F32(0), hf)
    s = hsv[..., 1].astype(F32) * INV255
    v = hsv[..., 2].astype(F32) * INV255
    one = np.ones_like(s)
    tab = np.stack([v, v * (one - s), v * _fma32(-s, hf, one), v * _fma32(-s, one - hf, one)], -1)
    idx = _SECTOR[sec]
    out = np.stack([np.take_along_axis(tab, idx[..., k:k + 1], -1)[..., 0] for k in (2, 1, 0)], -1)
    out = np.where((hsv[..., 1] == 0)[..., None], v[..., None], out)
    return np.clip(np.trunc(out * F32(255.0)), 0, 255).astype(np.uint8)


def hsv_luts(hue_shift, sat_shift, val_shift):
    """albumentations 0.5.2 functional._shift_hsv_uint8: the three 256-entry tables handed to cv2.LUT"""
    base = np.arange(0, 256, dtype=np.int16)
    lut_h = np.mod(base + hue_shift, 180).astype(np.uint8)
    lut_s = np.clip(base + sat_shift, 0, 255).astype(np.uint8)
    lut_v = np.clip(base + val_shift, 0, 255).astype(np.uint8)
    return lut_h, lut_s, lut_v


def shift_hsv(img, lut_h, lut_s, lut_v):
    """functional.shift_hsv on a uint8 colour image: RGB2HSV, LUT per plane, HSV2RGB"""
    hsv = rgb2hsv_u8(img)
    hsv = np.stack([lut_h[hsv[..., 0]], lut_s[hsv[..., 1]], lut_v[hsv[..., 2]]], -1)
    return hsv2rgb_u8(hsv)


def crop_coords(Hs, Ws, crop_h, crop_w, h_start, w_start):
    """functional.get_random_crop_coords / random_crop"""
    if Hs < crop_h or Ws < crop_w:
        raise ValueError(f"Requested crop size ({crop_h}, {crop_w}) is larger than the image size ({Hs}, {Ws})")
    y1 = int((Hs - crop_h) * h_start)
    x1 = int((Ws - crop_w) * w_start)
    return x1, y1


def filter2d_u8(img, kernel):
    """cv2.filter2D(img, ddepth=-1, kernel) for uint8 HxWxC and a small float32 kernel (anchor centre, BORDER_REFLECT_101)"""
    k = np.asarray(kernel, dtype=F32)
    ks = k.shape[0]
    r = ks // 2
    p = np.pad(img, ((r, r), (r, r), (0, 0)), mode="reflect")
    H, W = img.shape[:2]
    acc = np.zeros(img.shape, np.float64)
    for ky in range(ks):
        for kx in range(ks):
            if k[ky, kx] != 0:
                acc = (p[ky:ky + H, kx:kx + W].astype(np.float64) * np.float64(k[ky, kx]) + acc).astype(F32).astype(np.float64)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def gauss_noise_u8(img, gauss):
    """functional.gauss_noise under @clipped: float32 image + float64 noise, clip to [0, 255], astype(uint8) (truncation)"""
    return np.clip(img.astype(np.float32) + np.asarray(gauss, dtype=np.float64), 0, 255).astype(np.uint8)


def motion_blur_kernel(ksize, xs, ys, xe, ye):
    """MotionBlur.get_params: cv2.line((xs, ys) -> (xe, ye), thickness 1) on a ksize x ksize zero image, normalised.
    cv2.line with thickness 1 and LINE_8 is LineIterator's 8-connected Bresenham walk from the left endpoint (imgproc/src/drawing.cpp); pinned against cv2.line for every endpoint
    pair of the three kernel sizes by tests/test_augment_cpu.py."""
    k = np.zeros((ksize, ksize), dtype=np.uint8)
    if xe < xs:                                    # LineIterator(..., leftToRight=true): always start from the left end
        xs, ys, xe, ye = xe, ye, xs, ys
    dx, dy = xe - xs, ye - ys
    sy = 1 if dy > 0 else -1
    ady = abs(dy)
    x, y = xs, ys
    if dx >= ady:                                  # x is the major axis
        err = dx - 2 * ady
        for _ in range(dx + 1):
            k[y, x] = 1
            minor = err < 0
            err += -2 * ady + (2 * dx if minor else 0)
            x += 1
            y += sy if minor else 0
    else:
        err = ady - 2 * dx
        for _ in range(ady + 1):
            k[y, x] = 1
            minor = err < 0
            err += -2 * dx + (2 * ady if minor else 0)
            y += sy
            x += 1 if minor else 0
    return (k.astype(F32) / F32(k.sum())).astype(F32)


class AugmentParams:
    """the parameters one MyTransform(augment=True) call draws for one frame"""

    def __init__(self, lut_h, lut_s, lut_v, hsv, crop, kernel=None, gauss=None):
        self.lut_h, self.lut_s, self.lut_v, self.hsv = lut_h, lut_s, lut_v, bool(hsv)
        self.crop = tuple(int(c) for c in crop)   # x, y, w, h
        self.kernel = kernel                      # float32 [k,k] or None
        self.gauss = gauss                        # float64 [H,W,3] or None


def draw_params(Hs, Ws, height=120, width=160, rng=_random, line=None):
    """the `random` calls of albumentations 0.5.2 for one self.transform(image=, mask=) call of the augmenting pipeline,
    in the package's order (core/composition.py Compose.__call__ / OneOf.__call__, core/transforms_interface.py
    BasicTransform.__call__, augmentations/transforms.py get_params of the four transforms)."""
    rng.random()                                   # outer Compose: random.random() < p
    rng.random()                                   # inner Compose
    rng.random()                                   # HueSaturationValue.__call__ (p test is evaluated before always_apply)
    hue = rng.uniform(-20, 20)
    sat = rng.uniform(-30, 30)
    val = rng.uniform(-20, 20)
    rng.random()                                   # RandomSizedCrop.__call__
    crop_h = rng.randint(height // 2, height * 4)
    h_start, w_start = rng.random(), rng.random()
    crop_w = int(crop_h * (width / height))
    x1, y1 = crop_coords(Hs, Ws, crop_h, crop_w, h_start, w_start)
    rng.random()                                   # OneOf: random.random() < p
    choice = np.random.RandomState(rng.randint(0, 2 ** 32 - 1)).choice(2, p=[0.5, 0.5])
    rng.random()                                   # the chosen transform's own __call__
    kernel = gauss = None
    if choice == 0:                                # MotionBlur(blur_limit=(3, 7))
        ksize = int(rng.choice(np.arange(3, 8, 2)))
        xs, xe = rng.randint(0, ksize - 1), rng.randint(0, ksize - 1)
        if xs == xe:
            ys, ye = rng.sample(range(ksize), 2)
        else:
            ys, ye = rng.randint(0, ksize - 1), rng.randint(0, ksize - 1)
        kernel = (line or motion_blur_kernel)(ksize, xs, ys, xe, ye)
    else:                                          # GaussNoise(var_limit=(10, 50), mean=0)
        var = rng.uniform(10.0, 50.0)
        sigma = var ** 0.5
        gauss = np.random.RandomState(rng.randint(0, 2 ** 32 - 1)).normal(0, sigma, (height, width, 3))
    for _ in range(3):                             # ToGray / NoOp, Normalize, ToTensorV2
        rng.random()
    shifted = not (hue == 0 and sat == 0 and val == 0)
    lut_h, lut_s, lut_v = hsv_luts(hue, sat, val)
    return AugmentParams(lut_h, lut_s, lut_v, shifted, (x1, y1, crop_w, crop_h), kernel, gauss)


def augment_u8(img_bgr, label, P, height=120, width=160):
    """the uint8 stage: HueSaturationValue -> RandomSizedCrop -> MotionBlur | GaussNoise.  Returns (uint8 [h,w,3], label or None)"""
    x = shift_hsv(img_bgr, P.lut_h, P.lut_s, P.lut_v) if P.hsv else img_bgr
    cx, cy, cw, ch = P.crop
    x = pp.resize_linear_u8(np.ascontiguousarray(x[cy:cy + ch, cx:cx + cw]), height, width)
    if P.kernel is not None:
        x = filter2d_u8(x, P.kernel)
    if P.gauss is not None:
        x = gauss_noise_u8(x, P.gauss)
    y = None
    if label is not None:
        y = pp.resize_nearest_u8(np.ascontiguousarray(label[cy:cy + ch, cx:cx + cw]), height, width)
    return x, y


def my_transform_augment(img_bgr, label, P, width=160, height=120, gray=False):
    """MyTransform(width, height, gray, augment=True).__call__ for drawn parameters P: (float32 CHW, int64 HW or None)"""
    x, y = augment_u8(img_bgr, label if (label is not None and label.ndim >= 2) else None, P, height, width)
    if gray:
        x = pp.to_gray_u8(x)
    x = pp.normalize(x).transpose(2, 0, 1).copy()
    return x, (y.squeeze().astype(np.int64) if y is not None else None)
