"""GPU mirror of the reference's `dataManagement/myTransforms.py` (MyTransform, lines 5-31).

`MyTransform(width, height, gray, augment)(img, label)` takes what the reference's datasets hand it
(myDatasets.py:49-53): a uint8 BGR HxWx3 frame and a uint8 HxW label — numpy arrays or torch tensors, single or batched
([B,H,W,3] / [B,H,W]) — and returns `(float32 [3,h,w] or [B,3,h,w], int64 [h,w] or [B,h,w])` on the GPU.

augment=False: `rl_preprocess` (cv2-exact fixed-point bilinear resize, nearest resize of the label, optional ToGray,
ImageNet Normalize, HWC->CHW).

augment=True (myTransforms.py:8-13): HueSaturationValue -> RandomSizedCrop -> OneOf[MotionBlur, GaussNoise].  The random
parameters are drawn on the host exactly where albumentations 0.5.2 draws them — Python's `random` module (plus the
`numpy.random.RandomState` instances it seeds from it), one set per frame, same calls in the same order (`draw_params`)
— and the pixels are computed by `rl_preprocess_augment`, bit-exact with what albumentations makes cv2 compute for
those parameters.  `params=` injects pre-drawn parameters (a list of `AugmentParams`, one per frame) instead.
There is no CPU fallback: without CUDA the call raises."""
import ctypes
import random

import numpy as np
import torch

from .. import _lib


class _CParams(ctypes.Structure):
    """include/rightlane.h rl_augment_params"""
    _fields_ = [("lut_h", ctypes.c_ubyte * 256), ("lut_s", ctypes.c_ubyte * 256), ("lut_v", ctypes.c_ubyte * 256),
                ("hsv", ctypes.c_int), ("crop_x", ctypes.c_int), ("crop_y", ctypes.c_int), ("crop_w", ctypes.c_int),
                ("crop_h", ctypes.c_int), ("ksize", ctypes.c_int), ("kernel", ctypes.c_float * 49), ("noise", ctypes.c_int)]


PARAM_BYTES = ctypes.sizeof(_CParams)


class AugmentParams:
    """what one MyTransform(augment=True) call draws for one frame: the three cv2.LUT tables of HueSaturationValue
    (`hsv` False: all shifts were zero), the RandomSizedCrop window (x, y, w, h), the MotionBlur kernel (float32 [k,k]) or the
    GaussNoise plane (float64 [h,w,3])"""

    def __init__(self, lut_h, lut_s, lut_v, hsv, crop, kernel=None, gauss=None):
        self.lut_h, self.lut_s, self.lut_v, self.hsv = lut_h, lut_s, lut_v, bool(hsv)
        self.crop = tuple(int(c) for c in crop)
        self.kernel, self.gauss = kernel, gauss


def line_kernel(ksize, xs, ys, xe, ye):
    """MotionBlur.get_params: cv2.line((xs, ys), (xe, ye), thickness=1) on a ksize x ksize zero image, divided by its sum.
    cv2.line at thickness 1 is the 8-connected Bresenham walk of cv2.LineIterator started from the left endpoint."""
    k = np.zeros((ksize, ksize), dtype=np.float32)
    if xe < xs:
        xs, ys, xe, ye = xe, ye, xs, ys
    dx, ady = xe - xs, abs(ye - ys)
    sy = 1 if ye > ys else -1
    x, y = xs, ys
    major_x = dx >= ady
    long_, short_ = (dx, ady) if major_x else (ady, dx)
    err = long_ - 2 * short_
    for _ in range(long_ + 1):
        k[y, x] = 1.0
        step = err < 0
        err += -2 * short_ + (2 * long_ if step else 0)
        if major_x:
            x += 1
            y += sy if step else 0
        else:
            y += sy
            x += 1 if step else 0
    return (k / np.float32(k.sum())).astype(np.float32)


def draw_params(Hs, Ws, height, width, rng=random):
    """One frame's parameters with the `random` calls albumentations 0.5.2 makes for
    `Compose([Compose([HueSaturationValue, RandomSizedCrop, OneOf([MotionBlur, GaussNoise])]), ToGray|NoOp, Normalize,
    ToTensorV2])(image=, mask=)`, in its order: every Compose / transform / OneOf tests `random.random() < p` before
    anything else (even under always_apply), then draws its own parameters."""
    rng.random()                                   # outer Compose
    rng.random()                                   # augmentations Compose
    rng.random()                                   # HueSaturationValue
    hue, sat, val = rng.uniform(-20, 20), rng.uniform(-30, 30), rng.uniform(-20, 20)
    rng.random()                                   # RandomSizedCrop(min_max_height=(height // 2, height * 4))
    crop_h = rng.randint(height // 2, height * 4)
    h_start, w_start = rng.random(), rng.random()
    crop_w = int(crop_h * (width / height))
    if Hs < crop_h or Ws < crop_w:                 # albumentations' random_crop raises the same way
        raise ValueError(f"Requested crop size ({crop_h}, {crop_w}) is larger than the image size ({Hs}, {Ws})")
    y1, x1 = int((Hs - crop_h) * h_start), int((Ws - crop_w) * w_start)
    rng.random()                                   # OneOf(p=1)
    pick = np.random.RandomState(rng.randint(0, 2 ** 32 - 1)).choice(2, p=[0.5, 0.5])
    rng.random()                                   # the chosen transform
    kernel = gauss = None
    if pick == 0:                                  # MotionBlur(blur_limit=(3, 7))
        ksize = int(rng.choice(np.arange(3, 8, 2)))
        xs, xe = rng.randint(0, ksize - 1), rng.randint(0, ksize - 1)
        if xs == xe:
            ys, ye = rng.sample(range(ksize), 2)
        else:
            ys, ye = rng.randint(0, ksize - 1), rng.randint(0, ksize - 1)
        kernel = line_kernel(ksize, xs, ys, xe, ye)
    else:                                          # GaussNoise(var_limit=(10, 50), mean=0)
        sigma = rng.uniform(10.0, 50.0) ** 0.5
        gauss = np.random.RandomState(rng.randint(0, 2 ** 32 - 1)).normal(0, sigma, (height, width, 3))
    for _ in range(3):                             # ToGray | NoOp, Normalize, ToTensorV2
        rng.random()
    base = np.arange(0, 256, dtype=np.int16)       # functional._shift_hsv_uint8
    lut_h = np.mod(base + hue, 180).astype(np.uint8)
    lut_s = np.clip(base + sat, 0, 255).astype(np.uint8)
    lut_v = np.clip(base + val, 0, 255).astype(np.uint8)
    return AugmentParams(lut_h, lut_s, lut_v, not (hue == 0 and sat == 0 and val == 0), (x1, y1, crop_w, crop_h), kernel, gauss)


def pack_params(plist, height, width):
    """AugmentParams list -> (uint8 [B, PARAM_BYTES] host tensor, float64 [B,h,w,3] host tensor or None)"""
    arr = (_CParams * len(plist))()
    noise = None
    for b, P in enumerate(plist):
        c = arr[b]
        ctypes.memmove(c.lut_h, np.ascontiguousarray(P.lut_h, dtype=np.uint8).ctypes.data, 256)
        ctypes.memmove(c.lut_s, np.ascontiguousarray(P.lut_s, dtype=np.uint8).ctypes.data, 256)
        ctypes.memmove(c.lut_v, np.ascontiguousarray(P.lut_v, dtype=np.uint8).ctypes.data, 256)
        c.hsv = int(P.hsv)
        c.crop_x, c.crop_y, c.crop_w, c.crop_h = P.crop
        c.ksize = 0
        c.noise = 0
        if P.kernel is not None and P.kernel.size:
            k = np.ascontiguousarray(P.kernel, dtype=np.float32)
            if k.ndim != 2 or k.shape[0] != k.shape[1] or k.shape[0] not in (3, 5, 7):
                raise ValueError("MyTransform: the MotionBlur kernel must be 3x3, 5x5 or 7x7")
            c.ksize = k.shape[0]
            ctypes.memmove(c.kernel, k.ctypes.data, k.size * 4)
        if P.gauss is not None and np.size(P.gauss):
            if noise is None:
                noise = np.zeros((len(plist), height, width, 3), dtype=np.float64)
            noise[b] = np.asarray(P.gauss, dtype=np.float64).reshape(height, width, 3)
            c.noise = 1
    raw = np.frombuffer(arr, dtype=np.uint8).reshape(len(plist), PARAM_BYTES).copy()
    return torch.from_numpy(raw), (torch.from_numpy(noise) if noise is not None else None)


class MyTransform:
    def __init__(self, width=160, height=120, gray=False, augment=False):
        self.width, self.height, self.gray, self.augment = int(width), int(height), bool(gray), bool(augment)

    def __call__(self, img, label=None, params=None):
        if not torch.cuda.is_available():
            raise RuntimeError("MyTransform: the CUDA path is required (no CPU fallback)")
        L = _lib.lib()
        img_t = torch.as_tensor(np.ascontiguousarray(img) if isinstance(img, np.ndarray) else img)
        single = img_t.dim() == 3
        if single:
            img_t = img_t[None]
        if img_t.dtype != torch.uint8 or img_t.dim() != 4 or img_t.shape[-1] != 3:
            raise ValueError("MyTransform: expected uint8 [H,W,3] or [B,H,W,3] BGR frames")
        B, Hs, Ws, _ = img_t.shape
        plist = None
        if self.augment:   # draw before any device work, as the reference does (and so that a bad crop raises first)
            plist = params if params is not None else [draw_params(Hs, Ws, self.height, self.width) for _ in range(B)]
            if isinstance(plist, AugmentParams):
                plist = [plist]
            if len(plist) != B:
                raise ValueError("MyTransform: one AugmentParams per frame")
            for P in plist:
                x1, y1, cw, ch = P.crop
                if cw < 1 or ch < 1 or x1 < 0 or y1 < 0 or x1 + cw > Ws or y1 + ch > Hs:
                    raise ValueError(f"MyTransform: crop window {P.crop} outside the {Hs}x{Ws} frame")
        img_t = img_t.cuda(non_blocking=True).contiguous()
        have_label = label is not None and getattr(label, "ndim", 0) >= 2
        lab_t = None
        if have_label:
            lab_t = torch.as_tensor(np.ascontiguousarray(label) if isinstance(label, np.ndarray) else label)
            lab_t = lab_t.reshape(B, Hs, Ws).to(torch.uint8).cuda(non_blocking=True).contiguous()
        x = torch.empty(B, 3, self.height, self.width, dtype=torch.float32, device="cuda")
        y = torch.empty(B, self.height, self.width, dtype=torch.int64, device="cuda") if have_label else None
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        lab_p = ctypes.c_void_p(lab_t.data_ptr() if have_label else 0)
        y_p = ctypes.c_void_p(y.data_ptr() if have_label else 0)
        if self.augment:
            raw, noise = pack_params(plist, self.height, self.width)
            raw = raw.cuda(non_blocking=True)
            noise = noise.cuda(non_blocking=True) if noise is not None else None
            scratch = torch.empty(B, self.height, self.width, 3, dtype=torch.uint8, device="cuda")
            rc = L.rl_preprocess_augment(ctypes.c_void_p(img_t.data_ptr()), lab_p, B, Hs, Ws, self.height, self.width,
                                         int(self.gray), ctypes.c_void_p(raw.data_ptr()),
                                         ctypes.c_void_p(noise.data_ptr() if noise is not None else 0),
                                         ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(x.data_ptr()), y_p, stream)
        else:
            rc = L.rl_preprocess(ctypes.c_void_p(img_t.data_ptr()), lab_p, B, Hs, Ws, self.height, self.width, int(self.gray),
                                 ctypes.c_void_p(x.data_ptr()), y_p, stream)
        if rc != 0:
            L.rl_last_error.restype = ctypes.c_char_p
            raise RuntimeError(L.rl_last_error().decode())
        if single:
            x = x[0]
            y = y[0] if y is not None else None
        if not have_label:
            y = label   # the reference returns the label argument unchanged in this branch (myTransforms.py:28-31)
        return x, y
