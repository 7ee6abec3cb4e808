"""GPU mirror of the reference's `dataManagement/myTransforms.py` (MyTransform, lines 5-31), deterministic part.

`MyTransform(width, height, gray, augment=False)(img, label)` takes what the reference's datasets hand it
(myDatasets.py:49-53): a uint8 BGR HxWx3 frame and a uint8 HxW label — numpy arrays or torch tensors, single or batched
([B,H,W,3] / [B,H,W]) — and returns `(float32 [3,h,w] or [B,3,h,w], int64 [h,w] or [B,h,w])` on the GPU, computed by
`rl_preprocess` (cv2-exact fixed-point bilinear resize, nearest resize of the label, optional ToGray, ImageNet
Normalize, HWC->CHW).  `augment=True` (HueSaturationValue / RandomSizedCrop / MotionBlur-or-GaussNoise, all driven by
albumentations' host RNG) is not implemented on the GPU: it raises NotImplementedError, keep the reference's CPU
transform for that flag.  There is no CPU fallback: without CUDA the call raises."""
import ctypes

import numpy as np
import torch

from .. import _lib


class MyTransform:
    def __init__(self, width=160, height=120, gray=False, augment=False):
        if augment:
            raise NotImplementedError("MyTransform(augment=True): the random augmentations are not part of the CUDA path")
        self.width, self.height, self.gray = int(width), int(height), bool(gray)

    def __call__(self, img, label=None):
        if not torch.cuda.is_available():
            raise RuntimeError("MyTransform: the CUDA path is required (no CPU fallback)")
        L = _lib.lib()
        img_t = torch.as_tensor(np.ascontiguousarray(img) if isinstance(img, np.ndarray) else img)
        single = img_t.dim() == 3
        if single:
            img_t = img_t[None]
        if img_t.dtype != torch.uint8 or img_t.dim() != 4 or img_t.shape[-1] != 3:
            raise ValueError("MyTransform: expected uint8 [H,W,3] or [B,H,W,3] BGR frames")
        img_t = img_t.cuda(non_blocking=True).contiguous()
        B, Hs, Ws, _ = img_t.shape
        have_label = label is not None and getattr(label, "ndim", 0) >= 2
        lab_t = None
        if have_label:
            lab_t = torch.as_tensor(np.ascontiguousarray(label) if isinstance(label, np.ndarray) else label)
            lab_t = lab_t.reshape(B, Hs, Ws).to(torch.uint8).cuda(non_blocking=True).contiguous()
        x = torch.empty(B, 3, self.height, self.width, dtype=torch.float32, device="cuda")
        y = torch.empty(B, self.height, self.width, dtype=torch.int64, device="cuda") if have_label else None
        rc = L.rl_preprocess(ctypes.c_void_p(img_t.data_ptr()), ctypes.c_void_p(lab_t.data_ptr() if have_label else 0),
                             B, Hs, Ws, self.height, self.width, int(self.gray), ctypes.c_void_p(x.data_ptr()),
                             ctypes.c_void_p(y.data_ptr() if have_label else 0),
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            L.rl_last_error.restype = ctypes.c_char_p
            raise RuntimeError(L.rl_last_error().decode())
        if single:
            x = x[0]
            y = y[0] if y is not None else None
        if not have_label:
            y = label   # the reference returns the label argument unchanged in this branch (myTransforms.py:28-31)
        return x, y
