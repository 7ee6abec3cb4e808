"""TEST INFRASTRUCTURE — CPU restatement of the reference's deterministic input pipeline (no augmentation).

Follows rightLaneNetwork/dataManagement/myTransforms.py:8-31 (MyTransform with augment=False):
  albumentations.Resize(height, width)      -> cv2.resize(img, (w, h), INTER_LINEAR) for the image,
                                               cv2.resize(mask, (w, h), INTER_NEAREST) for the label
  albumentations.ToGray (gray=True)         -> cv2.cvtColor(RGB2GRAY) then GRAY2RGB, applied to the BGR image as it is
  albumentations.Normalize()                -> (img / 255 - mean) / std, ImageNet mean/std in channel order 0,1,2
  ToTensorV2                                -> HWC -> CHW; label .squeeze().long()   (myTransforms.py:26-27)
and myDatasets.py:49-53 (cv2.imread IMREAD_COLOR = uint8 BGR HWC, IMREAD_GRAYSCALE label).

The arithmetic lives in third-party code that is not vendored in the reference: opencv-python (requirements.txt:
`opencv-python~=4.5`) and albumentations (`albumentations~=0.5.2`).  cv2's uint8 INTER_LINEAR is a fixed-point
algorithm (11-bit coefficients, imgproc/src/resize.cpp: HResizeLinear / VResizeLinear<uchar>); it is restated here in
numpy integer arithmetic and pinned bit-exactly against cv2 itself by oracle/make_golden_preprocess.py (cv2 4.13 is
installed in the build container; albumentations is not, its three calls are restated from its published source).
"""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _cv_round(v):
    """cvRound: round half to even (lrint)"""
    return np.rint(v).astype(np.int64)


def linear_coeffs(dst, src):
    """cv2 resize.cpp (INTER_LINEAR, uint8): per destination index the left source index and the two 11-bit weights.
    x direction semantics: an index that falls off the right edge gets weight (2048, 0) on the last pixel."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(img, H, W):
    """bit-exact restatement of cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC input"""
    Hs, Ws = img.shape[:2]
    if (Hs, Ws) == (H, W):
        return img.copy()
    src = img.reshape(Hs, Ws, -1).astype(np.int64)
    sx, fx = linear_coeffs(W, Ws)
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx)
    sx = np.where(lo, 0, sx)
    hi = sx >= Ws - 1
    fx = np.where(hi, np.float32(0), fx)
    sx = np.where(hi, Ws - 1, sx)
    a0 = np.clip(_cv_round((np.float32(1) - fx) * np.float32(COEF_SCALE)), -32768, 32767)
    a1 = np.clip(_cv_round(fx * np.float32(COEF_SCALE)), -32768, 32767)
    sx1 = np.minimum(sx + 1, Ws - 1)
    sy, fy = linear_coeffs(H, Hs)
    b0 = np.clip(_cv_round((np.float32(1) - fy) * np.float32(COEF_SCALE)), -32768, 32767)
    b1 = np.clip(_cv_round(fy * np.float32(COEF_SCALE)), -32768, 32767)
    r0 = np.clip(sy, 0, Hs - 1)
    r1 = np.clip(sy + 1, 0, Hs - 1)
    # horizontal pass (int32 rows scaled by 2048)
    hrow = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]  # [Hs][W][C]
    S0, S1 = hrow[r0], hrow[r1]                                                    # [H][W][C]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((H, W) + img.shape[2:])


def resize_nearest_u8(img, H, W):
    """cv2.resize(..., INTER_NEAREST): sx = min(floor(dx * Ws / W), Ws - 1)"""
    Hs, Ws = img.shape[:2]
    if (Hs, Ws) == (H, W):
        return img.copy()
    ix = np.minimum(np.floor(np.arange(W) * (np.float64(Ws) / W)).astype(np.int64), Ws - 1)
    iy = np.minimum(np.floor(np.arange(H) * (np.float64(Hs) / H)).astype(np.int64), Hs - 1)
    return img[iy][:, ix]


def to_gray_u8(img):
    """cv2.cvtColor(img, COLOR_RGB2GRAY) fixed point then GRAY2RGB.  cv2 4.13 (the version the pin was taken with) uses
    15-bit weights: ch0 * 9798 + ch1 * 19235 + ch2 * 3735, rounded; albumentations hands it the BGR image unchanged."""
    i = img.astype(np.int64)
    g = (i[..., 0] * 9798 + i[..., 1] * 19235 + i[..., 2] * 3735 + (1 << 14)) >> 15
    return np.repeat(g.astype(np.uint8)[..., None], 3, axis=2)


def normalize(img_u8):
    """albumentations.augmentations.functional.normalize (0.5.2): img.astype(float32); img -= mean*255; img *= 1/(std*255)"""
    mean = MEAN * np.float32(255.0)
    denom = np.reciprocal(STD * np.float32(255.0), dtype=np.float32)
    x = img_u8.astype(np.float32)
    x -= mean
    x *= denom
    return x


def my_transform(img_bgr, label=None, width=160, height=120, gray=False):
    """MyTransform(width, height, gray, augment=False).__call__: returns (float32 CHW, int64 HW or None)"""
    x = resize_linear_u8(img_bgr, height, width)
    if gray:
        x = to_gray_u8(x)
    x = normalize(x).transpose(2, 0, 1).copy()
    y = None
    if label is not None and label.ndim >= 2:
        y = resize_nearest_u8(label, height, width).squeeze().astype(np.int64)
    return x, y


def overlay(frame_bgr, pred, width=160, height=120):
    """makeDemoVideo.py:38-42: `cv2.resize(frame, framesize, cv2.INTER_LANCZOS4)` passes the flag in the *dst* position, so
    the resize runs with the default INTER_LINEAR; then the three non-background classes are painted (BGR)."""
    out = resize_linear_u8(frame_bgr, height, width)
    out[pred == 1] = (0, 255, 0)
    out[pred == 2] = (255, 0, 0)
    out[pred == 3] = (0, 0, 255)
    return out
