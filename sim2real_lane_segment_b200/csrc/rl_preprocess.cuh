// rl_preprocess.cuh — cv2.resize(INTER_LINEAR) on uint8 as device functions shared by rl_preprocess.cu and rl_augment.cu:
// cv2's 11-bit fixed-point algorithm (imgproc/src/resize.cpp HResizeLinear / VResizeLinear<uchar>), bit-exact.
#pragma once
#include "rl_common.cuh"

namespace rl {

__device__ __forceinline__ int cv_round_to_short(float v) {
  int r = __float2int_rn(v);  // cvRound: round half to even
  return max(-32768, min(32767, r));
}

struct LinearTaps {
  int sx, sx1, r0, r1;  // source columns / rows of the four taps
  int a0, a1, b0, b1;   // 11-bit horizontal / vertical weights
  bool copy;            // same size: cv2 copies
};

// taps of destination pixel (dx, dy) for a Hs x Ws -> H x W resize
__device__ __forceinline__ LinearTaps linear_taps(int Hs, int Ws, int H, int W, double scale_x, double scale_y, int dx, int dy) {
  LinearTaps t;
  t.copy = (Hs == H && Ws == W);
  if (t.copy) {
    t.sx = t.sx1 = dx;
    t.r0 = t.r1 = dy;
    t.a0 = t.b0 = 2048;
    t.a1 = t.b1 = 0;
    return t;
  }
  float fx = (float)(((double)dx + 0.5) * scale_x - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) {
    fx = 0.f;
    sx = 0;
  }
  if (sx >= Ws - 1) {
    fx = 0.f;
    sx = Ws - 1;
  }
  t.a0 = cv_round_to_short((1.f - fx) * 2048.f);
  t.a1 = cv_round_to_short(fx * 2048.f);
  t.sx = sx;
  t.sx1 = min(sx + 1, Ws - 1);
  float fy = (float)(((double)dy + 0.5) * scale_y - 0.5);
  const int sy = (int)floorf(fy);
  fy -= (float)sy;
  t.b0 = cv_round_to_short((1.f - fy) * 2048.f);
  t.b1 = cv_round_to_short(fy * 2048.f);
  t.r0 = max(0, min(sy, Hs - 1));
  t.r1 = max(0, min(sy + 1, Hs - 1));
  return t;
}

// one channel: p00 p01 / p10 p11 = (row r0 | r1, column sx | sx1)
__device__ __forceinline__ int linear_blend(const LinearTaps& t, int p00, int p01, int p10, int p11) {
  if (t.copy) return p00;
  const int S0 = p00 * t.a0 + p01 * t.a1, S1 = p10 * t.a0 + p11 * t.a1;
  const int o = (((t.b0 * (S0 >> 4)) >> 16) + ((t.b1 * (S1 >> 4)) >> 16) + 2) >> 2;
  return max(0, min(255, o));
}

// cv2.resize(src, (W, H)) with INTER_LINEAR (the default), uint8, 3 channels: one output pixel
__device__ __forceinline__ void resize_linear_px(const uint8_t* src, int Hs, int Ws, int H, int W, double scale_x, double scale_y,
                                                 int dx, int dy, int* v) {
  const LinearTaps t = linear_taps(Hs, Ws, H, W, scale_x, scale_y, dx, dy);
  const uint8_t* p00 = src + ((size_t)t.r0 * Ws + t.sx) * 3;
  const uint8_t* p01 = src + ((size_t)t.r0 * Ws + t.sx1) * 3;
  const uint8_t* p10 = src + ((size_t)t.r1 * Ws + t.sx) * 3;
  const uint8_t* p11 = src + ((size_t)t.r1 * Ws + t.sx1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = linear_blend(t, p00[c], p01[c], p10[c], p11[c]);
}

}  // namespace rl
