// rl_preprocess.cu — the deterministic part of the reference's input pipeline on the GPU, one fused kernel:
//   uint8 BGR HWC frames (cv2.imread, myDatasets.py:49-53) -> Resize (cv2 INTER_LINEAR fixed point for the image,
//   INTER_NEAREST for the label) -> optional ToGray -> Normalize(ImageNet mean/std) -> CHW float32 + int64 label
//   (MyTransform with augment=False, myTransforms.py:8-31).
// The uint8 bilinear resize follows cv2's 11-bit fixed-point algorithm exactly (resize.cpp HResizeLinear /
// VResizeLinear<uchar>), so the resized bytes are identical to the reference's and the float result differs only by
// how the two fp32 operations of Normalize are fused.  One thread per output pixel; the 4 source pixels of the three
// channels are adjacent bytes, the output is written as three coalesced fp32 planes.
#include "rl_common.cuh"
#include "rl_preprocess.cuh"

namespace rl {

struct PrepArgs {
  const uint8_t* img;    // [B][Hs][Ws][3]
  const uint8_t* label;  // [B][Hs][Ws] or null
  float* x;              // [B][3][H][W]
  long long* y;          // [B][H][W] or null
  int B, Hs, Ws, H, W, gray;
  double scale_x, scale_y;
  float m[3], d[3];      // mean*255, 1/(std*255)
};

__global__ void __launch_bounds__(256) preprocess_kernel(PrepArgs a) {
  const long n = (long)a.B * a.H * a.W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % a.W);
    const int dy = (int)((i / a.W) % a.H);
    const int b = (int)(i / ((long)a.W * a.H));
    const uint8_t* src = a.img + (size_t)b * a.Hs * a.Ws * 3;
    int v[3];
    resize_linear_px(src, a.Hs, a.Ws, a.H, a.W, a.scale_x, a.scale_y, dx, dy, v);
    if (a.gray) {  // cv2 COLOR_RGB2GRAY (15-bit weights) on the channels as they are, then GRAY2RGB
      const int g = (v[0] * 9798 + v[1] * 19235 + v[2] * 3735 + (1 << 14)) >> 15;
      v[0] = v[1] = v[2] = g;
    }
    const size_t plane = (size_t)a.H * a.W, o = (size_t)b * 3 * plane + (size_t)dy * a.W + dx;
#pragma unroll
    for (int c = 0; c < 3; ++c) a.x[o + c * plane] = __fmul_rn(__fsub_rn((float)v[c], a.m[c]), a.d[c]);
    if (a.y != nullptr) {
      int lx = dx, ly = dy;
      if (!(a.Hs == a.H && a.Ws == a.W)) {  // INTER_NEAREST: floor(d * scale), clamped
        lx = min((int)floor((double)dx * a.scale_x), a.Ws - 1);
        ly = min((int)floor((double)dy * a.scale_y), a.Hs - 1);
      }
      a.y[i] = (long long)a.label[((size_t)b * a.Hs + ly) * a.Ws + lx];
    }
  }
}

// makeDemoVideo.py:38-42 / comparison.py: the frame resized to the network resolution with the predicted classes painted
// over it (BGR): 1 right lane (0,255,0), 2 left lane (255,0,0), 3 obstacles (0,0,255); class 0 keeps the frame.
__global__ void __launch_bounds__(256) overlay_kernel(const uint8_t* frames, int B, int Hs, int Ws, int H, int W, double scale_x,
                                                      double scale_y, const uint8_t* mask8, const long long* mask64, uint8_t* out) {
  const long n = (long)B * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % W), dy = (int)((i / W) % H), b = (int)(i / ((long)W * H));
    int v[3];
    resize_linear_px(frames + (size_t)b * Hs * Ws * 3, Hs, Ws, H, W, scale_x, scale_y, dx, dy, v);
    const int cls = mask8 ? (int)mask8[i] : (int)mask64[i];
    if (cls == 1) {
      v[0] = 0; v[1] = 255; v[2] = 0;
    } else if (cls == 2) {
      v[0] = 255; v[1] = 0; v[2] = 0;
    } else if (cls == 3) {
      v[0] = 0; v[1] = 0; v[2] = 255;
    }
    out[i * 3 + 0] = (uint8_t)v[0];
    out[i * 3 + 1] = (uint8_t)v[1];
    out[i * 3 + 2] = (uint8_t)v[2];
  }
}

}  // namespace rl

extern "C" int rl_overlay(const unsigned char* frames_bgr, int B, int Hs, int Ws, int H, int W, const void* mask, int mask_is_i64,
                          unsigned char* out, void* stream) {
  using namespace rl;
  if (!frames_bgr || !mask || !out) return rl_fail(RL_ERR_ARG, "rl_overlay: null argument");
  if (B < 1 || Hs < 1 || Ws < 1 || H < 1 || W < 1) return rl_fail(RL_ERR_SHAPE, "rl_overlay: bad shape");
  const long n = (long)B * H * W;
  const int grid = (int)std::min<long>((n + 255) / 256, 148L * 16);
  { ProfScope _ps(KC_MISC, (cudaStream_t)stream);
  overlay_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames_bgr, B, Hs, Ws, H, W, (double)Ws / W, (double)Hs / H,
                                                         mask_is_i64 ? nullptr : (const uint8_t*)mask,
                                                         mask_is_i64 ? (const long long*)mask : nullptr, out); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

extern "C" int rl_preprocess(const unsigned char* img_bgr, const unsigned char* label, int B, int Hs, int Ws, int H, int W,
                             int gray, float* x_out, long long* y_out, void* stream) {
  using namespace rl;
  if (!img_bgr || !x_out || (y_out && !label)) return rl_fail(RL_ERR_ARG, "rl_preprocess: null argument");
  if (B < 1 || Hs < 1 || Ws < 1 || H < 1 || W < 1) return rl_fail(RL_ERR_SHAPE, "rl_preprocess: bad shape");
  PrepArgs a;
  a.img = img_bgr;
  a.label = label;
  a.x = x_out;
  a.y = y_out;
  a.B = B; a.Hs = Hs; a.Ws = Ws; a.H = H; a.W = W; a.gray = gray;
  a.scale_x = (double)Ws / (double)W;
  a.scale_y = (double)Hs / (double)H;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) {
    a.m[c] = mean[c] * 255.0f;
    a.d[c] = 1.0f / (stdv[c] * 255.0f);
  }
  const long n = (long)B * H * W;
  const int grid = (int)std::min<long>((n + 255) / 256, 148L * 16);
  { ProfScope _ps(KC_MISC, (cudaStream_t)stream);
  preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}
