// rl_augment.cu — MyTransform(augment=True) (reference myTransforms.py:8-20) on the GPU for caller-supplied random
// parameters: HueSaturationValue -> RandomSizedCrop -> OneOf[MotionBlur, GaussNoise] -> ToGray / NoOp -> Normalize -> CHW.
// Every uint8 stage reproduces what albumentations 0.5.2 makes cv2 compute, bit for bit (pinned by the fixtures of
// oracle/make_golden_augment.py):
//   * RGB2HSV: cv2's integer algorithm with 12-bit reciprocal tables; the three cv2.LUT tables arrive in the parameters;
//     HSV2RGB: cv2's float32 vector path (tab2/tab3 with one fma, result * 255 truncated);
//   * crop + cv2.resize: the fixed-point INTER_LINEAR of rl_preprocess.cu applied to the crop window (the colour shift is
//     evaluated on the fly for the four source pixels, the shifted full frame is never materialised), INTER_NEAREST for
//     the label;
//   * MotionBlur: cv2.filter2D = fma chain over the kernel's non-zero taps in row-major order, BORDER_REFLECT_101, round
//     half to even;  GaussNoise: float32 pixel + float64 noise, clip, truncate.
// Two kernels: (1) colour shift + crop + resize -> uint8 [B][H][W][3] scratch + int64 label, (2) blur | noise, ToGray,
// Normalize, CHW float32.
#include "rl_common.cuh"
#include "rl_preprocess.cuh"

namespace rl {

struct AugArgs {
  const uint8_t* img;    // [B][Hs][Ws][3]
  const uint8_t* label;  // [B][Hs][Ws] or null
  const rl_augment_params* prm;  // [B]
  const double* noise;   // [B][H][W][3] or null
  uint8_t* scratch;      // [B][H][W][3]
  float* x;              // [B][3][H][W]
  long long* y;          // [B][H][W] or null
  int B, Hs, Ws, H, W, gray;
  float m[3], d[3];
};

// cv2 RGB2HSV_b (hsv_shift 12, hrange 180) -> cv2.LUT per plane -> cv2 HSV2RGB_b (float32 vector path)
__device__ __forceinline__ void hsv_shift_px(const rl_augment_params& p, const uint8_t* px, int* out) {
  const int r = px[0], g = px[1], b = px[2];
  const int v = max(max(r, g), b), vmin = min(min(r, g), b), diff = v - vmin;
  const int sdiv = v ? __double2int_rn((double)(255 << 12) / (double)v) : 0;
  const int hdiv = diff ? __double2int_rn((double)(180 << 12) / (6.0 * (double)diff)) : 0;
  const int s0 = (diff * sdiv + (1 << 11)) >> 12;
  int h0 = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
  h0 = (h0 * hdiv + (1 << 11)) >> 12;
  h0 += h0 < 0 ? 180 : 0;
  const int h = p.lut_h[h0 & 255], s = p.lut_s[s0 & 255], vv = p.lut_v[v];
  const float fv = __fmul_rn((float)vv, 0x1.010102p-8f);
  float cb, cg, cr;
  if (s == 0) {
    cb = cg = cr = fv;
  } else {
    const float fs = __fmul_rn((float)s, 0x1.010102p-8f);
    float fh = __fmul_rn((float)h, 0x1.111112p-5f);
    int sec = (int)floorf(fh);
    fh = __fsub_rn(fh, (float)sec);
    if ((unsigned)sec >= 6u) {
      sec = 0;
      fh = 0.f;
    }
    float tab[4];
    tab[0] = fv;
    tab[1] = __fmul_rn(fv, __fsub_rn(1.f, fs));
    tab[2] = __fmul_rn(fv, __fmaf_rn(-fs, fh, 1.f));
    tab[3] = __fmul_rn(fv, __fmaf_rn(-fs, __fsub_rn(1.f, fh), 1.f));
    // (b, g, r) table index per sector: {1,3,0} {1,0,2} {3,0,1} {0,2,1} {0,1,3} {2,1,0}
    const int ib = (0x200311 >> (4 * sec)) & 3, ig = (0x112003 >> (4 * sec)) & 3, ir = (0x031120 >> (4 * sec)) & 3;
    cb = tab[ib];
    cg = tab[ig];
    cr = tab[ir];
  }
  out[0] = max(0, min(255, (int)truncf(__fmul_rn(cr, 255.f))));
  out[1] = max(0, min(255, (int)truncf(__fmul_rn(cg, 255.f))));
  out[2] = max(0, min(255, (int)truncf(__fmul_rn(cb, 255.f))));
}

__global__ void __launch_bounds__(256) augment_crop_kernel(AugArgs a) {
  const long n = (long)a.B * a.H * a.W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % a.W), dy = (int)((i / a.W) % a.H), b = (int)(i / ((long)a.W * a.H));
    const rl_augment_params& p = a.prm[b];
    const int cw = p.crop_w, ch = p.crop_h;
    const uint8_t* src = a.img + ((size_t)b * a.Hs * a.Ws + (size_t)p.crop_y * a.Ws + p.crop_x) * 3;  // crop origin, row pitch Ws
    const double scale_x = (double)cw / (double)a.W, scale_y = (double)ch / (double)a.H;
    LinearTaps t = linear_taps(ch, cw, a.H, a.W, scale_x, scale_y, dx, dy);
    int q[4][3];
    const int ys[2] = {t.r0, t.r1}, xs[2] = {t.sx, t.sx1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* px = src + ((size_t)ys[k >> 1] * a.Ws + xs[k & 1]) * 3;
      if (p.hsv) {
        hsv_shift_px(p, px, q[k]);
      } else {
        q[k][0] = px[0];
        q[k][1] = px[1];
        q[k][2] = px[2];
      }
    }
    uint8_t* o = a.scratch + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)linear_blend(t, q[0][c], q[1][c], q[2][c], q[3][c]);
    if (a.y != nullptr) {
      int lx = dx, ly = dy;
      if (!(ch == a.H && cw == a.W)) {
        lx = min((int)floor((double)dx * scale_x), cw - 1);
        ly = min((int)floor((double)dy * scale_y), ch - 1);
      }
      a.y[i] = (long long)a.label[((size_t)b * a.Hs + p.crop_y + ly) * a.Ws + p.crop_x + lx];
    }
  }
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return max(0, min(n - 1, i));
}

__global__ void __launch_bounds__(256) augment_finish_kernel(AugArgs a) {
  const long n = (long)a.B * a.H * a.W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % a.W), dy = (int)((i / a.W) % a.H), b = (int)(i / ((long)a.W * a.H));
    const rl_augment_params& p = a.prm[b];
    const uint8_t* src = a.scratch + (size_t)b * a.H * a.W * 3;
    int v[3];
    if (p.ksize > 0) {  // cv2.filter2D: acc = fma(src, k, acc) over the non-zero taps, row-major
      const int ks = p.ksize, r = ks >> 1;
      float acc[3] = {0.f, 0.f, 0.f};
      for (int ky = 0; ky < ks; ++ky) {
        const int yy = reflect101(dy + ky - r, a.H);
        for (int kx = 0; kx < ks; ++kx) {
          const float w = p.kernel[ky * ks + kx];
          if (w == 0.f) continue;
          const uint8_t* px = src + ((size_t)yy * a.W + reflect101(dx + kx - r, a.W)) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] = __fmaf_rn((float)px[c], w, acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = max(0, min(255, __float2int_rn(acc[c])));
    } else {
      const uint8_t* px = src + ((size_t)dy * a.W + dx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = px[c];
    }
    if (p.noise && a.noise != nullptr) {  // float32 pixel + float64 noise, clip to [0, 255], truncate
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double s = __dadd_rn((double)v[c], a.noise[i * 3 + c]);
        v[c] = (int)fmin(fmax(s, 0.0), 255.0);
      }
    }
    if (a.gray) {
      const int g = (v[0] * 9798 + v[1] * 19235 + v[2] * 3735 + (1 << 14)) >> 15;
      v[0] = v[1] = v[2] = g;
    }
    const size_t plane = (size_t)a.H * a.W, o = (size_t)b * 3 * plane + (size_t)dy * a.W + dx;
#pragma unroll
    for (int c = 0; c < 3; ++c) a.x[o + c * plane] = __fmul_rn(__fsub_rn((float)v[c], a.m[c]), a.d[c]);
  }
}

}  // namespace rl

extern "C" int rl_preprocess_augment(const unsigned char* img_bgr, const unsigned char* label, int B, int Hs, int Ws, int H, int W,
                                     int gray, const rl_augment_params* params_dev, const double* noise_dev,
                                     unsigned char* scratch_u8, float* x_out, long long* y_out, void* stream) {
  using namespace rl;
  if (!img_bgr || !x_out || !params_dev || !scratch_u8 || (y_out && !label)) return rl_fail(RL_ERR_ARG, "rl_preprocess_augment: null argument");
  if (B < 1 || Hs < 1 || Ws < 1 || H < 1 || W < 1) return rl_fail(RL_ERR_SHAPE, "rl_preprocess_augment: bad shape");
  AugArgs a;
  a.img = img_bgr;
  a.label = label;
  a.prm = params_dev;
  a.noise = noise_dev;
  a.scratch = scratch_u8;
  a.x = x_out;
  a.y = y_out;
  a.B = B; a.Hs = Hs; a.Ws = Ws; a.H = H; a.W = W; a.gray = gray;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) {
    a.m[c] = mean[c] * 255.0f;
    a.d[c] = 1.0f / (stdv[c] * 255.0f);
  }
  const long n = (long)B * H * W;
  const int grid = (int)std::min<long>((n + 255) / 256, 148L * 16);
  {
    ProfScope _ps(KC_MISC, (cudaStream_t)stream);
    augment_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  {
    ProfScope _ps(KC_MISC, (cudaStream_t)stream);
    augment_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}
