// rl_misc.cu — memory-bound CUDA-core kernels: weight packing, first conv (K=27), channel statistics,
// transposed-conv col2im, the fused classifier head (normalize -> 1x1 -> /T -> softmax -> CE-on-probabilities ->
// argmax -> confusion counts, forward and backward), output-gradient preparation (deferred BatchNorm
// correction, max-pool scatter, transposed-conv column matrix), BN running statistics, AdamW / SGD, dropout masks.
#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ================================================================================================
// weight packing (fp32 reference layouts -> bf16 chunk-layout operands), one table-driven launch
// ================================================================================================
__global__ void pack_weights_kernel(const PackEntry* tab, const float* __restrict__ params, uint8_t* ws) {
  const PackEntry e = tab[blockIdx.y];
  const float* w = params + e.p_w;
  uint4* dst = reinterpret_cast<uint4*>(ws + e.dst_off);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < e.n16; i += (long)gridDim.x * blockDim.x) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    const int C = e.cin;
    switch (e.kind) {
      case 0: {  // dense fwd: [kc][jc][n=slot*16+co][8 c], tap slots ordered dx-major: slot = (dx+1)*3 + (dy+1)
        int kc = (int)(i / (8 * 144)), rem = (int)(i % (8 * 144)), jc = rem / 144, n = rem % 144;
        int c0 = kc * 64 + jc * 8, slot = n / 16, co = n % 16;
        int tap = (slot % 3) * 3 + slot / 3;  // reference tap index ky*3+kx
        if (c0 < C)
          for (int k = 0; k < 8; ++k) v[k] = w[((long)co * C + ref_channel(c0 + k, e.cs_perm)) * 9 + tap];
        break;
      }
      case 1: {  // dense bwd: [mb][j=tap*2+half][cl][8 co]
        int mb = (int)(i / (18 * 128)), rem = (int)(i % (18 * 128)), j = rem / 128, cl = rem % 128;
        int tap = j >> 1, half = j & 1, c = mb * 128 + cl;
        if (c < C)
          for (int k = 0; k < 8; ++k) v[k] = w[((long)(half * 8 + k) * C + ref_channel(c, e.cs_perm)) * 9 + tap];
        break;
      }
      case 2: {  // TD fwd: parts back to back, [kc][jc][nl][8 c], nl = accumulator column
        int nkc = (C + 63) / 64;
        long p0 = (long)nkc * 8 * e.np0;
        int n0 = 0, np = e.np0;
        long il = i;
        if (i >= p0) {
          il = i - p0;
          n0 = e.np0;
          np = C - e.np0;
        }
        int kc = (int)(il / (8 * np)), jc = (int)((il % (8 * np)) / np), nl = (int)(il % np);
        // accumulator column nl = 16 blk + 8k + 2q + e holds output channel 16 blk + 4q + 2k + e: the pooling epilogue's
        // tcgen05.ld.16x256b hands thread (lane % 4 = q) the columns 8k + 2q + e, i.e. four contiguous channels
        int c0 = kc * 64 + jc * 8, co = n0 + (nl & ~15) + 4 * ((nl >> 1) & 3) + 2 * ((nl >> 3) & 1) + (nl & 1);
        for (int k = 0; k < 8; ++k)
          if (c0 + k < C) v[k] = w[(long)co * C + c0 + k];
        break;
      }
      case 3: {  // TD bwd: [mb][kc][jc][cl][8 co]
        int nkc = (C + 63) / 64;
        int mb = (int)(i / ((long)nkc * 8 * 128)), rem = (int)(i % ((long)nkc * 8 * 128));
        int kc = rem / (8 * 128), jc = (rem / 128) % 8, cl = rem % 128;
        int c = mb * 128 + cl, co0 = kc * 64 + jc * 8;
        if (c < C)
          for (int k = 0; k < 8; ++k)
            if (co0 + k < C) v[k] = w[(long)(co0 + k) * C + c];
        break;
      }
      case 6: {  // dense bwd, chain launch: [j=tap*2+half][cl][8 co], arena channel c_lo + cl (rows >= nch stay zero)
        int j = (int)(i / 128), cl = (int)(i % 128);
        int tap = j >> 1, half = j & 1, c = e.c_lo + cl;
        if (cl < e.nch)
          for (int k = 0; k < 8; ++k) v[k] = w[((long)(half * 8 + k) * C + ref_channel(c, e.cs_perm)) * 9 + tap];
        break;
      }
      case 7: {  // dense bwd, block-input kernel: [cb][j=tap*2+half][cl][8 co], channel cb*cbs + cl < C0 (c_lo = C0, nch = cbs)
        int cb = (int)(i / (18 * kBulkCb)), rem = (int)(i % (18 * kBulkCb)), j = rem / kBulkCb, cl = rem % kBulkCb;
        int tap = j >> 1, half = j & 1, c = cb * e.nch + cl;
        if (cl < e.nch && c < e.c_lo)
          for (int k = 0; k < 8; ++k) v[k] = w[((long)(half * 8 + k) * C + ref_channel(c, e.cs_perm)) * 9 + tap];
        break;
      }
      case 9: {  // TU input gradient: [tap][co chunk j][ci][8 co]; weight [ci][co][3][3]
        int tap = (int)(i / 800), rem = (int)(i % 800), j = rem / 80, ci = rem % 80;
        for (int k = 0; k < 8; ++k) v[k] = w[((long)ci * 80 + (j * 8 + k)) * 9 + tap];
        break;
      }
      case 8: {  // TU fwd, parity kernel: [tap][ci chunk j][co][8 ci]; weight [ci][co][3][3]
        int tap = (int)(i / 800), rem = (int)(i % 800), j = rem / 80, co = rem % 80;
        for (int k = 0; k < 8; ++k) v[k] = w[((long)(j * 8 + k) * 80 + co) * 9 + tap];
        break;
      }
    }
    dst[i] = pack8(v);
  }
}

int launch_pack_weights(const Plan* p, const float* params, cudaStream_t st) {
  const PackEntry* tab = reinterpret_cast<const PackEntry*>((const uint8_t*)p->ws + p->packtab_off);
  { ProfScope _ps(KC_PACK, st);
  pack_weights_kernel<<<dim3(24, p->n_pack), 256, 0, st>>>(tab, params, (uint8_t*)p->ws); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// first convolution: Conv3x3(3 -> 48, pad 1) on fp32 NCHW input (reference tiramisu.py:33-35)
// ================================================================================================
__global__ void __launch_bounds__(256) first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* act, int Ct, int B,
                                                        int H, int W, __nv_bfloat16* xcol, __nv_bfloat16* xp) {
  __shared__ float sw[48 * 27 + 48];
  for (int i = threadIdx.x; i < 48 * 27; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < 48; i += blockDim.x) sw[48 * 27 + i] = bias[i];
  __syncthreads();
  const long npx = (long)B * H * W;
  const long HW = (long)H * W;
  for (long P = blockIdx.x * (long)blockDim.x + threadIdx.x; P < npx; P += (long)gridDim.x * blockDim.x) {
    const int P32 = (int)P, iHW = (int)HW, b = P32 / iHW, hw = P32 - b * iHW, y = hw / W, xx = hw - y * W;
    float in[27];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = y + ky - 1, xq = xx + kx - 1;
          in[ci * 9 + ky * 3 + kx] =
              (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(x + ((long)(b * 3 + ci) * H + yy) * W + xq) : 0.f;
        }
    if (xcol != nullptr) {  // bf16 im2col row kept for the weight gradient (27 taps + 5 zero columns)
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(2 * i < 27 ? in[2 * i] : 0.f, 2 * i + 1 < 27 ? in[2 * i + 1] : 0.f);
      uint4* d = reinterpret_cast<uint4*>(xcol + P * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    }
    uint4* dst = reinterpret_cast<uint4*>(act + P * Ct);
#pragma unroll
    for (int g = 0; g < 6; ++g) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int co = g * 8 + k;
        float acc = sw[48 * 27 + co];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc = fmaf(in[i], sw[co * 27 + i], acc);
        o[k] = acc;
      }
      dst[g] = pack8(o);
      if (xp != nullptr) *reinterpret_cast<uint4*>(xp + dact_index(npx, P, g * 8)) = dst[g];
    }
  }
}

int launch_first_fwd(const Plan* p, const float* x, const float* params, int train, cudaStream_t st) {
  const Layer& L = p->layers[0];
  __nv_bfloat16* act = (__nv_bfloat16*)((uint8_t*)p->ws + p->lv[0].act_off);
  __nv_bfloat16* xcol = p->training ? (__nv_bfloat16*)((uint8_t*)p->ws + p->xcol_off) : nullptr;
  { ProfScope _ps(KC_FIRST_FWD, st);
  __nv_bfloat16* xp = (p->training && train) ? (__nv_bfloat16*)((uint8_t*)p->ws + p->lv[0].xp_off) : nullptr;
  first_fwd_kernel<<<p->sm_count * 4, 256, 0, st>>>(x, params + L.p_w, params + L.p_b, act, p->lv[0].Ct, p->B, p->H, p->W,
                                                    xcol, xp); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// per-channel Σx, Σx² of an arena slice (channels [c_off, c_off+nch), nch % 8 == 0)
// ================================================================================================
__global__ void __launch_bounds__(256) channel_stats_kernel(const __nv_bfloat16* __restrict__ arena, long npx, int Ct,
                                                            int c_off, int nch, double* stats) {
  extern __shared__ float sred[];  // [blockDim][16]
  const int nchunk = nch / 8;
  const int ppb = blockDim.x / nchunk;  // pixel lanes per block
  const int j = threadIdx.x % nchunk, pl = threadIdx.x / nchunk;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  if (pl < ppb) {
    for (long P = (long)blockIdx.x * ppb + pl; P < npx; P += (long)gridDim.x * ppb) {
      uint4 v = ldg128(arena + P * Ct + c_off + j * 8);
      float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += f[k];
        q[k] += f[k] * f[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sred[threadIdx.x * 16 + k] = s[k];
    sred[threadIdx.x * 16 + 8 + k] = q[k];
  }
  __syncthreads();
  // thread t < nch*2 reduces one (channel, moment) over the pixel lanes
  for (int o = threadIdx.x; o < nch * 2; o += blockDim.x) {
    const int mom = o / nch, c = o % nch, jj = c / 8, k = c % 8;
    float acc = 0.f;
    for (int l = 0; l < ppb; ++l) acc += sred[(l * nchunk + jj) * 16 + mom * 8 + k];
    atomicAdd(stats + (mom ? Ct : 0) + c_off + c, (double)acc);
  }
}

int launch_channel_stats(const __nv_bfloat16* arena, long npx, int Ct, int c_off, int nch, double* stats, int sm,
                         cudaStream_t st) {
  int nchunk = nch / 8, ppb = 256 / nchunk;
  int grid = (int)std::min<long>((npx + ppb - 1) / ppb, (long)sm * 4);
  { ProfScope _ps(KC_STATS, st);
  channel_stats_kernel<<<grid, 256, 256 * 16 * sizeof(float), st>>>(arena, npx, Ct, c_off, nch, stats); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}


// ================================================================================================
// label histogram (getClassWeight, reference TrainingBase.py:12-23: integer counts, reciprocal later)
// ================================================================================================
__global__ void __launch_bounds__(256) label_hist_kernel(const void* y, int is64, long n, int num_cls, Scalars* sc) {
  __shared__ int h[9];
  if (threadIdx.x < 9) h[threadIdx.x] = 0;
  __syncthreads();
  int loc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    long v = is64 ? reinterpret_cast<const long long*>(y)[i] : (long)reinterpret_cast<const uint8_t*>(y)[i];
    int k = (v >= 0 && v < num_cls) ? (int)v : 8;
#pragma unroll
    for (int c = 0; c < 9; ++c) loc[c] += (k == c);
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    int s = loc[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&h[c], s);
  }
  __syncthreads();
  if (threadIdx.x < 8 && h[threadIdx.x]) atomicAdd(&sc->hist[threadIdx.x], h[threadIdx.x]);
  if (threadIdx.x == 8 && h[8]) atomicAdd(&sc->bad_label, h[8]);
}

int launch_label_hist(const void* y, int y_is_i64, long n, int num_cls, Scalars* sc, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 8);
  { ProfScope _ps(KC_MISC, st);
  label_hist_kernel<<<grid, 256, 0, st>>>(y, y_is_i64, n, num_cls, sc); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// classifier head, forward and backward
//   F.normalize(dim=1, eps=1e-12) (tiramisu.py:105) -> Conv1x1 288->C (+b) -> / T -> softmax (tiramisu.py:120-125)
//   -> cross_entropy applied to the *probabilities* (SimpleTrain.py:16, TrainingBase.py:85) -> argmax (first max)
//   -> confusion counts.  One warp handles 8 pixels: a quad of lanes per pixel, 9 chunks of 8 channels per lane.
// ================================================================================================
template <int C>
__global__ void __launch_bounds__(256, 2) head_kernel(HeadArgs a) {
  __shared__ float sw[C * 288];  // physical channel order
  __shared__ float sb[C], scw[C];
  __shared__ int scm[C * C];
  __shared__ float sred[8][2 + C];
  __shared__ float s_den;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C * 288; i += 256) {
    const int k = i / 288, c = i % 288;
    sw[i] = a.w[k * 288 + ref_channel(c, 128)];
  }
  if (tid < C) {
    sb[tid] = a.b[tid];
    const int hcount = a.sc->hist[tid];
    scw[tid] = a.weighted ? 1.0f / (float)hcount : 1.0f;  // 1/0 = inf for absent classes, as torch.reciprocal
  }
  for (int i = tid; i < C * C; i += 256) scm[i] = 0;
  if (tid == 0) {
    float den = 0.f;
    if (a.weighted) {
      for (int k = 0; k < C; ++k)
        if (a.sc->hist[k] > 0) den += (float)a.sc->hist[k] * (1.0f / (float)a.sc->hist[k]);
    } else {
      den = (float)a.npx;
    }
    s_den = den;
  }
  __syncthreads();
  const int q = lane & 3, pi = lane >> 2;
  float loss_num = 0.f, loss_den = 0.f, db[C];
#pragma unroll
  for (int k = 0; k < C; ++k) db[k] = 0.f;
  const float inv_den = 1.0f / s_den;
  const float gscale = a.grad_scale * (a.grad_scale_dev ? __ldg(a.grad_scale_dev) : 1.0f);

  for (long base = ((long)blockIdx.x * 8 + warp) * 8; base < a.npx; base += (long)gridDim.x * 64) {
    const long P = base + pi;
    const bool valid = P < a.npx;
    uint4 f[9];
    float nsq = 0.f, dot[C];
#pragma unroll
    for (int k = 0; k < C; ++k) dot[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int j = q + 4 * i;
      f[i] = valid ? ldg128(a.feat + P * a.Ct + j * 8) : make_uint4(0, 0, 0, 0);
      const uint32_t u[4] = {f[i].x, f[i].y, f[i].z, f[i].w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float2 v = unpack_bf16x2(u[h]);
        nsq = fmaf(v.x, v.x, fmaf(v.y, v.y, nsq));
#pragma unroll
        for (int k = 0; k < C; ++k) dot[k] = fmaf(v.x, sw[k * 288 + j * 8 + 2 * h], fmaf(v.y, sw[k * 288 + j * 8 + 2 * h + 1], dot[k]));
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      nsq += __shfl_xor_sync(0xffffffffu, nsq, o);
#pragma unroll
      for (int k = 0; k < C; ++k) dot[k] += __shfl_xor_sync(0xffffffffu, dot[k], o);
    }
    const float nrm = fmaxf(sqrtf(nsq), 1e-12f);
    const float inv_n = 1.0f / nrm;
    float z[C], p[C];
    float zmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      z[k] = (dot[k] * inv_n + sb[k]) / 0.05f;
      zmax = fmaxf(zmax, z[k]);
    }
    float esum = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      p[k] = expf(z[k] - zmax);
      esum += p[k];
    }
    int pred = 0;
    float pmax = -1.f;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      p[k] = p[k] / esum;
      if (p[k] > pmax) {
        pmax = p[k];
        pred = k;
      }
    }
    const int b = valid ? (int)P / a.HW : 0, hw = valid ? (int)P - b * a.HW : 0;
    if (valid && q == 0) {
      if (a.probs)
#pragma unroll
        for (int k = 0; k < C; ++k) a.probs[((long)b * C + k) * a.HW + hw] = p[k];
      if (a.mask_out) a.mask_out[P] = (uint8_t)pred;
    }
    int label = -1;
    float sm2[C];  // softmax of the probabilities (second softmax inside cross_entropy)
    float wl = 0.f;
    if (a.y != nullptr && valid) {
      long lv = a.y_is_i64 ? reinterpret_cast<const long long*>(a.y)[P] : (long)reinterpret_cast<const uint8_t*>(a.y)[P];
      label = (lv >= 0 && lv < C) ? (int)lv : -1;
    }
    {
      float e2 = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        sm2[k] = expf(p[k] - pmax);
        e2 += sm2[k];
      }
      const float lse = logf(e2) + pmax;
#pragma unroll
      for (int k = 0; k < C; ++k) sm2[k] = sm2[k] / e2;
      if (label >= 0) {
        float pl = 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) pl = (k == label) ? p[k] : pl;
#pragma unroll
        for (int k = 0; k < C; ++k) wl = (k == label) ? scw[k] : wl;
        if (q == 0) {
          loss_num += wl * (lse - pl);
          loss_den += wl;
          atomicAdd(&scm[label * C + pred], 1);
        }
      }
    }
    if (a.dact != nullptr) {
      // d(loss)/d(probs)
      float dp[C];
#pragma unroll
      for (int k = 0; k < C; ++k) {
        if (a.dprobs)
          dp[k] = valid ? a.dprobs[((long)b * C + k) * a.HW + hw] : 0.f;
        else
          dp[k] = (label >= 0) ? gscale * wl * inv_den * (sm2[k] - (k == label ? 1.f : 0.f)) : 0.f;
      }
      float dsum = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) dsum = fmaf(dp[k], p[k], dsum);
      float dl[C], gdot = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float dlogit = p[k] * (dp[k] - dsum) / 0.05f;
        if (q == 0 && valid) db[k] += dlogit;
        dl[k] = dlogit * inv_n;
        gdot = fmaf(dl[k], dot[k], gdot);
      }
      const float coef = gdot * inv_n * inv_n;
      if (valid) {
        if (q == 0) {
          uint32_t pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int k = 0; k < C; k += 2) pk[k >> 1] = pack_bf16x2(dl[k], k + 1 < C ? dl[k + 1] : 0.f);
          uint4* d = reinterpret_cast<uint4*>(a.dl + P * 16);
          d[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          d[1] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int j = q + 4 * i;
          const uint32_t u[4] = {f[i].x, f[i].y, f[i].z, f[i].w};
          float o[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float2 v = unpack_bf16x2(u[h]);
            float g0 = -coef * v.x, g1 = -coef * v.y;
#pragma unroll
            for (int k = 0; k < C; ++k) {
              g0 = fmaf(dl[k], sw[k * 288 + j * 8 + 2 * h], g0);
              g1 = fmaf(dl[k], sw[k * 288 + j * 8 + 2 * h + 1], g1);
            }
            o[2 * h] = g0 * a.feat_grad_sign;
            o[2 * h + 1] = g1 * a.feat_grad_sign;
          }
          *reinterpret_cast<uint4*>(a.dact + dact_index(a.npx, P, j * 8)) = pack8(o);
        }
      }
    }
  }
  // ---- block reductions
  loss_num = warp_sum(loss_num);
  loss_den = warp_sum(loss_den);
#pragma unroll
  for (int k = 0; k < C; ++k) db[k] = warp_sum(db[k]);
  if (lane == 0) {
    sred[warp][0] = loss_num;
    sred[warp][1] = loss_den;
#pragma unroll
    for (int k = 0; k < C; ++k) sred[warp][2 + k] = db[k];
  }
  __syncthreads();
  if (tid < 2 + C) {
    float v = 0.f;
    for (int w2 = 0; w2 < 8; ++w2) v += sred[w2][tid];
    if (tid == 0 && a.y) atomicAdd(&a.sc->loss_num, (double)v);
    if (tid == 1 && a.y) atomicAdd(&a.sc->loss_den, (double)v);
    if (tid >= 2 && a.g_b) atomicAdd(a.g_b + tid - 2, v);
  }
  if (a.y)
    for (int i = tid; i < C * C; i += 256)
      if (scm[i]) atomicAdd(&a.sc->cm[i], scm[i]);
}

int launch_head(const HeadArgs& a, int sm, cudaStream_t st) {
  int grid = (int)std::min<long>((a.npx + 63) / 64, (long)sm * 8);
  ProfScope _ps(KC_HEAD, st);
  switch (a.num_cls) {
    case 2: head_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: head_kernel<3><<<grid, 256, 0, st>>>(a); break;
    case 4: head_kernel<4><<<grid, 256, 0, st>>>(a); break;
    default: return rl_fail(RL_ERR_SHAPE, "head: num_cls %d unsupported", a.num_cls);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// output-gradient preparation:   dz = dropout_scale * (dX - A_c - xhat * B_c)
// (A_c, B_c): BatchNorm-backward mean terms summed over all consumer layers of channel c.
// ================================================================================================
__global__ void __launch_bounds__(256) dz_prepare_kernel(DzArgs a) {
  extern __shared__ float sm[];  // reduction scratch [256][8]
  float* s_red = sm;
  const int nchunk = a.nch / 8;
  // every thread owns one fixed chunk of 8 channels: its BatchNorm constants and deferred-correction coefficients
  // live in registers for the whole kernel
  const int jown = threadIdx.x % nchunk;
  float c_mean[8], c_rstd[8], c_A[8], c_B[8], mk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int gc = a.c_off + jown * 8 + k;
    double m = a.stats[gc] * a.inv_n_out, v = a.stats[a.Ct + gc] * a.inv_n_out - m * m;
    if (v < 0) v = 0;
    c_mean[k] = (float)m;
    c_rstd[k] = rsqrtf((float)v + kBnEps);
    c_A[k] = (float)(a.acc[gc] * a.inv_n_out);
    c_B[k] = (float)(a.acc[a.Ct + gc] * a.inv_n_out);
    mk[k] = 1.f;
  }
  int mk_img = -1;
  float bsum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
  int jfix = -1;

  // j is always this thread's own chunk (jown)
  auto corrected = [&](long Pout, int j, int img, float* o) {
    float dx[8], xv[8];
    unpack8(ldg128(a.dact + dact_index(a.npx_out, Pout, a.c_off + j * 8)), dx);
    unpack8(ldg128(a.xp ? a.xp + dact_index(a.npx_out, Pout, a.c_off + j * 8) : a.act + Pout * a.Ct + a.c_off + j * 8), xv);
    if (a.mask != nullptr && img != mk_img) {
      const float4* mp = reinterpret_cast<const float4*>(a.mask + (long)img * a.nch + j * 8);
      const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
      mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w;
      mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
      mk_img = img;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (xv[k] - c_mean[k]) * c_rstd[k];
      o[k] = mk[k] * (dx[k] - c_A[k] - xh * c_B[k]);
    }
  };

  if (a.kind == 0) {
    const int ppb = blockDim.x / nchunk;
    const int j = threadIdx.x % nchunk, pl = threadIdx.x / nchunk;
    jfix = j;
    if (pl < ppb)
      for (long P = (long)blockIdx.x * ppb + pl; P < a.npx_out; P += (long)gridDim.x * ppb) {
        float o[8];
        const int img = (int)P / a.HW_out;
        corrected(P, j, img, o);
        if (a.planes) {
          const int hw = (int)(P - (long)img * a.HW_out), y = hw / a.W, x = hw - y * a.W;
          const long pos = a.dz_guard + ((long)img * (a.H + 1) + y) * (a.W + 2) + x + 1;
          *reinterpret_cast<uint4*>(a.dz + ((long)j * a.dz_npos + pos) * 8) = pack8(o);
        } else {
          *reinterpret_cast<uint4*>(a.dz + P * a.nch + j * 8) = pack8(o);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) bsum[k] += o[k];
      }
  } else if (a.kind == 1) {
    // TD: loop over input-resolution pixels; gradient goes to the arg-max position of each 2x2 window
    const int ppb = blockDim.x / nchunk;
    const int j = threadIdx.x % nchunk, pl = threadIdx.x / nchunk;
    jfix = j;
    const long npx_in = (long)a.B * a.H * a.W;
    if (pl < ppb)
      for (long P = (long)blockIdx.x * ppb + pl; P < npx_in; P += (long)gridDim.x * ppb) {
        const int x = (int)(P % a.W), y = (int)((P / a.W) % a.H), b = (int)(P / ((long)a.W * a.H));
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
        const int py = y >> 1, px = x >> 1;
        if (py < a.Hp && px < a.Wp) {
          const long Pout = ((long)b * a.Hp + py) * a.Wp + px;
          const int qq = (y & 1) * 2 + (x & 1);
          float c8[8];
          corrected(Pout, j, b, c8);
          const uint2 am = *reinterpret_cast<const uint2*>(a.amax + Pout * a.nch + j * 8);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t word = k < 4 ? am.x : am.y;
            const int sel = (word >> (8 * (k & 3))) & 0xff;
            o[k] = (sel == qq) ? c8[k] : 0.f;
            bsum[k] += o[k];
          }
        }
        *reinterpret_cast<uint4*>(a.dz + P * a.nch + j * 8) = pack8(o);
      }
  } else {
    // TU: column matrix row (b,i,j) holds dz at output pixels (2i+ky, 2j+kx) for the 9 taps.
    // thread <-> fixed channel chunk j; loops over (input pixel, tap) pairs, so the bias sum stays in registers
    const int ppb = blockDim.x / nchunk;
    const int j = threadIdx.x % nchunk, pl = threadIdx.x / nchunk;
    const long nitems = (long)a.B * a.Hp * a.Wp * 9;
    if (pl < ppb)
      for (long it = (long)blockIdx.x * ppb + pl; it < nitems; it += (long)gridDim.x * ppb) {
        const long Pin = it / 9;
        const int tap = (int)(it - Pin * 9);
        const int jj = (int)(Pin % a.Wp), ii = (int)((Pin / a.Wp) % a.Hp), b = (int)(Pin / ((long)a.Wp * a.Hp));
        const int ky = tap / 3, kx = tap % 3, Y = 2 * ii + ky, X = 2 * jj + kx;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
        if (Y < a.H && X < a.W) {
          corrected(((long)b * a.H + Y) * a.W + X, j, b, o);
          // every output pixel appears in 1, 2 or 4 (pixel, tap) slots: count it once for the bias gradient
          const bool count = (ky < 2 || ii == a.Hp - 1) && (kx < 2 || jj == a.Wp - 1);
          if (count) {
#pragma unroll
            for (int k = 0; k < 8; ++k) bsum[k] += o[k];
          }
        }
        *reinterpret_cast<uint4*>(a.dz + Pin * 720 + tap * a.nch + j * 8) = pack8(o);
      }
  }
  if (a.g_bias != nullptr) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = bsum[k];
    __syncthreads();
    const int ppb = blockDim.x / nchunk;
    for (int c = threadIdx.x; c < a.nch; c += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < ppb; ++l) acc += s_red[(l * nchunk + c / 8) * 8 + (c % 8)];
      atomicAdd(a.g_bias + c, acc);
    }
  }
  (void)jfix;
}

// ---- specialised variants (one pass per producing layer; all index math hoisted, one `corrected` per value) ----
struct DzConst {  // per-channel constants of `corrected`: mean, rstd, A_c, B_c
  float mean, rstd, A, B;
};
__device__ __forceinline__ DzConst dz_const(const DzArgs& a, int c) {
  const int gc = a.c_off + c;
  double m = a.stats[gc] * a.inv_n_out, v = a.stats[a.Ct + gc] * a.inv_n_out - m * m;
  if (v < 0) v = 0;
  DzConst k;
  k.mean = (float)m;
  k.rstd = rsqrtf((float)v + kBnEps);
  k.A = (float)(a.acc[gc] * a.inv_n_out);
  k.B = (float)(a.acc[a.Ct + gc] * a.inv_n_out);
  return k;
}

// DenseLayer output (16 channels): one thread per pixel, 32-byte accesses, zero-padded two-plane result
__global__ void __launch_bounds__(256) dz_dense16_kernel(DzArgs a) {
  __shared__ float s_red[8][16];
  __shared__ float2 s_k[16];  // dz = mask * (dX - A' - x * B'),  B' = rstd * B,  A' = A - mean * B'
  if (threadIdx.x < 16) {
    const DzConst k = dz_const(a, threadIdx.x);
    const float b2 = k.rstd * k.B;
    s_k[threadIdx.x] = make_float2(k.A - k.mean * b2, b2);
  }
  __syncthreads();
  float2 k[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) k[c] = s_k[c];
  float bsum[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) bsum[c] = 0.f;
  const int npx = (int)a.npx_out;  // < 2^31: 32-bit index arithmetic
  const __nv_bfloat16* dplane = a.dact + (size_t)(a.c_off >> 4) * (size_t)a.npx_out * 16;
  const __nv_bfloat16* xplane = a.xp ? a.xp + (size_t)(a.c_off >> 4) * (size_t)a.npx_out * 16 : nullptr;
  for (int P = blockIdx.x * blockDim.x + threadIdx.x; P < npx; P += gridDim.x * blockDim.x) {
    const int img = P / a.HW_out;
    const int hw = P - img * a.HW_out, y = hw / a.W, x = hw - y * a.W;
    const u32x8 dv = ld256_nohint(dplane + (size_t)P * 16);
    const u32x8 xv = ldg256_nc_nohint(xplane ? xplane + (size_t)P * 16 : a.act + (size_t)P * a.Ct + a.c_off);
    float dx[16], xx[16], mk[16], o[16];
    unpack8(dv.lo, dx);
    unpack8(dv.hi, dx + 8);
    unpack8(xv.lo, xx);
    unpack8(xv.hi, xx + 8);
    if (a.mask != nullptr) {
      const float4* mp = reinterpret_cast<const float4*>(a.mask + (long)img * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 m4 = __ldg(mp + q);
        mk[4 * q] = m4.x; mk[4 * q + 1] = m4.y; mk[4 * q + 2] = m4.z; mk[4 * q + 3] = m4.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) mk[c] = 1.f;
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      o[c] = mk[c] * (dx[c] - fmaf(xx[c], k[c].y, k[c].x));
      bsum[c] += o[c];
    }
    const long pos = a.dz_guard + ((long)img * (a.H + 1) + y) * (a.W + 2) + x + 1;
    *reinterpret_cast<uint4*>(a.dz + pos * 8) = pack8(o);
    *reinterpret_cast<uint4*>(a.dz + (a.dz_npos + pos) * 8) = pack8(o + 8);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float v = warp_sum(bsum[c]);
    if (lane == 0) s_red[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16 && a.g_bias != nullptr) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    atomicAdd(a.g_bias + threadIdx.x, v);
  }
}

// TransitionDown output: one thread per (pooled pixel, 8-channel chunk) scatters to the 4 pixels of its 2x2 window
__global__ void __launch_bounds__(256) dz_td_kernel(DzArgs a) {
  extern __shared__ float sm[];
  const int nchunk = a.nch / 8, ppb = blockDim.x / nchunk;
  const int j = threadIdx.x % nchunk, pl = threadIdx.x / nchunk;
  DzConst* s_k = reinterpret_cast<DzConst*>(sm + 256 * 8);  // [nch]
  for (int c = threadIdx.x; c < a.nch; c += blockDim.x) s_k[c] = dz_const(a, c);
  __syncthreads();
  DzConst k[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) k[c] = s_k[j * 8 + c];
  float bsum[8], mk[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    bsum[c] = 0.f;
    mk[c] = 1.f;
  }
  int mk_img = -1;
  const long npool = (long)a.B * a.Hp * a.Wp;
  if (pl < ppb)
    for (long Pp = (long)blockIdx.x * ppb + pl; Pp < npool; Pp += (long)gridDim.x * ppb) {
      const int Pp32 = (int)Pp, b = Pp32 / (a.Wp * a.Hp), rem = Pp32 - b * (a.Wp * a.Hp), py = rem / a.Wp, px = rem - py * a.Wp;
      if (a.mask != nullptr && b != mk_img) {
        const float4* mp = reinterpret_cast<const float4*>(a.mask + (long)b * a.nch + j * 8);
        const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
        mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w; mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
        mk_img = b;
      }
      float dx[8], xv[8], c8[8];
      unpack8(ldg128(a.dact + dact_index(a.npx_out, Pp, a.c_off + j * 8)), dx);
      unpack8(ldg128(a.xp ? a.xp + dact_index(a.npx_out, Pp, a.c_off + j * 8) : a.act + Pp * a.Ct + a.c_off + j * 8), xv);
      const uint2 am = *reinterpret_cast<const uint2*>(a.amax + Pp * a.nch + j * 8);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float xh = (xv[c] - k[c].mean) * k[c].rstd;
        c8[c] = mk[c] * (dx[c] - k[c].A - xh * k[c].B);
        bsum[c] += c8[c];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t word = c < 4 ? am.x : am.y;
          o[c] = (((word >> (8 * (c & 3))) & 0xff) == (uint32_t)q) ? c8[c] : 0.f;
        }
        const long P = ((long)b * a.H + 2 * py + (q >> 1)) * a.W + 2 * px + (q & 1);
        *reinterpret_cast<uint4*>(a.dz + P * a.nch + j * 8) = pack8(o);
      }
    }
  // rows / columns outside every pooling window (odd H or W) get no gradient
  const int xw = a.W - 2 * a.Wp, xh = a.H - 2 * a.Hp;
  if (xw > 0 || xh > 0) {
    const long nextra = (long)a.B * ((long)xh * a.W + (long)(a.H - xh) * xw);
    for (long e = (long)blockIdx.x * ppb + pl; pl < ppb && e < nextra; e += (long)gridDim.x * ppb) {
      const long per = (long)xh * a.W + (long)(a.H - xh) * xw;
      const int b = (int)(e / per);
      long r = e - (long)b * per;
      int y, x;
      if (r < (long)xh * a.W) {
        y = 2 * a.Hp + (int)(r / a.W);
        x = (int)(r % a.W);
      } else {
        r -= (long)xh * a.W;
        y = (int)(r / xw);
        x = 2 * a.Wp + (int)(r % xw);
      }
      *reinterpret_cast<uint4*>(a.dz + (((long)b * a.H + y) * a.W + x) * a.nch + j * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  if (a.g_bias != nullptr) {
#pragma unroll
    for (int c = 0; c < 8; ++c) sm[threadIdx.x * 8 + c] = bsum[c];
    __syncthreads();
    for (int c = threadIdx.x; c < a.nch; c += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < ppb; ++l) acc += sm[(l * nchunk + c / 8) * 8 + (c % 8)];
      atomicAdd(a.g_bias + c, acc);
    }
  }
}

// TransitionUp output (80 channels): dz is written as four PARITY PLANES over the low-resolution position grid
// (I, J) in [0, h] x [0, w]:  plane (py, px) at position (b, I, J) holds dz of output pixel (2I + py, 2J + px), or zeros when
// that pixel lies outside the cropped output.  Layout [4][B][h+1][w+1][80] bf16: the transposed convolution's backward
// GEMMs (rl_tu.cu) read row-shifted views of tiles of these planes, the [pixels x 720] column matrix never exists.
// One thread per (position, 16-channel group); it handles the four parities of its position (one 2x2 output quad) with
// 32-byte accesses (a whole group of the planar arenas per load).
__global__ void __launch_bounds__(250) dz_tu_kernel(DzArgs a) {
  extern __shared__ float sm[];
  const int g = threadIdx.x % 5, pl = threadIdx.x / 5;   // 5 groups of 16 channels, 50 position lanes
  const int hp = a.Hp + 1, wp = a.Wp + 1;
  DzConst* s_k = reinterpret_cast<DzConst*>(sm + 250 * 16);  // [80]
  for (int c = threadIdx.x; c < 80; c += blockDim.x) s_k[c] = dz_const(a, c);
  __syncthreads();
  float kA[16], kB[16];  // dz = dX - A' - x * B',  B' = rstd * B,  A' = A - mean * B'
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const DzConst k = s_k[g * 16 + c];
    kB[c] = k.rstd * k.B;
    kA[c] = k.A - k.mean * kB[c];
  }
  float bsum[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) bsum[c] = 0.f;
  const long npos = (long)a.B * hp * wp;
  const __nv_bfloat16* dplane = a.dact + (size_t)((a.c_off >> 4) + g) * (size_t)a.npx_out * 16;
  const __nv_bfloat16* xplane = a.xp ? a.xp + (size_t)((a.c_off >> 4) + g) * (size_t)a.npx_out * 16 : nullptr;
  for (long q = (long)blockIdx.x * 50 + pl; q < npos; q += (long)gridDim.x * 50) {
    const int q32 = (int)q, b = q32 / (hp * wp), rem = q32 - b * (hp * wp), I = rem / wp, J = rem - I * wp;
#pragma unroll
    for (int par = 0; par < 4; ++par) {
      const int Y = 2 * I + (par >> 1), X = 2 * J + (par & 1);
      uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
      if (Y < a.H && X < a.W) {
        const long P = ((long)b * a.H + Y) * a.W + X;
        const u32x8 dv = ld256_nohint(dplane + (size_t)P * 16);
        const u32x8 xv = ldg256_nc_nohint(xplane ? xplane + (size_t)P * 16 : a.act + (size_t)P * a.Ct + a.c_off + g * 16);
        float dx[16], xx[16], o[16];
        unpack8(dv.lo, dx);
        unpack8(dv.hi, dx + 8);
        unpack8(xv.lo, xx);
        unpack8(xv.hi, xx + 8);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          o[c] = dx[c] - fmaf(xx[c], kB[c], kA[c]);
          bsum[c] += o[c];
        }
        lo = pack8(o);
        hi = pack8(o + 8);
      }
      st256(a.dz + ((long)par * npos + q) * 80 + g * 16, lo, hi);
    }
  }
  if (a.g_bias != nullptr) {
#pragma unroll
    for (int c = 0; c < 16; ++c) sm[threadIdx.x * 16 + c] = bsum[c];
    __syncthreads();
    for (int c = threadIdx.x; c < 80; c += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < 50; ++l) acc += sm[(l * 5 + c / 16) * 16 + (c % 16)];
      atomicAdd(a.g_bias + c, acc);
    }
  }
}

int launch_dz_prepare(const DzArgs& a, int sm, cudaStream_t st) {
  if (a.kind == 0 && a.planes && a.nch == 16) {
    int grid = (int)std::min<long>((a.npx_out + 255) / 256, (long)sm * 8);
    { ProfScope _ps(KC_DZ_PREPARE, st);
    dz_dense16_kernel<<<grid, 256, 0, st>>>(a); }
  RL_CUDA(cudaGetLastError());
    return RL_OK;
  }
  if (a.kind == 1) {
    const int nchunk = a.nch / 8, ppb = 256 / nchunk;
    long npool = (long)a.B * a.Hp * a.Wp;
    int grid = (int)std::max<long>(1, std::min<long>((npool + ppb - 1) / ppb, (long)sm * 8));
    { ProfScope _ps(KC_DZ_PREPARE, st);
    dz_td_kernel<<<grid, 256, 256 * 8 * sizeof(float) + a.nch * sizeof(DzConst), st>>>(a); }
  RL_CUDA(cudaGetLastError());
    return RL_OK;
  }
  if (a.kind == 2 && a.nch == 80) {
    if (a.c_off & 15) return rl_fail(RL_ERR_ARG, "dz_tu: the up-sampled channels must start on a 16-channel group");
    long npos = (long)a.B * (a.Hp + 1) * (a.Wp + 1);
    int grid = (int)std::max<long>(1, std::min<long>((npos + 49) / 50, (long)sm * 8));
    { ProfScope _ps(KC_DZ_PREPARE, st);
    dz_tu_kernel<<<grid, 250, 250 * 16 * sizeof(float) + 80 * sizeof(DzConst), st>>>(a); }
  RL_CUDA(cudaGetLastError());
    return RL_OK;
  }
  size_t smem = (size_t)(256 * 8) * sizeof(float);
  long work = (a.kind == 0) ? a.npx_out : (a.kind == 1 ? (long)a.B * a.H * a.W : (long)a.B * a.Hp * a.Wp * 9);
  int nchunk = a.nch / 8;
  int ppb = 256 / nchunk;
  long blocks = (work + ppb - 1) / ppb;
  int grid = (int)std::min<long>(blocks, (long)sm * 8);
  if (grid < 1) grid = 1;
  { ProfScope _ps(KC_DZ_PREPARE, st);
  dz_prepare_kernel<<<grid, 256, smem, st>>>(a); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// BatchNorm running statistics (momentum 0.1, unbiased variance), all 60 layers in one launch
// ================================================================================================
__global__ void bn_running_kernel(const BnTabEntry* tab, const uint8_t* ws, float* bnbuf, float momentum) {
  const BnTabEntry e = tab[blockIdx.x];
  const double* stats = reinterpret_cast<const double*>(ws + e.stats_byte_off);
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    double m = stats[c] * (double)e.inv_n, v = stats[e.Ct + c] * (double)e.inv_n - m * m;
    if (v < 0) v = 0;
    const int rc = ref_channel(c, e.cs_perm);
    bnbuf[e.rm + rc] = (1.f - momentum) * bnbuf[e.rm + rc] + momentum * (float)m;
    bnbuf[e.rv + rc] = (1.f - momentum) * bnbuf[e.rv + rc] + momentum * (float)(v * (double)e.unbias);
  }
}

int launch_bn_running(const Plan* p, float* bnbuf, float momentum, cudaStream_t st) {
  const BnTabEntry* tab = reinterpret_cast<const BnTabEntry*>((const uint8_t*)p->ws + p->bntab_off);
  { ProfScope _ps(KC_MISC, st);
  bn_running_kernel<<<p->n_bn, 256, 0, st>>>(tab, (const uint8_t*)p->ws, bnbuf, momentum); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// optimizers over the flat parameter arena (reference SimpleTrain.py:28 AdamW; MMETrainingModule.py:16-19 SGD)
// ================================================================================================
// (1 - beta), the decay factor, the step size and 1/sqrt(bias correction 2) are formed in double on the host, as
// torch.optim.AdamW forms them in Python floats, and only then rounded to fp32
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long n, float decay, float b1, float b2, float eps,
                                                    float omb1, float omb2, float step_size, float inv_sqrt_bc2, float gs) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * decay;
    const float mi = fmaf(b1, m[i], omb1 * gi);
    const float vi = fmaf(b2, v[i], omb2 * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

// the same update with every hyper-parameter read from device memory: hyper = double[6] {lr, beta1, beta2, eps, weight_decay,
// step}; nothing of the launch depends on host state, so a captured CUDA graph stays valid while lr and step change
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long n, const double* __restrict__ hyper, float gs) {
  const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5];
  const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
  const float decay = (float)(1.0 - lr * wd), fb1 = (float)b1, fb2 = (float)b2, feps = (float)eps, omb1 = (float)(1.0 - b1),
              omb2 = (float)(1.0 - b2), step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * decay;
    const float mi = fmaf(fb1, m[i], omb1 * gi);
    const float vi = fmaf(fb2, v[i], omb2 * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + feps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}
int launch_adamw_dev(float* p, const float* g, float* m, float* v, long n, const double* hyper, float grad_scale, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 8);
  { ProfScope _ps(KC_OPT, st);
  adamw_dev_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, hyper, grad_scale); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int launch_adamw(float* p, const float* g, float* m, float* v, long n, double lr, double b1, double b2, double eps, double wd,
                 long step, float grad_scale, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 8);
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  { ProfScope _ps(KC_OPT, st);
  adamw_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, (float)(1.0 - lr * wd), (float)b1, (float)b2, (float)eps, (float)(1.0 - b1),
                                     (float)(1.0 - b2), (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), grad_scale); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

__global__ void __launch_bounds__(256) sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ buf, long n, float lr, float mom, float wd,
                                                           int first, float gs) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float d = g[i] * gs + wd * p[i];
    float b = first ? d : mom * buf[i] + d;
    buf[i] = b;
    p[i] -= lr * (d + mom * b);
  }
}

int launch_sgd_nesterov(float* p, const float* g, float* buf, long n, float lr, float momentum, float wd, int first,
                        float grad_scale, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 8);
  { ProfScope _ps(KC_OPT, st);
  sgd_nesterov_kernel<<<grid, 256, 0, st>>>(p, g, buf, n, lr, momentum, wd, first, grad_scale); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// grads += scale * stage: end of a backward pass (the raw gradients of one pass are staged so that the loss's grad_output,
// a device scalar, can be applied without the head kernel having to wait for it)
__global__ void __launch_bounds__(256) grad_accum_kernel(float* __restrict__ g, const float* __restrict__ stage, long n, float scale,
                                                         const float* __restrict__ scale_dev) {
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) g[i] = fmaf(sc, stage[i], g[i]);
}
int launch_grad_accum(float* grads, const float* stage, long n, float scale, const float* scale_dev, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 8);
  { ProfScope _ps(KC_OPT, st);
  grad_accum_kernel<<<grid, 256, 0, st>>>(grads, stage, n, scale, scale_dev); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// Dropout2d(0.2) keep-scales for every (layer, sample, channel): counter-based hash RNG
// ================================================================================================
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void fill_masks_kernel(float* masks, long n, float p_drop, unsigned long long seed, unsigned long long step,
                                  const long long* step_dev) {
  if (step_dev != nullptr) step += (unsigned long long)*step_dev;  // device-side step counter: the launch can live in a CUDA graph
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned long long h = splitmix64(splitmix64(seed ^ (step * 0xD1B54A32D192ED03ull)) + (unsigned long long)i);
    float u = (float)(h >> 40) * (1.0f / 16777216.0f);
    masks[i] = (u >= p_drop) ? 1.0f / (1.0f - p_drop) : 0.f;
  }
}
int launch_fill_masks(float* masks, long n, float p_drop, unsigned long long seed, unsigned long long step,
                      const long long* step_dev, cudaStream_t st) {
  int grid = (int)std::min<long>((n + 255) / 256, 148L * 4);
  { ProfScope _ps(KC_MISC, st);
  fill_masks_kernel<<<grid, 256, 0, st>>>(masks, n, p_drop, seed, step, step_dev); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

}  // namespace rl
