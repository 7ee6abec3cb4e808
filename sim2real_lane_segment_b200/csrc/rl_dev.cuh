// rl_dev.cuh — device helpers shared by the kernels: BatchNorm coefficient derivation from the shared
// per-channel fp64 sums, tile loaders into the "chunk layout" of rl_umma.cuh, small reductions.
#pragma once
#include "rl_plan.h"
#include "rl_umma.cuh"

namespace rl {

constexpr float kBnEps = 1e-5f;

// Everything a consumer needs to turn channel c of an arena into relu(x*s+t) (reference layers.py:8-9,46-47).
struct BnSrc {
  const double* stats;  // [2][Ct]: Σx, Σx² over the level's pixels (training) — null in eval mode
  int Ct;
  double inv_n;         // 1 / (B*H*W)
  const float* gamma;   // reference-ordered
  const float* beta;
  const float* rmean;   // running stats (eval mode)
  const float* rvar;
  int cs_perm;
  int eval;
};

struct BnCoef {
  float s, t, mean, rstd, gamma;
};

__device__ __forceinline__ BnCoef bn_coef(const BnSrc& b, int c) {
  BnCoef k;
  const int rc = ref_channel(c, b.cs_perm);
  if (b.eval) {
    k.mean = b.rmean[rc];
    k.rstd = rsqrtf(b.rvar[rc] + kBnEps);
  } else {
    double m = b.stats[c] * b.inv_n;
    double v = b.stats[b.Ct + c] * b.inv_n - m * m;
    if (v < 0.0) v = 0.0;
    k.mean = (float)m;
    k.rstd = (float)(1.0 / sqrt(v + (double)kBnEps));
  }
  k.gamma = b.gamma[rc];
  k.s = k.gamma * k.rstd;
  k.t = b.beta[rc] - k.mean * k.s;
  return k;
}

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// relu(x*s+t) on 8 packed bf16 with per-channel s,t taken from shared memory
__device__ __forceinline__ uint4 bn_relu8(uint4 v, const float* s8, const float* t8) {
  const float4 s0 = *reinterpret_cast<const float4*>(s8), s1 = *reinterpret_cast<const float4*>(s8 + 4);
  const float4 t0 = *reinterpret_cast<const float4*>(t8), t1 = *reinterpret_cast<const float4*>(t8 + 4);
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  uint4 o;
  o.x = pack_bf16x2(fmaxf(fmaf(a.x, s0.x, t0.x), 0.f), fmaxf(fmaf(a.y, s0.y, t0.y), 0.f));
  o.y = pack_bf16x2(fmaxf(fmaf(b.x, s0.z, t0.z), 0.f), fmaxf(fmaf(b.y, s0.w, t0.w), 0.f));
  o.z = pack_bf16x2(fmaxf(fmaf(c.x, s1.x, t1.x), 0.f), fmaxf(fmaf(c.y, s1.y, t1.y), 0.f));
  o.w = pack_bf16x2(fmaxf(fmaf(d.x, s1.z, t1.z), 0.f), fmaxf(fmaf(d.y, s1.w, t1.w), 0.f));
  return o;
}

// ------------------------------------------------------------------------------------------------
// Tile loader, 128 threads: ROWS rows x up to 64 channels (8 chunks) of an NHWC arena -> chunk layout
// T[j][row][8] with ROWS rows.  Lane mapping: 8 consecutive lanes cover 8 consecutive rows of one chunk
// (conflict-free 128 B shared stores); the 4 lane groups cover 4 consecutive chunks (64 B per row per
// request: whole 32 B sectors).  ROWS = 128: 4 row groups x 2 chunk halves; ROWS = 64: 2 x 4.
// ------------------------------------------------------------------------------------------------
template <int ROWS>
struct TileLoader {
  static constexpr int RG = ROWS / 32;       // row groups per thread
  static constexpr int CH = 8 / RG;          // chunk slots per thread (RG*CH = 8 loads)
  static constexpr int CSTEP = 4;            // chunk stride between slots
  uint4 v[8];

  __device__ __forceinline__ static int row_of(int i) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return w * (ROWS / 4) + 8 * i + (lane & 7);
  }
  __device__ __forceinline__ static int chunk_of(int h) { return ((threadIdx.x & 31) >> 3) + CSTEP * h; }

  // rowptr[i]: pointer to channel 0 of the chunk window for row row_of(i), or nullptr (row is zero)
  __device__ __forceinline__ void load(const __nv_bfloat16* const* rowptr, int nchunk) {
#pragma unroll
    for (int i = 0; i < RG; ++i)
#pragma unroll
      for (int h = 0; h < CH; ++h) {
        const int j = chunk_of(h);
        uint4 z = make_uint4(0, 0, 0, 0);
        if (rowptr[i] != nullptr && j < nchunk) z = ldg128(rowptr[i] + j * 8);
        v[i * CH + h] = z;
      }
  }
  // st: per-channel s / t arrays (shared memory) already offset to the first channel of this window; null = raw copy
  __device__ __forceinline__ void store(uint8_t* smem, const __nv_bfloat16* const* rowptr, int nchunk, const float* s,
                                        const float* t) const {
#pragma unroll
    for (int i = 0; i < RG; ++i)
#pragma unroll
      for (int h = 0; h < CH; ++h) {
        const int j = chunk_of(h);
        if (j < nchunk) {
          uint4 o = v[i * CH + h];
          if (s != nullptr && rowptr[i] != nullptr) o = bn_relu8(o, s + j * 8, t + j * 8);
          *reinterpret_cast<uint4*>(smem + ((size_t)j * ROWS + row_of(i)) * 16) = o;
        }
      }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rl
