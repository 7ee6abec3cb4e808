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
    k.rstd = rsqrtf((float)v + kBnEps);  // fp64 only for the cancellation-prone E[x^2] - E[x]^2; fp32 rsqrt afterwards
  }
  k.gamma = b.gamma[rc];
  k.s = k.gamma * k.rstd;
  k.t = b.beta[rc] - k.mean * k.s;
  return k;
}

// Gradient arenas are planar by 16-channel group: [Ct/16][pixels][16] bf16.  A pixel's 16 channels of a group are 32
// contiguous bytes and consecutive pixels are contiguous, so the lane = pixel accesses of the input-gradient epilogues
// touch 8 lines per warp instead of 32 and every channel window is a contiguous stream.  Element index of channel c
// (a multiple of 8) of pixel P in an arena of npx pixels:
__host__ __device__ __forceinline__ size_t dact_index(long npx, long P, int c) {
  return ((size_t)(c >> 4) * (size_t)npx + (size_t)P) * 16 + (size_t)(c & 8);
}

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// relu(x*s+t) on 8 packed bf16 with per-channel s,t taken from shared memory
__device__ __forceinline__ uint4 bn_relu8(uint4 v, const float* s8, const float* t8) {
  const float4 s0 = *reinterpret_cast<const float4*>(s8), s1 = *reinterpret_cast<const float4*>(s8 + 4);
  const float4 t0 = *reinterpret_cast<const float4*>(t8), t1 = *reinterpret_cast<const float4*>(t8 + 4);
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  uint4 o;
  o.x = pack_bf16x2_relu(fmaf(a.x, s0.x, t0.x), fmaf(a.y, s0.y, t0.y));
  o.y = pack_bf16x2_relu(fmaf(b.x, s0.z, t0.z), fmaf(b.y, s0.w, t0.w));
  o.z = pack_bf16x2_relu(fmaf(c.x, s1.x, t1.x), fmaf(c.y, s1.y, t1.y));
  o.w = pack_bf16x2_relu(fmaf(d.x, s1.z, t1.z), fmaf(d.y, s1.w, t1.w));
  return o;
}

// ------------------------------------------------------------------------------------------------
// Tile loader, 128 threads: 128 rows x up to 64 channels (8 chunks) of an NHWC arena -> chunk layout T[j][row][8].
// Lane mapping: lane = (r8 = lane & 7, q = lane >> 3).  A lane moves 32 contiguous bytes (chunks 2q, 2q+1) of row
// 32*warp + 8*i + r8 with one 256-bit load, so 4 lanes cover a whole 128-byte line of a row and 8 consecutive lanes
// store 8 consecutive rows of one chunk plane (conflict-free 128 B shared stores).  nchunk is always even here
// (channel counts are multiples of 16).
// ------------------------------------------------------------------------------------------------
template <int ROWS>
struct TileLoader {
  static_assert(ROWS == 128, "only the 128-row tile is used");
  u32x8 v[4];

  __device__ __forceinline__ static int row_of(int i) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return w * 32 + 8 * i + (lane & 7);
  }
  // rowptr[i]: pointer to channel 0 of the chunk window for row row_of(i), or nullptr (row is zero)
  __device__ __forceinline__ void load(const __nv_bfloat16* const* rowptr, int nchunk) {
    const int q = (threadIdx.x & 31) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (rowptr[i] != nullptr && 2 * q < nchunk) {
        v[i] = ldg256_nc_nohint(rowptr[i] + 16 * q);
      } else {
        v[i].lo = v[i].hi = make_uint4(0, 0, 0, 0);
      }
    }
  }
  // s / t: per-channel arrays (shared memory) already offset to the first channel of this window; null = raw copy
  __device__ __forceinline__ void store(uint8_t* smem, const __nv_bfloat16* const* rowptr, int nchunk, const float* s,
                                        const float* t) const {
    const int q = (threadIdx.x & 31) >> 3;
    if (2 * q >= nchunk) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 lo = v[i].lo, hi = v[i].hi;
      if (s != nullptr && rowptr[i] != nullptr) {
        lo = bn_relu8(lo, s + 16 * q, t + 16 * q);
        hi = bn_relu8(hi, s + 16 * q + 8, t + 16 * q + 8);
      }
      const int row = row_of(i);
      *reinterpret_cast<uint4*>(smem + ((size_t)(2 * q) * ROWS + row) * 16) = lo;
      *reinterpret_cast<uint4*>(smem + ((size_t)(2 * q + 1) * ROWS + row) * 16) = hi;
    }
  }
};

__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// halo tile of the zero-padded dz16 planes: 128 + 2*(strip+2) + 2 <= 464 rows for strips up to 165 columns
constexpr int kDzRows = 464;

// Σ over the 32 lanes of v[k] for k = 0..15 (v is clobbered); lane l ends up holding the total of k = (l >> 1) & 15
__device__ __forceinline__ float warp_reduce16(float* v, int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rl
