// rl_conv_bwd.cu — backward tensor-core kernels (tcgen05 / TMEM):
//   dgrad_dense_kernel : input gradient of BN -> ReLU -> Conv3x3(Cin -> 16)  (autograd of reference layers.py:8-11)
//   dgrad_td_kernel    : input gradient of BN -> ReLU -> Conv1x1(C -> C)     (layers.py:46-50)
//   wgrad_kernel       : weight gradients  dW^T[c, n] = Σ_p A[p, c] · Bm[p, n]  for every convolution
//
// Input gradients are computed transposed, G^T[c, p] = Σ_k W^T[c, k] · dz[p - tap, k], so that the accumulator row
// (TMEM lane = thread) is a *channel*: the per-channel BatchNorm-backward sums Σg and Σg·x̂ are then plain
// per-thread accumulations, and every thread streams one channel column of the NHWC arenas.
// The 3x3 taps are 9 row-shifted views of one 16-channel halo tile in shared memory (start-address offsets).
#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

// ------------------------------------------------------------------------------------------------
// epilogue shared by both dgrad kernels.  Accumulator: rows (TMEM lanes = threads) are PIXELS, columns are the
// (up to 128) input channels of the current channel block, so every thread streams 32-byte channel groups of its
// own pixel from the planar arenas (x mirror, dX).  The per-channel BatchNorm-backward sums
//   S1_c = Σ_p g[p,c]      S2_c = Σ_p g[p,c]·(x[p,c] - mean_c)          (g = relu-masked input gradient)
// are reduced across the 32 pixels of a warp with transposing butterflies and kept in registers across tiles.
// ------------------------------------------------------------------------------------------------
struct DgradConsts {  // per channel of the current block, in shared memory: {s, t, gamma*rstd, mean}
  float4 k[128];
};

// Σ over the 32 lanes of v[k], k = 0..7; every lane ends up with the total of k = (lane >> 2) & 7
__device__ __forceinline__ float warp_reduce8(float* v, int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float r = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
  return r + __shfl_xor_sync(0xffffffffu, r, 1);
}

// x / dX of one 16-channel group of this thread's pixel, fetched two groups ahead of their use
struct DgradPre {
  uint4 x0, x1, d0, d1;
};
// xrow / drow_: this pixel's slot in the first 16-channel group plane of the channel block (planar layouts: dgs elements
// between group planes) or, for xrow with xgs == 0, its NHWC row
__device__ __forceinline__ void dgrad_prefetch(DgradPre& p, const __nv_bfloat16* xrow, const __nv_bfloat16* drow_, int c0,
                                               int nvalid, bool rmw, long dgs, long xgs) {
  const __nv_bfloat16* drow = drow_ + (long)(c0 >> 4) * dgs - c0;  // so that drow + c0 addresses group c0/16
  if (xgs != 0 && xrow != nullptr) xrow += (long)(c0 >> 4) * xgs - c0;  // planar activation mirror: same addressing
  if (xrow != nullptr && c0 < nvalid) {
    const u32x8 xv = ldg256_nc(xrow + c0);
    p.x0 = xv.lo;
    p.x1 = xv.hi;
    if (rmw) {
      const u32x8 dv = ld256(drow + c0);
      p.d0 = dv.lo;
      p.d1 = dv.hi;
    } else {
      p.d0 = p.d1 = make_uint4(0, 0, 0, 0);
    }
  }
}

// one 128-pixel tile x one channel block; P = this thread's pixel (or -1); S1: per-thread running totals, index =
// channel group (16 channels): lane (up, q, col) = (lane>>4, (lane>>3)&1, lane&7) keeps quantity q (0: Σg, 1: Σg(x-mean))
// of channel 8*up + col of the group.  pre[0..1] hold groups 0 and 1.  cbase: arena channel of the block's column 0.
__device__ __forceinline__ void dgrad_epilogue(uint32_t tmem_acc, int warp, int lane, const __nv_bfloat16* xrow,
                                               __nv_bfloat16* drow, int cbase, int nvalid, const DgradConsts* k,
                                               const DgradArgs& a, float* S1, DgradPre* pre) {
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) {
    const int c0 = gi * 16;
    if (c0 < nvalid) {  // warp-uniform
      float g[16];
      tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16) + c0, g);
      const DgradPre cur = pre[gi & 1];
      if (gi + 2 < 8) dgrad_prefetch(pre[gi & 1], xrow, drow, c0 + 32, nvalid, (cbase + c0 + 32) >= a.rmw_from, 16 * a.npx, a.xp ? 16 * a.npx : 0);
      // two halves of 8 channels keep the fp32 temporaries small (register budget: 3 CTAs per SM)
      uint4 dpk[2];
      float rs[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float gs[8], gx[8];
        if (xrow != nullptr) {
          float x[8], d[8];
          unpack8(h ? cur.x1 : cur.x0, x);
          unpack8(h ? cur.d1 : cur.d0, d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 kk = k->k[c0 + 8 * h + i];
            const float gv = (fmaf(x[i], kk.x, kk.y) > 0.f) ? g[8 * h + i] : 0.f;
            gs[i] = gv;
            gx[i] = gv * (x[i] - kk.w);
            d[i] = fmaf(kk.z, gv, d[i]);
          }
          dpk[h] = pack8(d);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) gs[i] = gx[i] = 0.f;
          dpk[h] = make_uint4(0, 0, 0, 0);
        }
        // butterflies: every lane ends with the total of column (lane >> 2) & 7; re-distribute to the (up, q, col)
        // convention: col = lane & 7, q = (lane >> 3) & 1
        const float r1 = warp_reduce8(gs, lane), r2 = warp_reduce8(gx, lane);
        const int src = (lane & 7) << 2;  // a lane that holds column (lane & 7)
        const float a1 = __shfl_sync(0xffffffffu, r1, src), a2 = __shfl_sync(0xffffffffu, r2, src);
        rs[h] = ((lane >> 3) & 1) ? a2 : a1;
      }
      if (xrow != nullptr) st256(drow + (long)(c0 >> 4) * 16 * a.npx, dpk[0], dpk[1]);
      S1[gi] += (lane & 16) ? rs[1] : rs[0];
    }
  }
}

__device__ __forceinline__ void dgrad_consts_init(DgradConsts* k, float* s_rstd, const DgradArgs& a, int mb) {
  float s = 0.f, t = 0.f, gr = 0.f, mean = 0.f, rstd = 0.f;
  if (mb * 128 + (int)threadIdx.x < a.Cin) {
    BnCoef b = bn_coef(a.bn, a.c_lo + mb * 128 + threadIdx.x);
    s = b.s;
    t = b.t;
    gr = b.gamma * b.rstd;
    mean = b.mean;
    rstd = b.rstd;
  }
  k->k[threadIdx.x] = make_float4(s, t, gr, mean);
  s_rstd[threadIdx.x] = rstd;
}

// block totals -> gradient slots: dβ = S1, dγ = rstd·S2, and the deferred mean terms A += γ·rstd·S1, B += γ·rstd·(rstd·S2)
__device__ __forceinline__ void dgrad_flush(float* s_sum /*[2][128]*/, const float* S1, int lane, const DgradConsts* k,
                                            const float* s_rstd, const DgradArgs& a, int mb, int nvalid) {
  for (int i = threadIdx.x; i < 256; i += 128) s_sum[i] = 0.f;
  __syncthreads();
  {
    const int cl = (lane >> 4) * 8 + (lane & 7), q = (lane >> 3) & 1;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi)
      if (gi * 16 < nvalid) atomicAdd(&s_sum[q * 128 + gi * 16 + cl], S1[gi]);
  }
  __syncthreads();
  const int cl = threadIdx.x;
  if (cl < nvalid) {
    const int c = a.c_lo + mb * 128 + cl;
    const float s1 = s_sum[cl], s2 = s_rstd[cl] * s_sum[128 + cl];
    if (a.g_beta != nullptr) {
      const int rc = ref_channel(c, a.bn.cs_perm);
      atomicAdd(a.g_beta + rc, s1);
      atomicAdd(a.g_gamma + rc, s2);
    }
    atomicAdd(a.acc + c, (double)(k->k[cl].z * s1));
    atomicAdd(a.acc + a.Ct + c, (double)(k->k[cl].z * s2));
  }
}

// ================================================================================================
// dgrad_dense_kernel:  G[p, c] = Σ_tap dz[p - tap, 0:16] · W[tap][0:16, c]   (M = 128 pixels, N = channel block, K = 16/tap)
// ================================================================================================
__global__ void __launch_bounds__(128, 3) dgrad_dense_kernel(DgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;                 // W of one channel block: [18][128][8] bf16 = 36864 B   (B operand, rows = channels)
  uint8_t* sDz = smem + 36864;        // two halo tiles [2][kDzRows][8] bf16 = 14848 B each    (A operand, rows = pixels)
  constexpr int kDzBytes = 2 * kDzRows * 16;
  DgradConsts* kc = reinterpret_cast<DgradConsts*>(smem + 36864 + 2 * kDzBytes);
  float* s_rstd = reinterpret_cast<float*>(kc + 1);
  float* s_sum = s_rstd + 128;
  __shared__ uint64_t bar_w, bar_mma, bar_dz[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int VH = a.H + 1;
  const int vrows_total = a.B * VH;
  const int run = blockIdx.x / a.nstrips, strip = blockIdx.x % a.nstrips;
  const int V0 = run * a.vrows_per_cta, V1 = min(V0 + a.vrows_per_cta, vrows_total);
  const int x0 = strip * a.strip_w, x1 = min(x0 + a.strip_w, a.W);
  const int SW = x1 - x0 + 2;
  const int t_begin = SW, t_end = (V1 - V0 + 1) * SW;  // rows V0..V1-1 of the region (region row 0 is V0-1)
  const int ntiles = (t_end - t_begin + 127) / 128;
  const int nmb = (a.Cin + 127) / 128;
  const int nrows = 128 + 2 * SW + 2;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_dz[0], 1);
    mbar_init(&bar_dz[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t n_w = 0, n_mma = 0;

  // Halo tile of dz for the tile starting at region position t0: smem row rr <-> region position t0 - SW - 1 + rr.
  // dz lives in global memory as two zero-padded planes whose position index is linear in (virtual row, column+1), so
  // for un-stripped levels the whole tile is one contiguous range per plane: a plain coalesced copy, no bounds checks
  // (pads and guards supply the zeros).
  const long PW = a.W + 2;  // padded row pitch (positions)
  const uint4* plane0 = reinterpret_cast<const uint4*>(a.dz);
  const uint4* plane1 = plane0 + a.dz_npos;
  // asynchronous copies (cp.async), issued one tile ahead into the other buffer: the copy of tile t+1 runs under the MMAs and
  // the epilogue of tile t, and no register-staged load sits in front of the fence that publishes a tile
  auto load_dz = [&](int tile, int buf) {
    const int t0 = t_begin + tile * 128;
    uint8_t* d0 = sDz + buf * kDzBytes;
    uint8_t* d1 = d0 + kDzRows * 16;
    if (SW == PW) {
      const long pos0 = a.dz_guard + (long)(V0 - 1) * PW + (t0 - SW - 1);
      for (int rr = tid; rr < nrows; rr += 128) {
        cp_async16(d0 + rr * 16, plane0 + pos0 + rr, true);
        cp_async16(d1 + rr * 16, plane1 + pos0 + rr, true);
      }
    } else {
      for (int rr = tid; rr < nrows; rr += 128) {
        const int t = t0 - SW - 1 + rr;
        const int rrow = (t >= 0) ? t / SW : -1;
        const int cc = t - rrow * SW;
        const long pos = a.dz_guard + (long)(V0 - 1 + rrow) * PW + (x0 - 1 + cc) + 1;
        cp_async16(d0 + rr * 16, plane0 + pos, true);
        cp_async16(d1 + rr * 16, plane1 + pos, true);
      }
    }
    cp_async_commit();
  };

  // channel blocks: all of them in this CTA (large levels: W^T is reused across the CTA's tiles), or one per
  // blockIdx.y (small levels, where the chain over blocks would be pure latency)
  const int mb_lo = gridDim.y > 1 ? blockIdx.y : 0, mb_hi = gridDim.y > 1 ? blockIdx.y + 1 : nmb;
  for (int mb = mb_lo; mb < mb_hi; ++mb) {
    const int nvalid = min(128, a.Cin - mb * 128);
    const uint32_t idesc = umma_idesc_bf16((uint32_t)nvalid, 0, 0);
    float S1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) S1[i] = 0.f;
    const int cb0 = a.c_lo + mb * 128;  // arena channel of accumulator column 0
    __syncthreads();  // every MMA of the previous block has been waited for: sW / constants / dz buffers may be overwritten
    dgrad_consts_init(kc, s_rstd, a, mb);
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_w, 36864);
      bulk_g2s(sW, a.wpk + (size_t)mb * 36864, 36864, &bar_w);
    }
    if (ntiles > 0) load_dz(0, 0);  // (the previous channel block's MMAs were all waited for: both buffers are free)
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = t_begin + tile * 128;
      int P = -1;  // accumulator row -> pixel
      {
        const int t = t0 + tid;
        if (t < t_end) {
          const int rrow = t / SW, cc = t - rrow * SW;
          const int vr = V0 - 1 + rrow, xc = x0 - 1 + cc;
          const int img = vr / VH, ri = vr - img * VH;
          if (ri < a.H && xc >= x0 && xc < x1) P = (img * a.H + ri) * a.W + xc;
        }
      }
      const __nv_bfloat16* xrow = P < 0 ? nullptr : (a.xp ? a.xp + dact_index(a.npx, P, cb0) : a.act + (size_t)P * a.Ct + cb0);
      __nv_bfloat16* drow = P >= 0 ? a.dact + dact_index(a.npx, P, cb0) : nullptr;
      DgradPre pre[2];
      cp_async_wait<0>();      // this thread's part of the tile (requested one tile ago)
      fence_proxy_async();
      // the epilogue's first x / dX groups are requested only now: fence.proxy.async waits for the loads a thread has
      // outstanding, so requested before it they would sit in front of the MMAs instead of behind them
      dgrad_prefetch(pre[0], xrow, drow, 0, nvalid, cb0 >= a.rmw_from, 16 * a.npx, a.xp ? 16 * a.npx : 0);
      dgrad_prefetch(pre[1], xrow, drow, 16, nvalid, (cb0 + 16) >= a.rmw_from, 16 * a.npx, a.xp ? 16 * a.npx : 0);
      __syncthreads();         // also: every thread is past the epilogue of tile-1 -> TMEM reusable
      if (tid == 0) {
        if (tile == 0) mbar_wait(&bar_w, n_w & 1);
        tc_fence_after();
        const uint32_t wa = smem_u32(sW), da = smem_u32(sDz + (tile & 1) * kDzBytes);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const uint32_t roff = (uint32_t)((1 - dy) * SW + (1 - dx));
          umma_ss(tmem, desc_kmajor(da + roff * 16, kDzRows), desc_kmajor(wa + tap * 2 * 16 * 128, 128), idesc, tap != 0);
        }
        umma_commit(&bar_mma);
      }
      // next tile's halo into the other buffer (its last readers, the MMAs of tile-1, were waited for before that epilogue)
      if (tile + 1 < ntiles) load_dz(tile + 1, (tile + 1) & 1);
      mbar_wait(&bar_mma, n_mma & 1);
      ++n_mma;
      tc_fence_after();
      dgrad_epilogue(tmem, warp, lane, xrow, drow, cb0, nvalid, kc, a, S1, pre);
      tc_fence_before();
    }
    if (ntiles == 0 && tid == 0) mbar_wait(&bar_w, n_w & 1);
    ++n_w;
    dgrad_flush(s_sum, S1, lane, kc, s_rstd, a, mb, nvalid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ================================================================================================
// dgrad_td_kernel:  G[p, c] = Σ_co dz[p, co] · W[co, c]   (M = 128 pixels, N = channel block, K = C streamed in chunks of 64)
// ================================================================================================
__global__ void __launch_bounds__(128, 3) dgrad_td_kernel(DgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stage_base = smem;  // 2 x (W chunk 16 KB + dz chunk 16 KB)
  DgradConsts* kcs = reinterpret_cast<DgradConsts*>(smem + 2 * 32768);
  float* s_rstd = reinterpret_cast<float*>(kcs + 1);
  float* s_sum = s_rstd + 128;
  __shared__ uint64_t bar_full[2], bar_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nmb = (a.Cin + 127) / 128;
  const int nkc = (a.Kdz + 63) / 64;
  const long npx = (long)a.B * a.H * a.W;

  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t g = 0;
  TileLoader<128> ld;
  const __nv_bfloat16* rowptr[4];

  for (int mb = 0; mb < nmb; ++mb) {
    const int nvalid = min(128, a.Cin - mb * 128);
    const uint32_t idesc = umma_idesc_bf16((uint32_t)nvalid, 0, 0);
    float S1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) S1[i] = 0.f;
    __syncthreads();
    dgrad_consts_init(kcs, s_rstd, a, mb);
    __syncthreads();
    for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long Pr = tile * 128 + TileLoader<128>::row_of(i);
        rowptr[i] = Pr < npx ? a.dz + (size_t)Pr * a.Kdz : nullptr;
      }
      const long Pl = tile * 128 + tid;
      const int P = Pl < npx ? (int)Pl : -1;
      const __nv_bfloat16* xrow = P < 0 ? nullptr : (a.xp ? a.xp + dact_index(a.npx, P, mb * 128) : a.act + (size_t)P * a.Ct + mb * 128);
      __nv_bfloat16* drow = P >= 0 ? a.dact + dact_index(a.npx, P, mb * 128) : nullptr;
      DgradPre pre[2];
      ld.load(rowptr, min(8, a.Kdz / 8));
      for (int kc = 0; kc < nkc; ++kc) {
        const int ck = min(64, a.Kdz - kc * 64);
        const uint32_t s = g & 1, u = g >> 1;
        uint8_t* sW = stage_base + s * 32768;
        uint8_t* sB = sW + 16384;
        if (u > 0) mbar_wait(&bar_done[s], (u - 1) & 1);
        if (tid == 0) {
          const uint32_t wbytes = (uint32_t)(ck / 8) * 128 * 16;
          mbar_arrive_expect_tx(&bar_full[s], wbytes);
          bulk_g2s(sW, a.wpk + ((size_t)mb * nkc + kc) * 16384, wbytes, &bar_full[s]);
        }
        ld.store(sB, rowptr, ck / 8, nullptr, nullptr);
        fence_proxy_async();  // before the next loads are issued (the fence waits for outstanding memory operations)
        if (kc + 1 < nkc) {
          const __nv_bfloat16* rp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] = rowptr[i] ? rowptr[i] + (kc + 1) * 64 : nullptr;
          ld.load(rp, min(8, (a.Kdz - (kc + 1) * 64) / 8));
        } else {
          // after the tile's last fence (fence.proxy.async waits for outstanding loads): the epilogue's first x / dX groups
          dgrad_prefetch(pre[0], xrow, drow, 0, nvalid, (mb * 128) >= a.rmw_from, 16 * a.npx, a.xp ? 16 * a.npx : 0);
          dgrad_prefetch(pre[1], xrow, drow, 16, nvalid, (mb * 128 + 16) >= a.rmw_from, 16 * a.npx, a.xp ? 16 * a.npx : 0);
        }
        __syncthreads();
        if (tid == 0) {
          mbar_wait(&bar_full[s], u & 1);
          tc_fence_after();
          const uint32_t wa = smem_u32(sW), ba = smem_u32(sB);
          for (int kk = 0; kk < ck / 16; ++kk)
            umma_ss(tmem, desc_kmajor(ba + kk * 2 * 16 * 128, 128), desc_kmajor(wa + kk * 2 * 16 * 128, 128), idesc,
                    (kc | kk) != 0);
          umma_commit(&bar_done[s]);
        }
        ++g;
      }
      {
        const uint32_t s = (g - 1) & 1, u = (g - 1) >> 1;
        mbar_wait(&bar_done[s], u & 1);
        tc_fence_after();
      }
      dgrad_epilogue(tmem, warp, lane, xrow, drow, mb * 128, nvalid, kcs, a, S1, pre);
      tc_fence_before();
      __syncthreads();
    }
    dgrad_flush(s_sum, S1, lane, kcs, s_rstd, a, mb, nvalid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

constexpr size_t kDgradDenseSmem = 36864 + 2 * (2 * kDzRows * 16) + sizeof(DgradConsts) + 384 * sizeof(float);
constexpr size_t kDgradTdSmem = 2 * 32768 + sizeof(DgradConsts) + 384 * sizeof(float);

int launch_dgrad(const DgradArgs& a, int grid, cudaStream_t st) {
  if (a.dense) {
    if (a.strip_w + 2 > 165) return rl_fail(RL_ERR_SHAPE, "dgrad: strip too wide");
    const int nmb = (a.Cin + 127) / 128;
    const long vrows = (long)a.B * (a.H + 1);
    const bool split = nmb > 1 && vrows * (a.W + 2) <= 40000;  // small level: channel blocks in parallel CTAs (rl_api.cu sizes grid.x accordingly)
    ProfScope _ps(KC_DGRAD_DENSE, st);
    dgrad_dense_kernel<<<dim3(grid, split ? nmb : 1), 128, kDgradDenseSmem, st>>>(a);
  } else {
    ProfScope _ps(KC_DGRAD_TD, st);
    dgrad_td_kernel<<<grid, 128, kDgradTdSmem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// wgrad_kernel:  D[c (128 per M-block), n] = Σ_p A[p, c] · Bm[p, n], K = pixels in chunks of 64
// ================================================================================================
__device__ __forceinline__ long wgrad_index(const WgradArgs& a, int c, int ng) {
  switch (a.kind) {
    case WG_DENSE: return ((long)(ng & 15) * a.Cin_w + ref_channel(c, a.cs_perm)) * 9 + (ng >> 4);
    case WG_TD: return (long)ng * a.Cin_w + c;
    case WG_HEAD: return ng < a.num_cls ? (long)ng * 288 + ref_channel(c, 128) : -1;
    default: return (long)ng * 27 + c;  // WG_FIRST
  }
}

constexpr int kWgThreads = 256;

__global__ void __launch_bounds__(kWgThreads, 2) wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int mb = blockIdx.y, part = blockIdx.z;
  const int n0 = a.part_n0[part], np = a.part_np[part];
  const int ca = min(128, a.Ca - mb * 128);  // valid channels of this M-block
  float* s_s = reinterpret_cast<float*>(smem);
  float* s_t = s_s + 128;
  uint8_t* stage_base = smem + 1024;
  const int stageB = np * 128;  // np/8 chunks x 64 rows x 16 B
  const int kStageBytes = 16384 + stageB;
  __shared__ uint64_t bar_done[3];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    mbar_init(&bar_done[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid < 128) {
    float s = 0.f, t = 0.f;
    if (a.transform && tid < ca) {
      BnCoef k = bn_coef(a.bn, mb * 128 + tid);
      s = k.s;
      t = k.t;
    }
    s_s[tid] = s;
    s_t[tid] = t;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16((uint32_t)np, 1, 1);

  const long P0 = (long)blockIdx.x * a.px_per_cta, P1 = min(P0 + a.px_per_cta, a.npx);
  const int nchunks = (int)((P1 - P0 + 63) / 64);
  const int a_nchunk = (ca + 7) / 8;
  const long HW = (long)a.H * a.W;
  const int nj = (a.kind == WG_DENSE) ? 18 : np / 8;  // 16-byte column chunks of the B tile
  // A tile: 64 pixel rows x 16 channel chunks -> two 32-byte items (= 2 adjacent chunks = one 16-channel group) per thread:
  // row = 8*warp + (lane&7), group = (lane>>3) + 4*(h>>1), chunk = 2*group + (h&1)
  const int a_row = warp * 8 + (lane & 7), a_j0 = lane >> 3;
  // B tile: row = tid & 63, column chunks j = (tid >> 6) + 4q
  const int b_row = tid & 63, b_j0 = tid >> 6;
  struct ARegs {
    uint4 v[4];
    bool valid;
  };

  // 3-stage pipeline, prefetch distance 2 chunks: A through registers (it needs the BN+ReLU transform),
  // B by cp.async straight into its operand layout
  auto load_a = [&](int kc, ARegs& r) {
    const long P = P0 + (long)kc * 64 + a_row;
    r.valid = kc < nchunks && P < P1;
    const long Pv = r.valid ? P : P0;
    const __nv_bfloat16* src = a.A + (size_t)Pv * a.lda + a.a_off + mb * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int j = 2 * (a_j0 + 4 * hh);  // even chunk of the group; the odd one follows in both layouts
      const __nv_bfloat16* pj = a.Axp ? a.Axp + dact_index(a.npx, Pv, a.a_off + mb * 128 + j * 8) : src + j * 8;
      if (r.valid && j < a_nchunk) {
        const u32x8 v = ldg256_nc_nohint(pj);
        r.v[2 * hh] = v.lo;
        r.v[2 * hh + 1] = v.hi;
      } else {
        r.v[2 * hh] = r.v[2 * hh + 1] = make_uint4(0, 0, 0, 0);
      }
    }
  };
  // DenseLayer B tile: this thread's pixel row advances by 64 pixels per chunk; its position in the zero-padded dz planes
  // is carried along (one division at the start, none in the loop).  The tap at (dy,dx) of pixel (b,y,x) sits at
  // pos - dy*(W+2) - dx, always in range; the 5 (tap, half) column chunks of a thread are fixed, so their plane offsets are too.
  long bpos = 0;   // plane position of pixel P0 + kc_next*64 + b_row
  int bx = 0, by = 0;
  long boff[5];
  if (a.kind == WG_DENSE) {
    const long P = P0 + b_row;
    const long b = P / HW;
    const int hw = (int)(P - b * HW);
    by = hw / a.W;
    bx = hw - by * a.W;
    bpos = a.dz_guard + (b * (a.H + 1) + by) * (a.W + 2) + bx + 1;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int j = b_j0 + 4 * q, tap = j >> 1, half = j & 1;
      boff[q] = (long)half * a.dz_npos - (long)(tap / 3 - 1) * (a.W + 2) - (tap % 3 - 1);
    }
  }
  auto issue_b = [&](int kc, uint8_t* sB) {
    if (kc < nchunks) {
      const long P = P0 + (long)kc * 64 + b_row;
      const bool pv = P < P1;
      if (a.kind == WG_DENSE) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int j = b_j0 + 4 * q;
          if (j < 18) cp_async16(sB + ((size_t)j * 64 + b_row) * 16, a.Bm + ((pv ? bpos : (long)a.dz_guard) + boff[q]) * 8, pv);
        }
        // advance to the next chunk (issue_b is called with consecutive kc)
        bx += 64;
        bpos += 64;
        while (bx >= a.W) {
          bx -= a.W;
          bpos += 2;  // the two pad columns of a plane row
          if (++by == a.H) {
            by = 0;
            bpos += a.W + 2;  // the pad row between images
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = b_j0 + 4 * q;
          if (j < nj) cp_async16(sB + ((size_t)j * 64 + b_row) * 16, pv ? a.Bm + (size_t)P * a.ldb + n0 + j * 8 : a.Bm, pv);
        }
      }
    }
    cp_async_commit();  // always: keeps the group count uniform
  };
  auto step = [&](int k, ARegs& r) {
    const int s = k % 3;
    uint8_t* sA = stage_base + s * kStageBytes;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int j = 2 * (a_j0 + 4 * (h >> 1)) + (h & 1);
      if (j < a_nchunk) {
        uint4 o = r.v[h];
        if (a.transform && r.valid) o = bn_relu8(o, s_s + j * 8, s_t + j * 8);
        *reinterpret_cast<uint4*>(sA + ((size_t)j * 64 + a_row) * 16) = o;
      }
    }
    // fence BEFORE the next global loads are issued: fence.proxy.async waits for the thread's outstanding memory operations
    cp_async_wait<1>();  // B(k) has landed (B(k+1) may still be in flight)
    fence_proxy_async();
    // stage (k+2)%3 == (k-1)%3 is recycled for chunk k+2 once the MMAs of chunk k-1 have retired
    if (k >= 1) mbar_wait(&bar_done[(k - 1) % 3], ((k - 1) / 3) & 1);
    load_a(k + 2, r);
    issue_b(k + 2, stage_base + ((k + 2) % 3) * kStageBytes + 16384);
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(sA), ba = smem_u32(sA + 16384);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem, desc_mnmajor(aa + kk * 256, 64), desc_mnmajor(ba + kk * 256, 64), idesc, (k | kk) != 0);
      umma_commit(&bar_done[s]);
    }
  };

  ARegs r0, r1;
  load_a(0, r0);
  issue_b(0, stage_base + 16384);
  load_a(1, r1);
  issue_b(1, stage_base + kStageBytes + 16384);
  for (int kc = 0; kc < nchunks; kc += 2) {
    step(kc, r0);
    if (kc + 1 < nchunks) step(kc + 1, r1);
  }
  cp_async_wait<0>();
  if (nchunks > 0 && warp < 4) {
    const uint32_t s = (nchunks - 1) % 3, u = (nchunks - 1) / 3;
    mbar_wait(&bar_done[s], u & 1);
    tc_fence_after();
    const int c = mb * 128 + tid;
    for (int c0 = 0; c0 < np; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      if (tid < ca) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const long idx = wgrad_index(a, c, n0 + c0 + i);
          if (idx >= 0) atomicAdd(a.gw + idx, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int launch_wgrad(const WgradArgs& a, int nruns, int nmb, int nparts, cudaStream_t st) {
  int npmax = 0;
  for (int i = 0; i < nparts; ++i) npmax = a.part_np[i] > npmax ? a.part_np[i] : npmax;
  size_t smem = 1024 + 3 * (16384 + (size_t)npmax * 128);
  if (smem > 160 * 1024) return rl_fail(RL_ERR_SHAPE, "wgrad: smem %zu too large", smem);
  {
    ProfScope _ps(KC_WGRAD, st);
    wgrad_kernel<<<dim3(nruns, nmb, nparts), kWgThreads, smem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int configure_bwd_kernels() {
  RL_CUDA(cudaFuncSetAttribute(dgrad_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024));
  RL_CUDA(cudaFuncSetAttribute(dgrad_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024));
  RL_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  return RL_OK;
}

}  // namespace rl
