// rl_bwd_fused.cu — backward of a whole DenseBlock (reference layers.py:18-40, autograd of layers.py:5-15) in two kernels
// that replace ten per-layer launches:
//
//   dgrad_bulk_kernel   : input gradient of all five DenseLayers with respect to the block's INPUT channels [0, C0)
//                           dX[p,c] (+)= Σ_j γ_j r_c · relu'_j(p,c) · (dz_j ⊛ W_j)[p,c]
//                         x and dX are streamed ONCE per block instead of once per layer (the per-layer kernel moved
//                         6·Cin bytes per pixel and layer; this one moves 6·C0 + 5·32 per pixel and block).  The layers'
//                         gradients with respect to the block's own new channels stay in the per-layer chain
//                         (dgrad_dense_kernel restricted to [C0, C0+16j)), because dz_{j-1} depends on them.
//   wgrad_dense2_kernel : weight gradients of the five layers in one launch and, from the same accumulators, the
//                         BatchNorm parameter gradients.  The relu-masked, normalised activation is split into two
//                         tensor-core operands built from x on the fly, m·x̂ and m (m = relu mask, exact in bf16):
//                           D2 = (m·x̂)ᵀ·im2col(dz)   D1 = mᵀ·im2col(dz)          (fp32 in TMEM)
//                           dW = γ·D2 + β·D1      dγ = Σ_n W[n,·]·D2[·,n]      dβ = Σ_n W[n,·]·D1[·,n]
//                         so the per-layer pixel sums Σg, Σg·x̂ of BatchNorm backward are contractions on the tensor
//                         pipe and no input-gradient epilogue has to reduce them with shuffles.
//
// The deferred mean terms (A_c, B_c, rl_plan.h) of the block-input channels only need the pixel sums of the TOTAL
// gradient the bulk kernel adds, Σ_p Δ and Σ_p Δ·x̂: one butterfly per 16 channels and tile instead of one per layer.
#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

// ================================================================================================
// dgrad_bulk_kernel — one CTA per SM, 10 warps:
//   warp 8    : loader. The five layers' packed weights of this CTA's channel block once (bulk TMA); per tile the five
//               dz halo tiles (two planes each, one contiguous range per plane: un-stripped levels only) by bulk TMA.
//   warp 9    : MMA issue. Per tile and half (<= 48 channels): 5 layers x 9 taps of M = 128 pixels, N = half, K = 16
//               into one of two TMEM buffers (5 x 48 columns each).
//   warps 0-7 : epilogue, set s = warp / 4 drains buffer s (lane quarter warp % 4): relu masks of the five layers from
//               x, Δ = Σ_j γ_j r · m_j · G_j in fp32, ONE read-modify-write of dX, pixel sums of Δ and Δ·x̂.
// ================================================================================================
constexpr int kBulkThreads = 320;
constexpr int kBulkWBytes = 18 * kBulkCb * 16;   // one layer's weights for a channel block: [18][96][8] bf16
constexpr int kBulkDzBytes = 2 * kDzRows * 16;   // one layer's halo tile: 2 planes
constexpr size_t kBulkSmem = (size_t)kBlockLayers * (kBulkWBytes + kBulkDzBytes) + kBlockLayers * kBulkCb * sizeof(float2) +
                             kBulkCb * sizeof(float2) + 2 * kBulkCb * sizeof(float);

__device__ __forceinline__ void bulk_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kBulkThreads, 1) dgrad_bulk_kernel(const __grid_constant__ DgradBulkArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sDz = sW + kBlockLayers * kBulkWBytes;
  float2* sK = reinterpret_cast<float2*>(sDz + kBlockLayers * kBulkDzBytes);  // [layer][96]: s = γ·rstd, t = β - mean·s
  float2* sX = sK + kBlockLayers * kBulkCb;                                   // [96]: rstd, -mean·rstd
  float* s_sum = reinterpret_cast<float*>(sX + kBulkCb);                      // [2][96]
  __shared__ uint64_t bar_w, bar_dzfull, bar_dzfree, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int cb_off = blockIdx.y * a.cbs;
  const int ncol = min(a.cbs, a.C0 - cb_off);
  const int n0 = min(kBulkHalf, ncol), n1 = ncol - n0;
  const int nh = n1 > 0 ? 2 : 1;
  const int VH = a.H + 1;
  const int V0 = blockIdx.x * a.vrows_per_cta, V1 = min(V0 + a.vrows_per_cta, a.B * VH);
  const int SW = a.W + 2;  // un-stripped: the region pitch equals the padded plane pitch
  const int t_begin = SW, t_end = (V1 - V0 + 1) * SW;
  const int ntiles = (t_end - t_begin + 127) / 128;
  const int nrows = 128 + 2 * SW + 2;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_dzfull, 1);
    mbar_init(&bar_dzfree, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp of the set
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < a.nl * kBulkCb; i += kBulkThreads) {
    const int j = i / kBulkCb, cl = i - j * kBulkCb;
    float2 k = make_float2(0.f, 0.f), xk = make_float2(0.f, 0.f);
    if (cl < ncol) {
      BnSrc b;
      b.stats = a.stats;
      b.Ct = a.Ct;
      b.inv_n = a.inv_n;
      b.gamma = a.gamma[j];
      b.beta = a.beta[j];
      b.rmean = b.rvar = nullptr;
      b.cs_perm = a.cs_perm;
      b.eval = 0;
      const BnCoef c = bn_coef(b, cb_off + cl);
      k = make_float2(c.s, c.t);
      xk = make_float2(c.rstd, -c.mean * c.rstd);
    }
    sK[i] = k;
    if (j == 0) sX[cl] = xk;
  }
  for (int i = tid; i < 2 * kBulkCb; i += kBulkThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // =============================================================== loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, (uint32_t)a.nl * kBulkWBytes);
      for (int j = 0; j < a.nl; ++j) bulk_g2s(sW + j * kBulkWBytes, a.wpk[j] + (size_t)blockIdx.y * kBulkWBytes, kBulkWBytes, &bar_w);
      const uint32_t plane_bytes = (uint32_t)nrows * 16;
      for (int tile = 0; tile < ntiles; ++tile) {
        if (tile > 0) mbar_wait(&bar_dzfree, (tile - 1) & 1);  // the MMAs that read the previous tile's dz retired
        // smem row rr <-> region position t0 - SW - 1 + rr; the planes' position index is linear in (virtual row, column + 1)
        const long pos0 = a.dz_guard + (long)(V0 - 1) * SW + (t_begin + tile * 128 - SW - 1);
        mbar_arrive_expect_tx(&bar_dzfull, (uint32_t)a.nl * 2 * plane_bytes);
        for (int j = 0; j < a.nl; ++j) {
          bulk_g2s(sDz + j * kBulkDzBytes, a.dz[j] + pos0 * 8, plane_bytes, &bar_dzfull);
          bulk_g2s(sDz + j * kBulkDzBytes + kDzRows * 16, a.dz[j] + (a.dz_npos + pos0) * 8, plane_bytes, &bar_dzfull);
        }
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issue
    if (lane == 0) {
      mbar_wait(&bar_w, 0);
      const uint32_t wa0 = smem_u32(sW), da0 = smem_u32(sDz);
      uint32_t ht = 0;
      for (int tile = 0; tile < ntiles; ++tile) {
        mbar_wait(&bar_dzfull, tile & 1);
        tc_fence_after();
        for (int half = 0; half < nh; ++half, ++ht) {
          const uint32_t buf = ht & 1, u = ht >> 1;
          if (u > 0) {
            mbar_wait(&acc_empty[buf], (u - 1) & 1);  // the epilogue set drained this buffer
            tc_fence_after();
          }
          const uint32_t idesc = umma_idesc_bf16((uint32_t)(half ? n1 : n0), 0, 0);
          for (int j = 0; j < a.nl; ++j) {
            const uint32_t dacc = tmem + buf * 256 + j * kBulkHalf;
            const uint32_t da = da0 + j * kBulkDzBytes, wa = wa0 + j * kBulkWBytes + half * kBulkHalf * 16;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              const uint32_t roff = (uint32_t)((1 - dy) * SW + (1 - dx));
              umma_ss(dacc, desc_kmajor(da + roff * 16, kDzRows), desc_kmajor(wa + tap * 2 * 16 * kBulkCb, kBulkCb), idesc, tap != 0);
            }
          }
          umma_commit(&acc_full[buf]);
        }
        umma_commit(&bar_dzfree);
      }
    }
  } else {
    // =============================================================== epilogue warps 0-7
    const int q = warp & 3, set = warp >> 2;
    float S[3] = {0.f, 0.f, 0.f};
    const int nht = ntiles * nh;
    const long gstride = 16 * a.npx;  // elements between 16-channel group planes
    for (int ht = set; ht < nht; ht += 2) {
      const int tile = nh == 2 ? (ht >> 1) : ht, half = nh == 2 ? (ht & 1) : 0;
      const int ng = (half ? n1 : n0) >> 4;
      const int hoff = half * kBulkHalf;
      long P = -1;  // accumulator row -> pixel
      {
        const int t = t_begin + tile * 128 + q * 32 + lane;
        if (t < t_end) {
          const int rrow = t / SW, cc = t - rrow * SW;
          const int vr = V0 - 1 + rrow, xc = cc - 1;
          const int img = vr / VH, ri = vr - img * VH;
          if (ri < a.H && xc >= 0 && xc < a.W) P = ((long)img * a.H + ri) * a.W + xc;
        }
      }
      const bool valid = P >= 0;
      const size_t e0 = valid ? dact_index(a.npx, P, cb_off + hoff) : 0;
      u32x8 xv[3], dv[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        xv[g].lo = xv[g].hi = dv[g].lo = dv[g].hi = make_uint4(0, 0, 0, 0);
        if (g < ng && valid) {
          xv[g] = ldg256_nc(a.xp + e0 + (size_t)g * gstride);
          if (a.rmw) dv[g] = ld256(a.dact + e0 + (size_t)g * gstride);
        }
      }
      mbar_wait(&acc_full[set], (ht >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + set * 256;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (g < ng) {  // warp-uniform
          float x[16], delta[16];
          unpack8(xv[g].lo, x);
          unpack8(xv[g].hi, x + 8);
#pragma unroll
          for (int i = 0; i < 16; ++i) delta[i] = 0.f;
          for (int j = 0; j < a.nl; ++j) {
            float G[16];
            tmem_ld16(tacc + j * kBulkHalf + 16 * g, G);
            const float2* kk = sK + j * kBulkCb + hoff + 16 * g;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 k = kk[i];
              if (fmaf(x[i], k.x, k.y) > 0.f) delta[i] = fmaf(k.x, G[i], delta[i]);
            }
          }
          if (g == ng - 1) {  // both halves of this buffer's columns have been read: the MMAs two half-tiles ahead may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
          }
          if (valid) {
            float d[16];
            unpack8(dv[g].lo, d);
            unpack8(dv[g].hi, d + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] += delta[i];
            st256(a.dact + e0 + (size_t)g * gstride, pack8(d), pack8(d + 8));
          }
          float gx[16];
          const float2* xk = sX + hoff + 16 * g;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 k = xk[i];
            if (!valid) delta[i] = 0.f;
            gx[i] = delta[i] * fmaf(x[i], k.x, k.y);
          }
          const float r0 = warp_reduce16(delta, lane), r1 = warp_reduce16(gx, lane);
          S[g] += (lane & 1) ? r1 : r0;  // lane keeps quantity (lane & 1) of channel (lane >> 1) & 15 of the group
        }
      }
    }
    // ---- pixel sums of the total: A_c += Σ Δ, B_c += Σ Δ·x̂   (fp32 per CTA -> fp64 atomics)
    const int hoff_set = (nh == 2) ? set * kBulkHalf : 0;
    const int ng_set = ((nh == 2 && set == 1) ? n1 : n0) >> 4;
#pragma unroll
    for (int g = 0; g < 3; ++g)
      if (g < ng_set) atomicAdd(&s_sum[(lane & 1) * kBulkCb + hoff_set + 16 * g + ((lane >> 1) & 15)], S[g]);
    bulk_epi_sync();
    if (tid < ncol) {
      atomicAdd(a.acc + cb_off + tid, (double)s_sum[tid]);
      atomicAdd(a.acc + a.Ct + cb_off + tid, (double)s_sum[kBulkCb + tid]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_dgrad_bulk(const DgradBulkArgs& a, int nruns, int ncb, cudaStream_t st) {
  if (a.W + 2 > 165) return rl_fail(RL_ERR_SHAPE, "dgrad_bulk: level wider than 163 columns needs the per-layer path");
  if (a.nl < 1 || a.nl > kBlockLayers || a.cbs > kBulkCb || (a.cbs & 15)) return rl_fail(RL_ERR_ARG, "dgrad_bulk: bad block geometry");
  {
    ProfScope _ps(KC_DGRAD_BULK, st);
    dgrad_bulk_kernel<<<dim3(nruns, ncb), kBulkThreads, kBulkSmem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ================================================================================================
// wgrad_dense2_kernel — D[row, n] = Σ_p A[p, row] · Bm[p, n],  K = pixels in chunks of 64, N = 144 (9 taps x 16).
// A CTA owns a pixel range and up to three 64-channel blocks of one layer (grid.y = block group, grid.z = layer): rows
// 0-63 of a block are m·x̂ of its channels, rows 64-127 their masks m.  The im2col tile of dz is gathered once per chunk
// and shared by the blocks' MMAs; accumulators: 3 x 144 TMEM columns.
// ================================================================================================
constexpr int kWg2Threads = 256;
constexpr int kWg2MaxMb = 3;
constexpr int kWg2AccStride = 160;
constexpr int kWg2BBytes = 18 * 64 * 16;  // 18432
constexpr size_t kWg2Smem = 3072 + 3 * (size_t)(kWg2BBytes + kWg2MaxMb * 16384);

__global__ void __launch_bounds__(kWg2Threads, 1) wgrad_dense2_kernel(const __grid_constant__ WgDenseArgs a) {
  const WgDenseLayer& L = a.L[blockIdx.z];
  const int nb_total = (L.Cin + 63) / 64;
  const int b0 = blockIdx.y * kWg2MaxMb;
  if (b0 >= nb_total) return;  // uniform per CTA, before any barrier / allocation
  const int nb = min(kWg2MaxMb, nb_total - b0);
  const int cbase = b0 * 64;

  extern __shared__ __align__(128) uint8_t smem[];
  float4* s_k = reinterpret_cast<float4*>(smem);  // [192]: s, t, rstd, -mean·rstd
  uint8_t* stage_base = smem + 3072;
  const int kStageBytes = kWg2BBytes + nb * 16384;  // [B im2col tile | A block 0 | A block 1 | A block 2]
  __shared__ uint64_t bar_done[3];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    mbar_init(&bar_done[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < nb * 64; i += kWg2Threads) {
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cbase + i < L.Cin) {
      BnSrc b;
      b.stats = a.stats;
      b.Ct = a.Ct;
      b.inv_n = a.inv_n;
      b.gamma = L.gamma;
      b.beta = L.beta;
      b.rmean = b.rvar = nullptr;
      b.cs_perm = L.cs_perm;
      b.eval = 0;
      const BnCoef c = bn_coef(b, cbase + i);
      k = make_float4(c.s, c.t, c.rstd, -c.mean * c.rstd);
    }
    s_k[i] = k;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16(144, 1, 1);

  const long P0 = (long)blockIdx.x * a.px_per_cta, P1 = min(P0 + a.px_per_cta, a.npx);
  const int nchunks = P1 > P0 ? (int)((P1 - P0 + 63) / 64) : 0;
  const long HW = (long)a.H * a.W;
  // A tile: thread <-> (pixel row of the chunk, 16-channel group of the 64-channel block): one 32-byte load per block
  const int a_row = warp * 8 + (lane & 7), a_g = lane >> 3;
  // B tile: row = tid & 63, column chunks j = (tid >> 6) + 4q
  const int b_row = tid & 63, b_j0 = tid >> 6;
  struct ARegs {
    u32x8 v[kWg2MaxMb];
    bool valid;
  };
  auto load_a = [&](int kc, ARegs& r) {
    const long P = P0 + (long)kc * 64 + a_row;
    r.valid = kc < nchunks && P < P1;
#pragma unroll
    for (int b = 0; b < kWg2MaxMb; ++b) {
      const int c = cbase + 64 * b + 16 * a_g;
      if (b < nb && r.valid && c < L.Cin) {
        r.v[b] = ldg256_nc(a.xp + dact_index(a.npx, P, c));
      } else {
        r.v[b].lo = r.v[b].hi = make_uint4(0, 0, 0, 0);
      }
    }
  };
  // im2col gather of dz: this thread's pixel row advances by 64 pixels per chunk; its position in the zero-padded planes is
  // carried along.  The tap at (dy,dx) of pixel (b,y,x) sits at pos - dy*(W+2) - dx, always in range.
  long bpos = 0;
  int bx = 0, by = 0;
  long boff[5];
  {
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
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int j = b_j0 + 4 * q;
        if (j < 18) cp_async16(sB + ((size_t)j * 64 + b_row) * 16, L.dz + ((pv ? bpos : (long)a.dz_guard) + boff[q]) * 8, pv);
      }
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
    }
    cp_async_commit();  // always: keeps the group count uniform
  };
  auto step = [&](int k, ARegs& r) {
    const int s = k % 3;
    uint8_t* sB = stage_base + s * kStageBytes;
#pragma unroll
    for (int b = 0; b < kWg2MaxMb; ++b) {
      if (b < nb) {
        uint8_t* sA = sB + kWg2BBytes + b * 16384;
        uint32_t xm[8], mm[8];
        if (r.valid) {
          float x[16];
          unpack8(r.v[b].lo, x);
          unpack8(r.v[b].hi, x + 8);
          const float4* kp = s_k + 64 * b + 16 * a_g;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 k0 = kp[2 * i], k1 = kp[2 * i + 1];
            const bool m0 = fmaf(x[2 * i], k0.x, k0.y) > 0.f, m1 = fmaf(x[2 * i + 1], k1.x, k1.y) > 0.f;
            xm[i] = pack_bf16x2(m0 ? fmaf(x[2 * i], k0.z, k0.w) : 0.f, m1 ? fmaf(x[2 * i + 1], k1.z, k1.w) : 0.f);
            mm[i] = (m0 ? 0x3F80u : 0u) | (m1 ? 0x3F800000u : 0u);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) xm[i] = mm[i] = 0u;
        }
        // chunk layout T[j][row][8], 64 rows: channels 16g .. 16g+15 -> chunks 2g, 2g+1 (m·x̂) and 8+2g, 8+2g+1 (m)
        uint4* d = reinterpret_cast<uint4*>(sA);
        d[(2 * a_g) * 64 + a_row] = make_uint4(xm[0], xm[1], xm[2], xm[3]);
        d[(2 * a_g + 1) * 64 + a_row] = make_uint4(xm[4], xm[5], xm[6], xm[7]);
        d[(8 + 2 * a_g) * 64 + a_row] = make_uint4(mm[0], mm[1], mm[2], mm[3]);
        d[(8 + 2 * a_g + 1) * 64 + a_row] = make_uint4(mm[4], mm[5], mm[6], mm[7]);
      }
    }
    // stage (k+2)%3 == (k-1)%3 is recycled for chunk k+2 once the MMAs of chunk k-1 have retired
    if (k >= 1) mbar_wait(&bar_done[(k - 1) % 3], ((k - 1) / 3) & 1);
    load_a(k + 2, r);
    issue_b(k + 2, stage_base + ((k + 2) % 3) * kStageBytes);
    cp_async_wait<2>();  // B(k) has landed (B(k+1), B(k+2) may still be in flight)
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ba = smem_u32(sB);
      for (int b = 0; b < nb; ++b) {
        const uint32_t aa = ba + kWg2BBytes + b * 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + b * kWg2AccStride, desc_mnmajor(aa + kk * 256, 64), desc_mnmajor(ba + kk * 256, 64), idesc, (k | kk) != 0);
      }
      umma_commit(&bar_done[s]);
    }
  };

  ARegs r0, r1;
  load_a(0, r0);
  issue_b(0, stage_base);
  load_a(1, r1);
  issue_b(1, stage_base + kStageBytes);
  for (int kc = 0; kc < nchunks; kc += 2) {
    step(kc, r0);
    if (kc + 1 < nchunks) step(kc + 1, r1);
  }
  cp_async_wait<0>();
  if (nchunks > 0 && warp < 4) {
    const uint32_t s = (nchunks - 1) % 3, u = (nchunks - 1) / 3;
    mbar_wait(&bar_done[s], u & 1);
    tc_fence_after();
    const int cl = tid & 63, is_mask = tid >> 6;
    for (int b = 0; b < nb; ++b) {
      const int c = cbase + 64 * b + cl;
      const bool valid = c < L.Cin;
      const int rc = valid ? ref_channel(c, L.cs_perm) : 0;
      const float coef = valid ? (is_mask ? L.beta[rc] : L.gamma[rc]) : 0.f;
      // packed weights of channel c: [mb][j = tap*2 + half][cl128][8 co]
      const uint4* wrow = reinterpret_cast<const uint4*>(L.wpk_bwd) + ((size_t)(c >> 7) * 18) * 128 + (c & 127);
      float acc = 0.f;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + b * kWg2AccStride + tap * 16, v);
        if (valid) {
          float w[16];
          unpack8(__ldg(wrow + (size_t)(2 * tap) * 128), w);
          unpack8(__ldg(wrow + (size_t)(2 * tap + 1) * 128), w + 8);
          float* g = L.gw + (long)rc * 9 + tap;  // weight [co][Cin][3][3]
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            acc = fmaf(v[i], w[i], acc);
            atomicAdd(g + (long)i * L.Cin * 9, coef * v[i]);
          }
        }
      }
      if (valid) atomicAdd((is_mask ? L.g_beta : L.g_gamma) + rc, acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_wgrad_dense2(const WgDenseArgs& a, int nruns, int ngroups, cudaStream_t st) {
  if (a.nl < 1 || a.nl > kBlockLayers) return rl_fail(RL_ERR_ARG, "wgrad_dense2: bad layer count");
  {
    ProfScope _ps(KC_WGRAD_DENSE, st);
    wgrad_dense2_kernel<<<dim3(nruns, ngroups, a.nl), kWg2Threads, kWg2Smem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int configure_fused_kernels() {
  RL_CUDA(cudaFuncSetAttribute(dgrad_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBulkSmem));
  RL_CUDA(cudaFuncSetAttribute(wgrad_dense2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWg2Smem));
  return RL_OK;
}

}  // namespace rl
