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
//                         tensor-core operands built from x on the fly, m·x̂ (as m·(x - c), see the kernel) and m
//                         (m = relu mask, exact in bf16):
//                           D2 = (m·x̂)ᵀ·im2col(dz)   D1 = mᵀ·im2col(dz)          (fp32 in TMEM)
//                           dW = γ·D2 + β·D1      dγ = Σ_n W[n,·]·D2[·,n]      dβ = Σ_n W[n,·]·D1[·,n]
//                         so the per-layer pixel sums Σg, Σg·x̂ of BatchNorm backward are contractions on the tensor
//                         pipe and no input-gradient epilogue has to reduce them with shuffles.
//
// The deferred mean terms (A_c, B_c, rl_plan.h) of the block-input channels are the pixel sums of the TOTAL gradient the
// bulk kernel adds, Σ_p Δ = Σ_j γ_j r·dβ_j and Σ_p Δ·x̂ = Σ_j γ_j r·dγ_j: wgrad_dense2_kernel adds them from its dβ_j / dγ_j.
#include <cuda.h>
#include <string.h>

#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

// ================================================================================================
// dgrad_bulk_kernel — one CTA per SM, 10 warps.  Work item: a 2-D tile of the level, TR image rows x (TP - 2) columns
// with TP * TR = 128 (TP = 32 for levels at least 30 wide).  The accumulator rows (M = 128) are the positions of TR rows of
// a TP-wide strip that includes one halo column on each side (2 of TP rows per line are discarded), so that the nine
// taps stay row-shifted views (start-address offsets TP·(1-dy) - dx) of ONE small halo tile per layer:
// (TR + 2) x TP positions, 6 KB, fetched by a single tensor-map TMA box per layer straight from the zero-padded dz
// planes (out-of-range rows / columns arrive as zeros), double-buffered.
//   warp 8    : loader. The five layers' packed weights of this CTA's channel block once (bulk TMA); per tile five
//               cp.async.bulk.tensor.3d boxes [2 planes][TR+2][TP x 8].
//   warp 9    : MMA issue. Per tile and half (<= 48 channels): 5 layers x 9 taps of M = 128, N = half, K = 16 into one of
//               two TMEM buffers (5 x 48 columns each).
//   warps 0-7 : epilogue, set s = warp / 4 drains buffer s (lane quarter warp % 4): relu masks of the five layers from
//               x, Δ = Σ_j γ_j r · m_j · G_j in fp32, ONE read-modify-write of dX.  (The pixel sums Σ Δ, Σ Δ·x̂ that feed
//               the deferred BatchNorm mean terms of these channels equal Σ_j γ_j r·dβ_j, Σ_j γ_j r·dγ_j and are added by
//               wgrad_dense2_kernel, which already holds dβ_j / dγ_j: no cross-pixel reduction here at all.)
// ================================================================================================
constexpr int kBulkThreads = 320;
constexpr int kBulkWBytes = 18 * kBulkCb * 16;   // one layer's weights for a channel block: [18][96][8] bf16
constexpr int kBulkDzBytes = 6400;               // one layer's halo tile: 2 planes x (TR+2)*TP <= 192 positions x 16 B (+ slack)
constexpr size_t kBulkSmem = (size_t)kBlockLayers * kBulkWBytes + 2 * (size_t)kBlockLayers * kBulkDzBytes +
                             kBlockLayers * kBulkCb * sizeof(float2);

struct alignas(64) DgradBulkParams {
  CUtensorMap tmap[kBlockLayers];
  DgradBulkArgs a;
};

__global__ void __launch_bounds__(kBulkThreads, 1) dgrad_bulk_kernel(const __grid_constant__ DgradBulkParams P) {
  const DgradBulkArgs& a = P.a;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sDz = sW + kBlockLayers * kBulkWBytes;                               // [2 sets][layer][kBulkDzBytes]
  float2* sK = reinterpret_cast<float2*>(sDz + 2 * kBlockLayers * kBulkDzBytes);  // [layer][96]: s = γ·rstd, t = β - mean·s
  __shared__ uint64_t bar_w, dz_full[2], dz_free[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int cb_off = blockIdx.y * a.cbs;
  const int ncol = min(a.cbs, a.C0 - cb_off);
  const int n0 = min(kBulkHalf, ncol), n1 = ncol - n0;
  const int nh = n1 > 0 ? 2 : 1;
  const int TP = a.TP, TR = 128 / TP, tp_shift = 31 - __clz(TP);
  const int VH = a.H + 1;
  const long T0 = (long)blockIdx.x * a.tiles_per_cta, T1 = min(T0 + a.tiles_per_cta, a.ntiles);
  const int ntiles = T1 > T0 ? (int)(T1 - T0) : 0;
  const int plane_rows = (TR + 2) * TP;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dz_full[i], 1);
      mbar_init(&dz_free[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp of the set
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < a.nl * kBulkCb; i += kBulkThreads) {
    const int j = i / kBulkCb, cl = i - j * kBulkCb;
    float2 k = make_float2(0.f, 0.f);
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
    }
    sK[i] = k;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // =============================================================== loader
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, (uint32_t)a.nl * kBulkWBytes);
      for (int j = 0; j < a.nl; ++j) bulk_g2s(sW + j * kBulkWBytes, a.wpk[j] + (size_t)blockIdx.y * kBulkWBytes, kBulkWBytes, &bar_w);
      const uint32_t box_bytes = (uint32_t)(2 * plane_rows * 16);
      for (int it = 0; it < ntiles; ++it) {
        const int set = it & 1;
        if (it >= 2) mbar_wait(&dz_free[set], ((it >> 1) - 1) & 1);  // the MMAs that read this set two tiles ago retired
        const long tile = T0 + it;
        const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
        mbar_arrive_expect_tx(&dz_full[set], (uint32_t)a.nl * box_bytes);
        for (int j = 0; j < a.nl; ++j)  // box origin: padded column tc*(TP-2) (= image column - 1), virtual row tr*TR - 1
          tma_load_3d(sDz + (set * kBlockLayers + j) * kBulkDzBytes, &P.tmap[j], tc * (TP - 2) * 8, tr * TR - 1, 0, &dz_full[set]);
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issue.  The issuing thread is a serial instruction
    // stream (90 MMAs per tile): descriptors are formed once and advanced by constants (16-byte units of the address field).
    if (lane == 0) {
      mbar_wait(&bar_w, 0);
      const uint64_t ad0 = desc_kmajor(smem_u32(sDz), (uint32_t)plane_rows), bd0 = desc_kmajor(smem_u32(sW), kBulkCb);
      long roff[9];  // tap -> first operand row: TP·(1-dy) - dx in -1 .. 2·TP+1 (row -1 / plane_rows only feed discarded halo columns)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) roff[tap] = (long)(TP * (1 - (tap / 3 - 1)) - (tap % 3 - 1));
      const uint32_t idesc0 = umma_idesc_bf16((uint32_t)n0, 0, 0), idesc1 = umma_idesc_bf16((uint32_t)max(n1, 16), 0, 0);
      uint32_t ht = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int set = it & 1;
        mbar_wait(&dz_full[set], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t ad_set = ad0 + (uint64_t)(set * (kBlockLayers * kBulkDzBytes >> 4));
        for (int half = 0; half < nh; ++half, ++ht) {
          const uint32_t buf = ht & 1, u = ht >> 1;
          if (u > 0) {
            mbar_wait(&acc_empty[buf], (u - 1) & 1);  // the epilogue set drained this buffer
            tc_fence_after();
          }
          const uint32_t idesc = half ? idesc1 : idesc0;
          // layer-outer, tap-inner (measured: tap-outer / layer-inner, which interleaves five accumulation chains, is 7-60 % slower)
          uint64_t adj = ad_set, bdj = bd0 + (uint64_t)(half * (kBulkHalf * 16 >> 4));
          uint32_t dacc = tmem + buf * 256;
          for (int j = 0; j < a.nl; ++j, adj += (kBulkDzBytes >> 4), bdj += (kBulkWBytes >> 4), dacc += kBulkHalf) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
              umma_ss(dacc, adj + (uint64_t)roff[tap], bdj + (uint64_t)(tap * (2 * 16 * kBulkCb >> 4)), idesc, tap != 0);
          }
          umma_commit(&acc_full[buf]);
        }
        umma_commit(&dz_free[set]);
      }
    }
  } else {
    // =============================================================== epilogue warps 0-7
    const int q = warp & 3, set = warp >> 2;
    const int nht = ntiles * nh;
    const long gstride = 16 * a.npx;  // elements between 16-channel group planes
    const int m = q * 32 + lane, ry = m >> tp_shift, cx = m & (TP - 1);
    // accumulator row -> pixel of half-tile ht (or -1)
    auto pixel_of = [&](int ht) -> long {
      const int it = nh == 2 ? (ht >> 1) : ht;
      const long tile = T0 + it;
      const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
      const int vr = tr * TR + ry, xc = tc * (TP - 2) + cx - 1;
      const int img = vr / VH, ri = vr - img * VH;
      return (cx >= 1 && cx <= TP - 2 && xc < a.W && img < a.B && ri < a.H) ? ((long)img * a.H + ri) * a.W + xc : -1;
    };
    // x / dX of the three 16-channel groups of a half-tile live in the same registers across half-tiles: as soon as a
    // group's values have been unpacked (x) or stored (dX), the next half-tile's are requested into them, so a load has a
    // whole half-tile of work to hide behind
    u32x8 xv[3], dv[3];
    long Pp = -1;
    size_t e0 = 0;
    if (set < nht) {
      Pp = pixel_of(set);
      const int half = nh == 2 ? (set & 1) : 0;
      e0 = Pp >= 0 ? dact_index(a.npx, Pp, cb_off + half * kBulkHalf) : 0;
      const int ng = (half ? n1 : n0) >> 4;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        xv[g].lo = xv[g].hi = dv[g].lo = dv[g].hi = make_uint4(0, 0, 0, 0);
        if (g < ng && Pp >= 0) {
          xv[g] = ldg256_nc(a.xp + e0 + (size_t)g * gstride);
          if (a.rmw) dv[g] = ld256(a.dact + e0 + (size_t)g * gstride);
        }
      }
    }
    for (int ht = set; ht < nht; ht += 2) {
      const int half = nh == 2 ? (ht & 1) : 0;
      const int ng = (half ? n1 : n0) >> 4;
      const int hoff = half * kBulkHalf;
      const bool valid = Pp >= 0;
      const size_t e_cur = e0;
      // next half-tile of this set (same half when the block has two halves)
      const bool more = ht + 2 < nht;
      const long Pn = more ? pixel_of(ht + 2) : -1;
      const size_t e_next = Pn >= 0 ? dact_index(a.npx, Pn, cb_off + hoff) : 0;
      mbar_wait(&acc_full[set], (ht >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + set * 256;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (g < ng) {  // warp-uniform
          float x[16], delta[16];
          unpack8(xv[g].lo, x);
          unpack8(xv[g].hi, x + 8);
          if (Pn >= 0) xv[g] = ldg256_nc(a.xp + e_next + (size_t)g * gstride);
#pragma unroll
          for (int i = 0; i < 16; ++i) delta[i] = 0.f;
          float G[2][16];
          tmem_ld16_nowait(tacc + 16 * g, G[0]);
#pragma unroll
          for (int j = 0; j < kBlockLayers; ++j) {
            if (j < a.nl) {
              tmem_wait_ld();
              reg_fence16(G[j & 1]);
              if (j + 1 < a.nl) tmem_ld16_nowait(tacc + (j + 1) * kBulkHalf + 16 * g, G[(j + 1) & 1]);  // next layer streams in behind
              const float4* kk = reinterpret_cast<const float4*>(sK + j * kBulkCb + hoff + 16 * g);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 k = kk[i];  // (s, t) of channels 2i, 2i+1
                if (fmaf(x[2 * i], k.x, k.y) > 0.f) delta[2 * i] = fmaf(k.x, G[j & 1][2 * i], delta[2 * i]);
                if (fmaf(x[2 * i + 1], k.z, k.w) > 0.f) delta[2 * i + 1] = fmaf(k.z, G[j & 1][2 * i + 1], delta[2 * i + 1]);
              }
            }
          }
          if (g == ng - 1) {  // this buffer's columns have all been read: the MMAs two half-tiles ahead may overwrite it
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
            st256(a.dact + e_cur + (size_t)g * gstride, pack8(d), pack8(d + 8));
          }
          if (a.rmw) {
            if (Pn >= 0) dv[g] = ld256(a.dact + e_next + (size_t)g * gstride);
            else dv[g].lo = dv[g].hi = make_uint4(0, 0, 0, 0);
          }
        }
      }
      Pp = Pn;
      e0 = e_next;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_dgrad_bulk(const DgradBulkArgs& a, const void* const* tmaps, int nruns, int ncb, cudaStream_t st) {
  if (a.nl < 1 || a.nl > kBlockLayers || a.cbs > kBulkCb || (a.cbs & 15)) return rl_fail(RL_ERR_ARG, "dgrad_bulk: bad block geometry");
  if (a.TP != 8 && a.TP != 16 && a.TP != 32) return rl_fail(RL_ERR_ARG, "dgrad_bulk: bad tile pitch");
  DgradBulkParams P;
  for (int j = 0; j < a.nl; ++j) memcpy(&P.tmap[j], tmaps[j], sizeof(CUtensorMap));
  for (int j = a.nl; j < kBlockLayers; ++j) memcpy(&P.tmap[j], tmaps[0], sizeof(CUtensorMap));
  P.a = a;
  {
    ProfScope _ps(KC_DGRAD_BULK, st);
    dgrad_bulk_kernel<<<dim3(nruns, ncb), kBulkThreads, kBulkSmem, st>>>(P);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// host: the two zero-padded planes of one dz buffer as a 3-D tensor [plane][virtual row][(W+2) x 8 bf16]; box =
// [2][TR+2][TP x 8] (rows of up to 512 contiguous bytes; a 16-byte innermost dimension would make every position its own
// TMA request).  plane0 points at position 0 of plane 0 (guard included), guard positions precede virtual row 0.
int encode_dz_tmap(void* tmap_out, const void* plane0, long npos, long guard, int W, long nvrows, int TP) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    RL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (f == nullptr || q != cudaDriverEntryPointSuccess) return rl_fail(RL_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = (EncodeFn)f;
  }
  CUtensorMap tm;
  const int TR = 128 / TP;
  const cuuint64_t dims[3] = {(cuuint64_t)(W + 2) * 8, (cuuint64_t)nvrows, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)(W + 2) * 16, (cuuint64_t)npos * 16};
  const cuuint32_t box[3] = {(cuuint32_t)TP * 8, (cuuint32_t)(TR + 2), 2};
  const cuuint32_t estr[3] = {1, 1, 1};
  void* base = (void*)((const uint8_t*)plane0 + (size_t)guard * 16);
  const CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rl_fail(RL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the dz planes (W=%d)", (int)r, W);
  memcpy(tmap_out, &tm, sizeof(tm));
  return RL_OK;
}

// im2col of the dz planes as ONE tensor-map box.  The tap (dy, dx) of position pos sits at pos - dy (W+2) - dx, so the
// nine taps of 64 consecutive positions are a box over the *overlapping* dimensions
//   (64 positions x 8 channels [contiguous], 3 x-offsets [stride 16 B], 3 y-offsets [stride (W+2) 16 B], 2 planes)
// starting at position pos0 - (W+2) - 1.  It lands in shared memory as [plane][iy][ix][64 rows][8] = the chunk layout
// T[j][row][8] with j = 9 plane + 3 iy + ix, tap = 3 (2 - iy) + (2 - ix): the weight-gradient kernel's B operand, without a
// single address computation.  positions: positions addressable along dimension 1 (the rest of the plane is slack for the
// offset dimensions; out-of-range coordinates are zero-filled).
int encode_dz_im2col_tmap(void* tmap_out, const void* plane0, long npos, long positions, int W) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    RL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (f == nullptr || q != cudaDriverEntryPointSuccess) return rl_fail(RL_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = (EncodeFn)f;
  }
  CUtensorMap tm;
  // the 64 positions x 8 channels of one (plane, iy, ix) are 1024 contiguous bytes: declared as 256 32-bit elements so that
  // the box's inner dimension is one 1 KB row (a 16-byte inner dimension made the copy 1152 separate rows and slower than
  // the cp.async gather it replaced)
  const cuuint64_t dims[4] = {(cuuint64_t)positions * 4, 3, 3, 2};
  const cuuint64_t strides[3] = {16, (cuuint64_t)(W + 2) * 16, (cuuint64_t)npos * 16};
  const cuuint32_t box[4] = {256, 3, 3, 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(plane0), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rl_fail(RL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the dz im2col map (W=%d)", (int)r, W);
  memcpy(tmap_out, &tm, sizeof(tm));
  return RL_OK;
}

__global__ void __launch_bounds__(128, 1) im2col_tmap_probe_kernel(const __grid_constant__ CUtensorMap tm, int c1, uint4* out) {
  __shared__ __align__(128) uint8_t tile[18 * 64 * 16];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 18 * 64 * 16);
    tma_load_4d(tile, &tm, 4 * c1, 0, 0, 0, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 18 * 64; i += 128) out[i] = reinterpret_cast<const uint4*>(tile)[i];
}

// ================================================================================================
// wgrad_dense2_kernel — D[row, n] = Σ_p A[p, row] · Bm[p, n],  K = pixels in chunks of 64, N = 144 (9 taps x 16).
// A CTA owns a pixel range and up to two 64-channel blocks of one layer (grid.y = block group, grid.z = layer): rows
// 0-63 of a block are m·(x - c) of its channels (c = the channel mean rounded to bf16, so the operand is centred and as
// accurate as m·x̂), rows 64-127 their masks m.  Both are formed from packed bf16 pairs in five instructions per pair:
// the relu decision fma(x, s, t) > 0 of a channel is monotone in x, so it equals a comparison of ±x with a per-channel
// bf16 threshold (relu_threshold), exact for every bf16 x.  With D2'' and D1 the two accumulators of a channel,
//   D2 = r·D2'' - r·(mean - c)·D1          (x̂ = r·(x - c) - r·(mean - c))
// and dW, dγ, dβ follow as in the file header.  K runs over the positions of the zero-padded dz planes, so the im2col tile
// of a 64-position chunk is ONE TMA box (encode_dz_im2col_tmap above) shared by the blocks' MMAs, and the A rows of pad
// positions are zero (4-deep rings); accumulators: 2 x 144 TMEM columns.  18 warps: 16 A-operand producers, one MMA-issuing
// warp, one TMA warp; mbarriers only, no block-wide barrier in the main loop.  (The tile used to be gathered with 1152
// cp.async per chunk: 38 % of the kernel's shared-memory wavefronts and half of the producers' instructions.)
// ================================================================================================
constexpr int kWg2Producers = 512;              // warps 0-15: A-operand producers (warps 0-7 also read the accumulators out)
constexpr int kWg2Threads = kWg2Producers + 64;  // warp 16: MMA issue, warp 17: the im2col TMA
constexpr int kWg2MaxMb = 2;
constexpr int kWg2AccStride = 160;
constexpr int kWg2BBytes = 18 * 64 * 16;  // 18432
constexpr int kWg2BStages = 4, kWg2AStages = 4;
constexpr int kWg2ConstBytes = kWg2MaxMb * 4 * 24 * 4;  // per (block, 16-channel group): 8 x (c2, sgn2, thr2)
constexpr size_t kWg2Smem = kWg2ConstBytes + (size_t)kWg2BStages * kWg2BBytes + (size_t)kWg2AStages * kWg2MaxMb * 16384;

// monotone integer key of a bf16 bit pattern (-inf < ... < -0 < +0 < ... < +inf) and back
__device__ __forceinline__ uint32_t bf16_key(uint32_t b) { return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u); }
__device__ __forceinline__ uint32_t bf16_unkey(uint32_t k) { return (k & 0x8000u) ? (k & 0x7FFFu) : (~k & 0xFFFFu); }

// For one channel (s, t of its BatchNorm): sign σ = ±1 and the bf16 threshold θ such that for EVERY bf16 value x
//   fmaf(x, s, t) > 0   <=>   σ·x >= θ.
// fmaf is monotone in x, so {x : fmaf(x,s,t) > 0} is an up-set of σ·x on the bf16 grid; θ is its smallest element,
// found by stepping from the rounded real threshold -t/s (at most a couple of bf16 steps away).
__device__ __forceinline__ void relu_threshold(float s, float t, uint32_t* sgn_bits, uint32_t* thr_bits) {
  constexpr uint32_t kMin = 0x007Fu, kMax = 0xFF80u;  // keys of -inf, +inf
  if (s == 0.f) {
    *sgn_bits = 0x3F80u;
    *thr_bits = (t > 0.f) ? 0xFF80u : 0x7F80u;  // always / never
    return;
  }
  const bool up = s > 0.f;
  const float ty = up ? (-t / s) : (t / s);
  __nv_bfloat16 r = __float2bfloat16_rn(ty);
  uint32_t key = bf16_key((uint32_t)__bfloat16_as_ushort(r));
  key = min(max(key, kMin), kMax);
  auto f = [&](uint32_t k) {
    const float y = __uint_as_float(bf16_unkey(k) << 16);
    return fmaf(up ? y : -y, s, t) > 0.f;
  };
  for (int it = 0; it < 16 && key > kMin && f(key - 1); ++it) --key;
  for (int it = 0; it < 16 && key < kMax && !f(key); ++it) ++key;
  *sgn_bits = up ? 0x3F80u : 0xBF80u;
  *thr_bits = bf16_unkey(key);
}

struct alignas(64) WgDenseParams {
  CUtensorMap tm[kBlockLayers];  // im2col maps of the layers' dz planes (encode_dz_im2col_tmap)
  WgDenseArgs a;
};

__global__ void __launch_bounds__(kWg2Threads, 1) wgrad_dense2_kernel(const __grid_constant__ WgDenseParams P) {
  const WgDenseArgs& a = P.a;
  const WgDenseLayer& L = a.L[blockIdx.z];
  const int nb_total = (L.Cin + 63) / 64;
  const int ngroups = (nb_total + kWg2MaxMb - 1) / kWg2MaxMb;
  if ((int)blockIdx.y >= ngroups) return;  // uniform per CTA, before any barrier / allocation
  const int bpg = (nb_total + ngroups - 1) / ngroups;
  const int b0 = blockIdx.y * bpg;
  const int nb = min(bpg, nb_total - b0);
  const int cbase = b0 * 64;

  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* s_k = reinterpret_cast<uint32_t*>(smem);  // [block][group16][c2 x8 | sgn2 x8 | thr2 x8]
  uint8_t* sBring = smem + kWg2ConstBytes;
  uint8_t* sAring = sBring + kWg2BStages * kWg2BBytes;
  constexpr int kAStage = kWg2MaxMb * 16384;
  // full[k % 4]: the producers stored + fenced their part of chunk k's operands;  done[k % 4]: its MMAs retired
  __shared__ uint64_t bar_full[kWg2BStages], bar_done[kWg2BStages];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  BnSrc bn;
  bn.stats = a.stats;
  bn.Ct = a.Ct;
  bn.inv_n = a.inv_n;
  bn.gamma = L.gamma;
  bn.beta = L.beta;
  bn.rmean = bn.rvar = nullptr;
  bn.cs_perm = L.cs_perm;
  bn.eval = 0;

  if (tid == 0) {
    for (int i = 0; i < kWg2BStages; ++i) {
      mbar_init(&bar_full[i], 4 + 1);  // the four warps of the stage's producer group + the TMA thread's expect-tx arrival
      mbar_init(&bar_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < nb * 64; i += kWg2Threads) {
    uint32_t cb = 0, sg = 0x3F80u, th = 0x7F80u;  // channel beyond Cin: mask never set
    if (cbase + i < L.Cin) {
      const BnCoef c = bn_coef(bn, cbase + i);
      cb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(c.mean));
      relu_threshold(c.s, c.t, &sg, &th);
    }
    // 16-bit halves of the pair words: channel i -> block i/64, group (i%64)/16, pair (i%16)/2, half i&1
    uint16_t* base = reinterpret_cast<uint16_t*>(s_k + ((i >> 6) * 4 + ((i & 63) >> 4)) * 24);
    const int h = (i & 15);
    base[h] = (uint16_t)cb;
    base[16 + h] = (uint16_t)sg;
    base[32 + h] = (uint16_t)th;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16(144, 1, 1);

  // K runs over the POSITIONS of the zero-padded dz planes (vrow * (W+2) + col; col 0 / W+1 and row H of every image are
  // pads, 2 % of the positions at 120x160): the im2col tile of 64 consecutive positions is then one TMA box, and the A rows
  // of pad positions are zero
  const int PW = a.W + 2;
  const long nposK = (long)a.B * (a.H + 1) * PW;
  const long Q0 = (long)blockIdx.x * a.pos_per_cta, Q1 = min(Q0 + a.pos_per_cta, nposK);
  const int nchunks = Q1 > Q0 ? (int)((Q1 - Q0 + 63) / 64) : 0;

  if (warp == 16) {
    // =============================================================== MMA issue: no block-wide barrier anywhere in the main
    // loop; the producers run ahead of the tensor pipe by the depth of the rings
    if (lane == 0) {
      // descriptors are formed once and advanced by constants (16-byte units of the address field): the issuing thread is
      // a serial instruction stream and must stay far below the 72 cycles an MMA takes
      const uint64_t bd0 = desc_mnmajor(smem_u32(sBring), 64), ad0 = desc_mnmajor(smem_u32(sAring), 64);
      uint32_t sb = 0, pb = 0;
      for (int k = 0; k < nchunks; ++k) {
        mbar_wait(&bar_full[sb], pb);
        tc_fence_after();
        const uint64_t bd = bd0 + (uint64_t)(sb * (kWg2BBytes >> 4));
        uint64_t ad = ad0 + (uint64_t)(sb * (kAStage >> 4));
        uint32_t dacc = tmem;
        for (int b = 0; b < nb; ++b, ad += (16384 >> 4), dacc += kWg2AccStride) {
          umma_ss(dacc, ad, bd, idesc, k != 0);
          umma_ss(dacc, ad + 16, bd + 16, idesc, 1);
          umma_ss(dacc, ad + 32, bd + 32, idesc, 1);
          umma_ss(dacc, ad + 48, bd + 48, idesc, 1);
        }
        umma_commit(&bar_done[sb]);
        if (++sb == kWg2BStages) {
          sb = 0;
          pb ^= 1;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    return;
  }
  if (warp == 17) {
    // =============================================================== im2col of dz: one TMA box per chunk
    if (lane == 0) {
      const long c1 = a.dz_guard + Q0 - PW - 1;  // tap (dy,dx) of position q sits at q - dy*PW - dx: box offsets iy = 1 - dy, ix = 1 - dx
      for (int k = 0; k < nchunks; ++k) {
        const uint32_t s = (uint32_t)k & 3;
        if (k >= 4) mbar_wait(&bar_done[s], ((k - 4) >> 2) & 1);  // the MMAs that read this stage have retired
        mbar_arrive_expect_tx(&bar_full[s], kWg2BBytes);
        tma_load_4d(sBring + s * kWg2BBytes, &P.tm[blockIdx.z], (int)(4 * (c1 + 64L * k)), 0, 0, 0, &bar_full[s]);
      }
    }
    tc_fence_before();
    __syncthreads();
    return;
  }

  // ================================================================= producers, warps 0-15, in four groups of four warps.
  // Group g produces the A operands of the chunks k = g, g+4, ... into ring stage g.  fence.proxy.async, which every
  // published stage needs, waits for the loads its thread has outstanding: a producer that prefetches its next chunk before
  // publishing the current one runs at one DRAM latency per chunk (that was this kernel's pace: ~1600 cycles per chunk).  Here a
  // thread issues the loads of its next chunk right AFTER the fence and meets them again three chunks of other groups' work
  // later.
  // thread <-> (64-channel block, 16-channel group) fixed, four pixel rows of the chunk: one 32-byte load per row
  const int grp = warp >> 2, tg = tid & 127;
  const int a_blk = tg >> 6, a_g = (tg >> 4) & 3, a_r0 = tg & 15;
  const int a_c = cbase + 64 * a_blk + 16 * a_g;
  const bool a_on = a_blk < nb && a_c < L.Cin;
  // carried position state of each of the thread's rows: column inside the padded plane row, image row (H = the pad row),
  // pixel index of that row's first pixel ((b*H + y)*W; the pad row takes the next image's first row, so a wrap leaves it
  // unchanged)
  long a_q[4], a_pxrow[4];
  int a_col[4], a_y[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a_q[i] = Q0 + 64L * grp + a_r0 + 16 * i;
    const long vrow = a_q[i] / PW;
    a_col[i] = (int)(a_q[i] - vrow * PW);
    const long b = vrow / (a.H + 1);
    a_y[i] = (int)(vrow - b * (a.H + 1));
    a_pxrow[i] = (b * a.H + a_y[i]) * a.W;
  }
  const __nv_bfloat16* a_base = a.xp + (size_t)((a_on ? a_c : 0) >> 4) * (size_t)a.npx * 16;
  // this thread's pair constants stay in registers for the whole kernel
  uint32_t kc2[8], kg2[8], kt2[8];
  {
    const uint4* kp = reinterpret_cast<const uint4*>(s_k + ((a_on ? a_blk : 0) * 4 + a_g) * 24);
    const uint4 c0 = kp[0], c1 = kp[1], g0 = kp[2], g1 = kp[3], t0 = kp[4], t1 = kp[5];
    kc2[0] = c0.x; kc2[1] = c0.y; kc2[2] = c0.z; kc2[3] = c0.w; kc2[4] = c1.x; kc2[5] = c1.y; kc2[6] = c1.z; kc2[7] = c1.w;
    kg2[0] = g0.x; kg2[1] = g0.y; kg2[2] = g0.z; kg2[3] = g0.w; kg2[4] = g1.x; kg2[5] = g1.y; kg2[6] = g1.z; kg2[7] = g1.w;
    kt2[0] = t0.x; kt2[1] = t0.y; kt2[2] = t0.z; kt2[3] = t0.w; kt2[4] = t1.x; kt2[5] = t1.y; kt2[6] = t1.z; kt2[7] = t1.w;
  }
  u32x8 av[4];
  bool avalid[4];
  auto load_a = [&](int kc) {  // this group's chunk kc; advances the rows by four chunks
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      avalid[i] = kc < nchunks && a_q[i] < Q1 && a_col[i] >= 1 && a_col[i] <= a.W && a_y[i] < a.H;
      if (a_on && avalid[i]) av[i] = ldg256_nc(a_base + (size_t)(a_pxrow[i] + a_col[i] - 1) * 16);
      else av[i].lo = av[i].hi = make_uint4(0, 0, 0, 0);
      a_q[i] += 256;
      a_col[i] += 256;
      while (a_col[i] >= PW) {
        a_col[i] -= PW;
        if (a_y[i] == a.H) {
          a_y[i] = 0;
        } else {
          ++a_y[i];
          a_pxrow[i] += a.W;
        }
      }
    }
  };
  const uint32_t sA_thr = smem_u32(sAring) + (uint32_t)(grp * kAStage + a_blk * 16384 + ((2 * a_g) * 64 + a_r0) * 16);
  load_a(grp);
  for (int k = grp, u = 0; k < nchunks; k += 4, ++u) {
    if (u > 0) mbar_wait(&bar_done[grp], (u - 1) & 1);  // the MMAs of chunk k-4 (this stage's previous tenant) have retired
    if (a_blk < nb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t xm[8], mm[8];
        const uint32_t xs[8] = {av[i].lo.x, av[i].lo.y, av[i].lo.z, av[i].lo.w, av[i].hi.x, av[i].hi.y, av[i].hi.z, av[i].hi.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&xs[e]);
          const __nv_bfloat162 cv = *reinterpret_cast<const __nv_bfloat162*>(&kc2[e]);
          const __nv_bfloat162 gv = *reinterpret_cast<const __nv_bfloat162*>(&kg2[e]);
          const __nv_bfloat162 tv = *reinterpret_cast<const __nv_bfloat162*>(&kt2[e]);
          uint32_t m = __hge2_mask(__hmul2(xv, gv), tv);  // 0xFFFF per channel whose relu is open
          if (!(a_on && avalid[i])) m = 0u;
          const __nv_bfloat162 d = __hsub2(xv, cv);
          xm[e] = *reinterpret_cast<const uint32_t*>(&d) & m;
          mm[e] = m & 0x3F803F80u;
        }
        // chunk layout T[j][row][8], 64 rows: channels 16g .. 16g+15 -> chunks 2g, 2g+1 (m·(x-c)) and 8+2g, 8+2g+1 (m)
        const uint32_t d = sA_thr + i * 16 * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d), "r"(xm[0]), "r"(xm[1]), "r"(xm[2]), "r"(xm[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d + 64 * 16), "r"(xm[4]), "r"(xm[5]), "r"(xm[6]), "r"(xm[7]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d + 8 * 64 * 16), "r"(mm[0]), "r"(mm[1]), "r"(mm[2]), "r"(mm[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(d + 9 * 64 * 16), "r"(mm[4]), "r"(mm[5]), "r"(mm[6]), "r"(mm[7]) : "memory");
      }
    }
    fence_proxy_async();  // nothing of this thread is in flight here: the next chunk's loads are issued below
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_full[grp]);
    load_a(k + 4);
  }
  // The groups finish at different times.  A parity wait is only meaningful within one phase of its barrier: a warp that asked
  // for the last chunk's phase while that stage's barrier was still two phases behind would be let through at once.  After
  // this rendezvous every chunk has been published, so each stage's barrier is in (or past) its final phase.
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (nchunks > 0 && warp < 8 && (warp >> 2) < nb) {
    const uint32_t s = (nchunks - 1) % kWg2BStages, u = (nchunks - 1) / kWg2BStages;
    mbar_wait(&bar_done[s], u & 1);
    tc_fence_after();
    // warps 0-3 read block 0's accumulator, warps 4-7 block 1's (TMEM lane quarter = warp % 4)
    const int b = warp >> 2, row = tid & 127, cl = row & 63, is_mask = row >> 6;
    const int c = cbase + 64 * b + cl;
    const bool valid = c < L.Cin;
    int rc = 0;
    float cw = 0.f, cg = 0.f, gr = 0.f;  // this row's weight in dW and in dγ (dβ: the mask row itself); γ·rstd
    if (valid) {
      rc = ref_channel(c, L.cs_perm);
      const BnCoef k = bn_coef(bn, c);
      const float resid = k.mean - __bfloat162float(__float2bfloat16_rn(k.mean));  // mean - c
      gr = k.gamma * k.rstd;
      cw = is_mask ? (L.beta[rc] - gr * resid) : gr;
      cg = is_mask ? (-k.rstd * resid) : k.rstd;
    }
    // packed weights of channel c: [mb][j = tap*2 + half][cl128][8 co]
    const uint4* wrow = reinterpret_cast<const uint4*>(L.wpk_bwd) + ((size_t)(c >> 7) * 18) * 128 + (c & 127);
    float acc = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      // accumulator column of (tap, co): the TMA tile's chunk order is j = 9 * (co / 8) + 8 - tap (plane, y offset, x offset)
      float v[16];
      tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + b * kWg2AccStride + (8 - tap) * 8, v);
      tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + b * kWg2AccStride + (17 - tap) * 8, v + 8);
      if (valid) {
        float w[16];
        unpack8(__ldg(wrow + (size_t)(2 * tap) * 128), w);
        unpack8(__ldg(wrow + (size_t)(2 * tap + 1) * 128), w + 8);
        float* g = L.gw + (long)rc * 9 + tap;  // weight [co][Cin][3][3]
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          acc = fmaf(v[i], w[i], acc);
          atomicAdd(g + (long)i * L.Cin * 9, cw * v[i]);
        }
      }
    }
    if (valid) {
      atomicAdd(L.g_gamma + rc, cg * acc);
      if (is_mask) atomicAdd(L.g_beta + rc, acc);
      // deferred mean terms of the block-input channels (the chain launches cover the block's own new channels):
      // A_c += γ r · dβ_j[c],  B_c += γ r · dγ_j[c]   (= Σ_p, Σ_p x̂· of what dgrad_bulk adds for this layer)
      if (c < a.C0) {
        atomicAdd(a.acc + a.Ct + c, (double)(gr * cg * acc));
        if (is_mask) atomicAdd(a.acc + c, (double)(gr * acc));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_wgrad_dense2(const WgDenseArgs& a, const void* const* tmaps, int nruns, int ngroups, cudaStream_t st) {
  if (a.nl < 1 || a.nl > kBlockLayers) return rl_fail(RL_ERR_ARG, "wgrad_dense2: bad layer count");
  WgDenseParams P;
  P.a = a;
  for (int j = 0; j < a.nl; ++j) memcpy(&P.tm[j], tmaps[j], sizeof(CUtensorMap));
  for (int j = a.nl; j < kBlockLayers; ++j) memcpy(&P.tm[j], tmaps[0], sizeof(CUtensorMap));
  {
    ProfScope _ps(KC_WGRAD_DENSE, st);
    wgrad_dense2_kernel<<<dim3(nruns, ngroups, a.nl), kWg2Threads, kWg2Smem, st>>>(P);
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

extern "C" int rl_selftest_im2col_tmap(const void* planes, long npos, long positions, int W, int c1, void* out, void* stream) {
  using namespace rl;
  if (!planes || !out) return rl_fail(RL_ERR_ARG, "rl_selftest_im2col_tmap: null argument");
  CUtensorMap tm;
  int rc = encode_dz_im2col_tmap(&tm, planes, npos, positions, W);
  if (rc) return rc;
  im2col_tmap_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(tm, c1, (uint4*)out);
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}
