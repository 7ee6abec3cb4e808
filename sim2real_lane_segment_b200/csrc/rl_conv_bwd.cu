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

constexpr int kDzRows = 464;  // halo tile rows: 128 + 2*(strip+2) + 2 <= 464 for strips up to 165 columns

// ------------------------------------------------------------------------------------------------
// epilogue shared by both dgrad kernels: thread = channel, 128 accumulator columns = pixels
// ------------------------------------------------------------------------------------------------
struct DgradChan {
  bool active;
  int c;
  float s, t, mean, rstd, gr;  // gr = gamma * rstd
  float S1, S2;
};

__device__ __forceinline__ void dgrad_epilogue(uint32_t tmem_acc, int warp, const int* P_s, const DgradArgs& a,
                                               DgradChan& ch) {
  const unsigned short* actu = reinterpret_cast<const unsigned short*>(a.act);
  unsigned short* dactu = reinterpret_cast<unsigned short*>(a.dact);
  const bool rmw = ch.c >= a.rmw_from;
#pragma unroll 1
  for (int n0 = 0; n0 < 128; n0 += 16) {
    float g[16];
    tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16) + n0, g);
    if (ch.active) {
      unsigned short xs[16], ds[16];
      long idx[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int P = P_s[n0 + i];
        idx[i] = P >= 0 ? (long)P * a.Ct + ch.c : -1;
        xs[i] = idx[i] >= 0 ? __ldg(actu + idx[i]) : (unsigned short)0;
        ds[i] = (idx[i] >= 0 && rmw) ? dactu[idx[i]] : (unsigned short)0;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (idx[i] >= 0) {
          const float x = __uint_as_float((uint32_t)xs[i] << 16);
          const float gi = (fmaf(x, ch.s, ch.t) > 0.f) ? g[i] : 0.f;
          const float xh = (x - ch.mean) * ch.rstd;
          ch.S1 += gi;
          ch.S2 = fmaf(gi, xh, ch.S2);
          const float old = __uint_as_float((uint32_t)ds[i] << 16);
          dactu[idx[i]] = __bfloat16_as_ushort(__float2bfloat16_rn(fmaf(ch.gr, gi, old)));
        }
      }
    }
  }
}

__device__ __forceinline__ void dgrad_chan_init(DgradChan& ch, const DgradArgs& a, int mb) {
  ch.c = mb * 128 + threadIdx.x;
  ch.active = ch.c < a.Cin;
  ch.S1 = ch.S2 = 0.f;
  if (ch.active) {
    BnCoef k = bn_coef(a.bn, ch.c);
    ch.s = k.s;
    ch.t = k.t;
    ch.mean = k.mean;
    ch.rstd = k.rstd;
    ch.gr = k.gamma * k.rstd;
  } else {
    ch.s = ch.t = ch.mean = ch.rstd = ch.gr = 0.f;
  }
}

__device__ __forceinline__ void dgrad_chan_flush(const DgradChan& ch, const DgradArgs& a) {
  if (!ch.active) return;
  const int rc = ref_channel(ch.c, a.bn.cs_perm);
  atomicAdd(a.g_beta + rc, ch.S1);
  atomicAdd(a.g_gamma + rc, ch.S2);
  atomicAdd(a.acc + ch.c, (double)(ch.gr * ch.S1));
  atomicAdd(a.acc + a.Ct + ch.c, (double)(ch.gr * ch.S2));
}

// ================================================================================================
// dgrad_dense_kernel
// ================================================================================================
__global__ void __launch_bounds__(128, 3) dgrad_dense_kernel(DgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;                 // W^T of one M-block: [18][128][8] bf16 = 36864 B
  uint8_t* sDz = smem + 36864;        // halo tile [2][kDzRows][8] bf16 = 14848 B
  int* P_s = reinterpret_cast<int*>(smem + 36864 + 2 * kDzRows * 16);
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  const int VH = a.H + 1;
  const int vrows_total = a.B * VH;
  const int run = blockIdx.x / a.nstrips, strip = blockIdx.x % a.nstrips;
  const int V0 = run * a.vrows_per_cta, V1 = min(V0 + a.vrows_per_cta, vrows_total);
  const int x0 = strip * a.strip_w, x1 = min(x0 + a.strip_w, a.W);
  const int SW = x1 - x0 + 2;
  const int t_begin = SW, t_end = (V1 - V0 + 1) * SW;  // rows V0..V1-1 of the region (region row 0 is V0-1)
  const int ntiles = (t_end - t_begin + 127) / 128;
  const int nmb = (a.Cin + 127) / 128;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16(128, 0, 0);
  uint32_t n_w = 0, n_mma = 0;

  for (int mb = 0; mb < nmb; ++mb) {
    DgradChan ch;
    dgrad_chan_init(ch, a, mb);
    __syncthreads();  // every MMA of the previous M-block has been waited for: sW may be overwritten
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_w, 36864);
      bulk_g2s(sW, a.wpk + (size_t)mb * 36864, 36864, &bar_w);
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = t_begin + tile * 128;
      // ---- halo tile of dz: smem row rr <-> region position t0 - SW - 1 + rr
      const int nrows = 128 + 2 * SW + 2;
      for (int rr = tid; rr < nrows; rr += 128) {
        const int t = t0 - SW - 1 + rr;
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
        if (t >= 0) {
          const int rrow = t / SW, cc = t - rrow * SW;
          const int vr = V0 - 1 + rrow, xc = x0 - 1 + cc;
          if (vr >= 0 && vr < vrows_total && xc >= 0 && xc < a.W) {
            const int img = vr / VH, ri = vr - img * VH;
            if (ri < a.H) {
              const __nv_bfloat16* src = a.dz + (((size_t)img * a.H + ri) * a.W + xc) * 16;
              v0 = ldg128(src);
              v1 = ldg128(src + 8);
            }
          }
        }
        *reinterpret_cast<uint4*>(sDz + (size_t)rr * 16) = v0;
        *reinterpret_cast<uint4*>(sDz + ((size_t)kDzRows + rr) * 16) = v1;
      }
      {  // accumulator column -> pixel
        const int t = t0 + tid;
        int P = -1;
        if (t < t_end) {
          const int rrow = t / SW, cc = t - rrow * SW;
          const int vr = V0 - 1 + rrow, xc = x0 - 1 + cc;
          const int img = vr / VH, ri = vr - img * VH;
          if (ri < a.H && xc >= x0 && xc < x1) P = (img * a.H + ri) * a.W + xc;
        }
        P_s[tid] = P;
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        if (tile == 0) {
          mbar_wait(&bar_w, n_w & 1);
        }
        tc_fence_after();
        const uint32_t wa = smem_u32(sW), da = smem_u32(sDz);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const uint32_t roff = (uint32_t)((1 - dy) * SW + (1 - dx));
          umma_ss(tmem, desc_kmajor(wa + tap * 2 * 16 * 128, 128), desc_kmajor(da + roff * 16, kDzRows), idesc, tap != 0);
        }
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, n_mma & 1);
      ++n_mma;
      tc_fence_after();
      dgrad_epilogue(tmem, warp, P_s, a, ch);
      tc_fence_before();
      __syncthreads();
    }
    if (ntiles > 0) ++n_w;
    else if (tid == 0) { mbar_wait(&bar_w, n_w & 1); ++n_w; }
    dgrad_chan_flush(ch, a);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ================================================================================================
// dgrad_td_kernel
// ================================================================================================
__global__ void __launch_bounds__(128, 3) dgrad_td_kernel(DgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stage_base = smem;  // 2 x (W^T chunk 16 KB + dz chunk 16 KB)
  int* P_s = reinterpret_cast<int*>(smem + 2 * 32768);
  __shared__ uint64_t bar_full[2], bar_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
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
  const uint32_t idesc = umma_idesc_bf16(128, 0, 0);
  uint32_t g = 0;
  TileLoader<128> ld;
  const __nv_bfloat16* rowptr[4];

  for (int mb = 0; mb < nmb; ++mb) {
    DgradChan ch;
    dgrad_chan_init(ch, a, mb);
    for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long P = tile * 128 + TileLoader<128>::row_of(i);
        rowptr[i] = P < npx ? a.dz + (size_t)P * a.Kdz : nullptr;
      }
      {
        const long P = tile * 128 + tid;
        P_s[tid] = P < npx ? (int)P : -1;
      }
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
        if (kc + 1 < nkc) {
          const __nv_bfloat16* rp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] = rowptr[i] ? rowptr[i] + (kc + 1) * 64 : nullptr;
          ld.load(rp, min(8, (a.Kdz - (kc + 1) * 64) / 8));
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
          mbar_wait(&bar_full[s], u & 1);
          tc_fence_after();
          const uint32_t wa = smem_u32(sW), ba = smem_u32(sB);
          for (int kk = 0; kk < ck / 16; ++kk)
            umma_ss(tmem, desc_kmajor(wa + kk * 2 * 16 * 128, 128), desc_kmajor(ba + kk * 2 * 16 * 128, 128), idesc,
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
      dgrad_epilogue(tmem, warp, P_s, a, ch);
      tc_fence_before();
      __syncthreads();
    }
    dgrad_chan_flush(ch, a);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int launch_dgrad(const DgradArgs& a, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    RL_CUDA(cudaFuncSetAttribute(dgrad_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    RL_CUDA(cudaFuncSetAttribute(dgrad_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    configured = true;
  }
  if (a.dense) {
    if (a.strip_w + 2 > 165) return rl_fail(RL_ERR_SHAPE, "dgrad: strip too wide");
    size_t smem = 36864 + 2 * kDzRows * 16 + 512;
    { ProfScope _ps(KC_DGRAD_DENSE, st);
  dgrad_dense_kernel<<<grid, 128, smem, st>>>(a); }
  } else {
    size_t smem = 2 * 32768 + 512;
    { ProfScope _ps(KC_DGRAD_TD, st);
  dgrad_td_kernel<<<grid, 128, smem, st>>>(a); }
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
    case WG_TU: return ((long)c * kUpCh + ng % kUpCh) * 9 + ng / kUpCh;
    case WG_HEAD: return ng < a.num_cls ? (long)ng * 288 + ref_channel(c, 128) : -1;
    default: return (long)ng * 27 + c;  // WG_FIRST
  }
}

__global__ void __launch_bounds__(128, 2) wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int mb = blockIdx.y, part = blockIdx.z;
  const int n0 = a.part_n0[part], np = a.part_np[part];
  const int ca = min(128, a.Ca - mb * 128);  // valid channels of this M-block
  float* s_s = reinterpret_cast<float*>(smem);
  float* s_t = s_s + 128;
  uint8_t* stage_base = smem + 1024;
  const int stageB = np * 128;  // np/8 chunks x 64 rows x 16 B
  const int kStageBytes = 16384 + stageB;
  __shared__ uint64_t bar_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  {
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
  TileLoader<64> ld;
  const __nv_bfloat16* rowptr[2];
  const long HW = (long)a.H * a.W;

  for (int kc = 0; kc < nchunks; ++kc) {
    const long Pc = P0 + (long)kc * 64;
    const uint32_t s = kc & 1, u = kc >> 1;
    uint8_t* sA = stage_base + s * kStageBytes;
    uint8_t* sB = sA + 16384;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long P = Pc + TileLoader<64>::row_of(i);
      rowptr[i] = P < P1 ? a.A + (size_t)P * a.lda + a.a_off + mb * 128 : nullptr;
    }
    ld.load(rowptr, a_nchunk);
    if (u > 0) mbar_wait(&bar_done[s], (u - 1) & 1);
    ld.store(sA, rowptr, a_nchunk, a.transform ? s_s : nullptr, s_t);
    // zero-fill the channel chunks beyond Ca once per stage use is unnecessary: those accumulator rows are ignored
    // ---- B tile
    if (a.kind == WG_DENSE) {
      for (int it = tid; it < 64 * 18; it += 128) {
        const int row = it & 63, j = it >> 6;
        const int tap = j >> 1, half = j & 1;
        const long P = Pc + row;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (P < P1) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const long b = P / HW;
          const int hw = (int)(P - b * HW), y = hw / a.W, x = hw - y * a.W;
          const int qy = y - dy, qx = x - dx;
          if (qy >= 0 && qy < a.H && qx >= 0 && qx < a.W) v = ldg128(a.Bm + ((b * a.H + qy) * a.W + qx) * 16 + half * 8);
        }
        *reinterpret_cast<uint4*>(sB + ((size_t)j * 64 + row) * 16) = v;
      }
    } else {
      const int nj = np / 8;
      for (int it = tid; it < 64 * nj; it += 128) {
        const int row = it & 63, j = it >> 6;
        const long P = Pc + row;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (P < P1) v = ldg128(a.Bm + (size_t)P * a.ldb + n0 + j * 8);
        *reinterpret_cast<uint4*>(sB + ((size_t)j * 64 + row) * 16) = v;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(sA), ba = smem_u32(sB);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem, desc_mnmajor(aa + kk * 256, 64), desc_mnmajor(ba + kk * 256, 64), idesc, (kc | kk) != 0);
      umma_commit(&bar_done[s]);
    }
  }
  if (nchunks > 0) {
    const uint32_t s = (nchunks - 1) & 1, u = (nchunks - 1) >> 1;
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
  static bool configured = false;
  if (!configured) {
    RL_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  int npmax = 0;
  for (int i = 0; i < nparts; ++i) npmax = a.part_np[i] > npmax ? a.part_np[i] : npmax;
  size_t smem = 1024 + 2 * (16384 + (size_t)npmax * 128);
  if (smem > 110 * 1024) return rl_fail(RL_ERR_SHAPE, "wgrad: smem %zu too large", smem);
  { ProfScope _ps(KC_WGRAD, st);
  wgrad_kernel<<<dim3(nruns, nmb, nparts), 128, smem, st>>>(a); }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

}  // namespace rl
