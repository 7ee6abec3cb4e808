// rl_conv_fwd.cu — forward tensor-core kernels (tcgen05 / TMEM, operands staged in shared memory, weights by
// 1-D bulk TMA):
//   dense_fwd_kernel : BN -> ReLU -> Conv3x3(Cin -> 16) -> +bias -> Dropout2d scale -> concat-slice store (+ Σx, Σx²)
//                      reference layers.py:5-12 (DenseLayer) and the torch.cat at layers.py:33,39
//   gemm_px_kernel   : [pixels x K] x [K x N] pointwise GEMM with either the TransitionDown epilogue
//                      (+bias, Dropout2d scale, MaxPool2d(2) floor; layers.py:43-52) or a plain bf16 store
//                      (ConvTranspose2d column matrix, layers.py:61-63, and its input gradient)
//
// The growth-16 convolution is computed "taps-as-N": one GEMM  Y[p, (tap,co)] = Σ_c relu(bn(x))[p,c] · W[(tap,co), c]
// with N = 9*16 = 144 and K = Cin, so every activation is read exactly once and never needs a halo; the
// 3x3 structure is resolved on the 16-channel side by a shift-add (col2im) of the 9 tap planes in shared memory.
#include <cuda.h>

#include <algorithm>
#include <string.h>

#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

constexpr int kRing = 512;  // col2im ring, pixels

// ================================================================================================
// dense_fwd_kernel
// ================================================================================================
constexpr int kFwdThreads = 160;  // warps 0-3: operand producers + epilogue; warp 4: weight TMA + MMA issue

__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(kFwdThreads, 2) dense_fwd_kernel(DenseFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  // smem carve-up: s[Cin], t[Cin] floats | zs ring [4][kRing][4] floats | 2 stages x (A 16 KB + W 18 KB)
  float* s_s = reinterpret_cast<float*>(smem);
  float* s_t = s_s + a.Cin;
  float* zs = reinterpret_cast<float*>(smem + (((size_t)a.Cin * 8 + 127) & ~(size_t)127));
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(zs) + 4 * kRing * 16;
  constexpr int kStageBytes = 16384 + 18432;
  // ready: 128 producers stored + fenced their part of A[s];  wfull: weight chunk landed (TMA tx bytes);
  // done: the MMAs that read stage s retired;  acc_full: a tile's accumulator is complete;  acc_empty: 128 readers drained it
  __shared__ uint64_t bar_ready[2], bar_wfull[2], bar_done[2], bar_acc_full, bar_acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[4][32];
  __shared__ __align__(16) float s_bias16[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- work item: rows [R0,R1) of the level (all images stacked), columns [x0,x1)
  const int run = blockIdx.x / a.nstrips, strip = blockIdx.x % a.nstrips;
  const int R0 = run * a.rows_per_cta, R1 = min(R0 + a.rows_per_cta, a.rows_total);
  const int x0 = strip * a.strip_w, x1 = min(x0 + a.strip_w, a.W);
  const int rin0 = max(R0 - 1, 0), rin1 = min(R1 + 1, a.rows_total);
  const int xi0 = max(x0 - 1, 0), xi1 = min(x1 + 1, a.W);
  const int SW = xi1 - xi0;
  const int n_in = (rin1 - rin0) * SW;
  const int ntiles = (n_in + 127) / 128;
  const int nkc = (a.Cin + 63) / 64;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_ready[i], 128);
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_done[i], 1);
    }
    mbar_init(&bar_acc_full, 1);
    mbar_init(&bar_acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 4 * kRing * 4; i += kFwdThreads) zs[i] = 0.f;
  if (tid < 16) s_bias16[tid] = a.bias[tid];
  for (int c = tid; c < a.Cin; c += kFwdThreads) {
    BnCoef k = bn_coef(a.bn, c);
    s_s[c] = k.s;
    s_t[c] = k.t;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 4) {
    // =============================================================== weight TMA + MMA issue (one elected lane)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(144, 0, 0);
      auto issue_w = [&](uint32_t gg, int kc) {
        const uint32_t s = gg & 1, u = gg >> 1;
        if (u > 0) mbar_wait(&bar_done[s], (u - 1) & 1);  // MMAs that last read this stage have retired
        const int ck = min(64, a.Cin - kc * 64);
        const uint32_t wbytes = (uint32_t)(ck / 8) * 144 * 16;
        mbar_arrive_expect_tx(&bar_wfull[s], wbytes);
        bulk_g2s(stage_base + s * kStageBytes + 16384, a.wpk + (size_t)kc * 18432, wbytes, &bar_wfull[s]);
      };
      uint32_t g = 0;
      if (ntiles > 0) issue_w(0, 0);
      for (int tile = 0; tile < ntiles; ++tile) {
        if (tile > 0) {
          mbar_wait(&bar_acc_empty, (tile - 1) & 1);  // previous tile's accumulator has been read out
          tc_fence_after();
        }
        for (int kc = 0; kc < nkc; ++kc, ++g) {
          const uint32_t s = g & 1, u = g >> 1;
          const int ck = min(64, a.Cin - kc * 64);
          // next chunk's weights go out one step ahead (this tile's kc+1, or the next tile's chunk 0)
          if (kc + 1 < nkc) issue_w(g + 1, kc + 1);
          else if (tile + 1 < ntiles) issue_w(g + 1, 0);
          mbar_wait(&bar_ready[s], u & 1);
          mbar_wait(&bar_wfull[s], u & 1);
          tc_fence_after();
          const uint32_t aaddr = smem_u32(stage_base + s * kStageBytes), waddr = aaddr + 16384;
          for (int kk = 0; kk < ck / 16; ++kk)
            umma_ss(tmem, desc_kmajor(aaddr + kk * 2 * 16 * 128, 128), desc_kmajor(waddr + kk * 2 * 16 * 144, 144), idesc,
                    (kc | kk) != 0);
          umma_commit(&bar_done[s]);
          if (kc == nkc - 1) umma_commit(&bar_acc_full);
        }
      }
    }
  } else {
    // =============================================================== producers (BN+ReLU -> smem A) and epilogue
    float st_sum[16], st_sq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) st_sum[i] = st_sq[i] = 0.f;
    uint32_t g = 0;
    int emit_pos = 0;
    TileLoader<128> ld;
    const __nv_bfloat16* rowptr[4];
    auto set_rowptrs = [&](int t0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + TileLoader<128>::row_of(i);
        if (t < n_in) {
          const int rr = t / SW, cc = t - rr * SW;
          rowptr[i] = a.act + ((size_t)(rin0 + rr) * a.W + (xi0 + cc)) * a.Ct;
        } else {
          rowptr[i] = nullptr;
        }
      }
    };
    if (ntiles > 0) {
      set_rowptrs(0);
      ld.load(rowptr, min(8, a.Cin / 8));
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = tile * 128;
      // ---------------- K chunks of 64 channels: registers -> BN+ReLU -> smem A[s]; no block-wide barrier
      for (int kc = 0; kc < nkc; ++kc, ++g) {
        const int ck = min(64, a.Cin - kc * 64);
        const uint32_t s = g & 1, u = g >> 1;
        if (u > 0) mbar_wait(&bar_done[s], (u - 1) & 1);
        ld.store(stage_base + s * kStageBytes, rowptr, ck / 8, s_s + kc * 64, s_t + kc * 64);
        fence_proxy_async();  // before the next loads are issued (the fence waits for outstanding memory operations)
        mbar_arrive(&bar_ready[s]);
        if (kc + 1 < nkc) {  // prefetch the next chunk while this one is consumed
          const __nv_bfloat16* rp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] = rowptr[i] ? rowptr[i] + (kc + 1) * 64 : nullptr;
          ld.load(rp, min(8, (a.Cin - (kc + 1) * 64) / 8));
        }
      }
      // first chunk of the next tile streams in behind this tile's epilogue
      if (tile + 1 < ntiles) {
        set_rowptrs(t0 + 128);
        ld.load(rowptr, min(8, a.Cin / 8));
      }
      mbar_wait(&bar_acc_full, tile & 1);
      tc_fence_after();
      // ---------------- col2im scatter of the 9 tap planes
      const int t = t0 + tid;
      const bool in_valid = t < n_in;
      int rr = 0, cc = 0;
      if (in_valid) {
        rr = t / SW;
        cc = t - rr * SW;
      }
      const int r = rin0 + rr, xx = xi0 + cc;
      const int ri = r % a.H;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int dx = tap / 3 - 1, dy = tap % 3 - 1;  // tap slots are packed dx-major
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + tap * 16, v);
        if (tap == 8) {  // accumulator fully read by this thread: the next tile's MMAs may overwrite it
          tc_fence_before();
          mbar_arrive(&bar_acc_empty);
        }
        const int qr = r - dy, qx = xx - dx, qi = ri - dy;
        if (in_valid && qi >= 0 && qi < a.H && qr >= R0 && qr < R1 && qx >= x0 && qx < x1) {
          const int uidx = (t - dy * SW - dx + SW + 1) & (kRing - 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4* p = reinterpret_cast<float4*>(zs + ((size_t)q * kRing + uidx) * 4);
            float4 o = *p;
            o.x += v[q * 4 + 0];
            o.y += v[q * 4 + 1];
            o.z += v[q * 4 + 2];
            o.w += v[q * 4 + 3];
            *p = o;
          }
        }
        producers_sync();
      }
      // ---------------- emit finished output pixels
      const int lim = (tile == ntiles - 1) ? n_in : max(0, min(t0 + 128 - SW - 1, n_in));
      for (int e0 = emit_pos; e0 < lim; e0 += 128) {
        const int e = e0 + tid;
        if (e < lim) {
          const int er = e / SW, ec = e - er * SW;
          const int orow = rin0 + er, ox = xi0 + ec;
          if (orow >= R0 && orow < R1 && ox >= x0 && ox < x1) {
            const int uidx = (e + SW + 1) & (kRing - 1);
            const int img = orow / a.H;
            float o[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4* p = reinterpret_cast<float4*>(zs + ((size_t)q * kRing + uidx) * 4);
              float4 z = *p;
              *p = make_float4(0.f, 0.f, 0.f, 0.f);
              o[q * 4 + 0] = z.x;
              o[q * 4 + 1] = z.y;
              o[q * 4 + 2] = z.z;
              o[q * 4 + 3] = z.w;
            }
            uint32_t pk[8];
            float mk[16];
            if (a.mask != nullptr) {
              const float4* mp = reinterpret_cast<const float4*>(a.mask + (size_t)img * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 m4 = __ldg(mp + q);
                mk[4 * q] = m4.x;
                mk[4 * q + 1] = m4.y;
                mk[4 * q + 2] = m4.z;
                mk[4 * q + 3] = m4.w;
              }
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) mk[c] = 1.f;
            }
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              float y0 = (o[c] + s_bias16[c]) * mk[c], y1 = (o[c + 1] + s_bias16[c + 1]) * mk[c + 1];
              pk[c >> 1] = pack_bf16x2(y0, y1);
              float2 rb = unpack_bf16x2(pk[c >> 1]);
              st_sum[c] += rb.x;
              st_sq[c] += rb.x * rb.x;
              st_sum[c + 1] += rb.y;
              st_sq[c + 1] += rb.y * rb.y;
            }
            st256(a.act_out + ((size_t)orow * a.W + ox) * a.Ct + a.Cin, make_uint4(pk[0], pk[1], pk[2], pk[3]),
                  make_uint4(pk[4], pk[5], pk[6], pk[7]));
            if (a.x16 != nullptr)
              st256(a.x16 + ((size_t)orow * a.W + ox) * 16, make_uint4(pk[0], pk[1], pk[2], pk[3]),
                    make_uint4(pk[4], pk[5], pk[6], pk[7]));
          }
        }
      }
      emit_pos = max(emit_pos, lim);
      producers_sync();
    }

    // ---------------- per-channel Σx, Σx² of the 16 new channels (fp32 per thread -> fp64 atomics)
    if (a.stats_out != nullptr) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float s1 = warp_sum(st_sum[c]), s2 = warp_sum(st_sq[c]);
        if (lane == 0) {
          red[warp][c] = s1;
          red[warp][16 + c] = s2;
        }
      }
      producers_sync();
      if (tid < 32) {
        float v = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
        const int c = tid & 15;
        atomicAdd(a.stats_out + (tid < 16 ? 0 : a.Ct) + a.Cin + c, (double)v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ================================================================================================
// dense_fwd_tma_kernel — the same layer, fed by the tensor-map TMA engine.  One CTA per SM, 10 warps:
//   warp 9     : activation producer. One lane issues cp.async.bulk.tensor.2d (128 pixels x 64 channels, 128-byte
//                swizzle) into a 6-deep ring: up to 96 KB of raw activations in flight per SM with no registers held.
//   warps 0-3  : BN+ReLU in place on a landed stage (LDS.128 -> fma/relu -> STS.128 at the same address; a thread's
//                8-channel group is the same for all of its 8 rows, so s/t are fetched once per stage).
//   warp 8     : weight chunks by 1-D bulk TMA (3-deep ring) and tcgen05.mma issue (A: SWIZZLE_128B K-major straight
//                from the tensor-map layout, B: the packed chunk layout), two 144-column accumulators in TMEM.
//   warps 4-7  : epilogue (col2im of the 9 tap planes, bias, Dropout2d scale, concat-slice store, Σx, Σx²) of tile i
//                while the other roles are already on tile i+1.
// Requires a full-width strip (W <= 160): the 128 input pixels of a tile are then 128 consecutive arena rows.
// ================================================================================================
constexpr int kF2Threads = 448;  // warps 0-3 transform, 4-7 + 10-13 epilogue, 8 MMA issue, 9 activation TMA
constexpr int kF2A = 6, kF2W = 3;  // maximum activation ring depth; weight ring depth in streaming mode
constexpr int kF2ABytes = 16384, kF2WBytes = 18432;
constexpr size_t kF2Cap = 224 * 1024;

struct alignas(64) DenseFwdTmaArgs {
  CUtensorMap tmap;
  DenseFwdArgs a;
  int na;          // activation ring depth (3..6 stages of 16 KB)
  int w_resident;  // all weight chunks stay in shared memory for the whole CTA (else a 3-deep streaming ring)
  int strip_mode;  // 0: full-width rows through a 2-D map; 1: 128-column strips through a 3-D map (ch, x, row)
};

__device__ __forceinline__ void epilogue_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory"); }

__global__ void __launch_bounds__(kF2Threads, 1) dense_fwd_tma_kernel(const __grid_constant__ DenseFwdTmaArgs P) {
  const DenseFwdArgs& a = P.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B atoms want 1024-byte alignment
  const int nkc = (a.Cin + 63) / 64;
  const uint32_t NA = (uint32_t)P.na;
  const bool wres = P.w_resident != 0;
  uint8_t* sA = smem;
  uint8_t* sW = smem + NA * kF2ABytes;
  float* zs = reinterpret_cast<float*>(sW + (wres ? nkc : kF2W) * kF2WBytes);  // col2im ring [4][kRing][4]
  // per-channel BN scale / shift, stored per 64-channel chunk as [half][group j][4]: channel kc*64 + 8j + 4*half + i sits at
  // kc*64 + 32*half + 4j + i, so that the eight groups read by a quarter-warp are 128 contiguous bytes (no bank conflict)
  float* s_s = zs + 4 * kRing * 4;
  float* s_t = s_s + ((a.Cin + 63) & ~63);
  // a_full: TMA bytes landed; a_ready: 128 transform threads done; a_empty: the MMAs reading the stage retired
  __shared__ uint64_t a_full[kF2A], a_ready[kF2A], a_empty[kF2A], w_full[kF2W], w_empty[kF2W], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[2][4][16];
  __shared__ __align__(16) float s_bias16[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // work item: output rows [R0,R1) of the level (all images stacked) x output columns [x0,x1).  Full-width mode: the
  // input window is the whole row (SW = W, 128 consecutive arena pixels per tile).  Strip mode (levels wider than 160):
  // the input window of a strip is always the 128 columns [xs, xs+128) with xs = x0-1, one image row per tile, fetched
  // through a 3-D tensor map whose out-of-range columns (x = -1, x >= W) arrive as zeros.
  const bool strips = P.strip_mode != 0;
  const int run = strips ? blockIdx.x / a.nstrips : blockIdx.x, strip = strips ? blockIdx.x % a.nstrips : 0;
  const int R0 = run * a.rows_per_cta, R1 = min(R0 + a.rows_per_cta, a.rows_total);
  const int rin0 = max(R0 - 1, 0), rin1 = min(R1 + 1, a.rows_total);
  const int x0 = strips ? strip * a.strip_w : 0, x1 = strips ? min(x0 + a.strip_w, a.W) : a.W;
  const int xs = strips ? x0 - 1 : 0;
  const int SW = strips ? 128 : a.W;
  const int n_in = (rin1 - rin0) * SW;
  const int ntiles = (n_in + 127) / 128;
  const uint32_t total = (uint32_t)ntiles * (uint32_t)nkc;

  if (tid == 0) {
    for (int i = 0; i < kF2A; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 4);  // one arrival per transform warp
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kF2W; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < 4 * kRing * 4; i += kF2Threads) zs[i] = 0.f;
  if (tid < 16) s_bias16[tid] = a.bias[tid];
  for (int c = tid; c < a.Cin; c += kF2Threads) {
    BnCoef k = bn_coef(a.bn, c);
    const int o = (c & ~63) + 32 * ((c >> 2) & 1) + 4 * ((c >> 3) & 7) + (c & 3);
    s_s[o] = k.s;
    s_t[o] = k.t;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 9) {
    // =============================================================== activation producer
    if (lane == 0) {
      const int P0 = rin0 * a.W;
      uint32_t s = 0, ph = 1;  // ring slot and the parity of "slot free" (a fresh barrier passes a parity-1 wait)
      for (int tile = 0; tile < ntiles; ++tile)
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(&a_empty[s], ph);
          mbar_arrive_expect_tx(&a_full[s], kF2ABytes);
          if (strips) tma_load_3d(sA + s * kF2ABytes, &P.tmap, kc * 64, xs, rin0 + tile, &a_full[s]);
          else tma_load_2d(sA + s * kF2ABytes, &P.tmap, kc * 64, P0 + tile * 128, &a_full[s]);
          if (++s == NA) {
            s = 0;
            ph ^= 1;
          }
        }
    }
  } else if (warp == 8) {
    // =============================================================== weight TMA + MMA issue
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(144, 0, 0);
      auto issue_w = [&](uint32_t gg) {
        const int kc = (int)(gg % (uint32_t)nkc);
        const uint32_t s = gg % kF2W, u = gg / kF2W;
        if (u > 0) mbar_wait(&w_empty[s], (u - 1) & 1);
        const int ck = min(64, a.Cin - kc * 64);
        const uint32_t wbytes = (uint32_t)(ck / 8) * 144 * 16;
        mbar_arrive_expect_tx(&w_full[s], wbytes);
        bulk_g2s(sW + s * kF2WBytes, a.wpk + (size_t)kc * kF2WBytes, wbytes, &w_full[s]);
      };
      if (wres) {  // the layer's whole weight matrix: loaded once, reused by every tile of this CTA
        if (total > 0) {
          uint32_t wbytes = 0;
          for (int kc = 0; kc < nkc; ++kc) wbytes += (uint32_t)(min(64, a.Cin - kc * 64) / 8) * 144 * 16;
          mbar_arrive_expect_tx(&w_full[0], wbytes);
          for (int kc = 0; kc < nkc; ++kc)
            bulk_g2s(sW + kc * kF2WBytes, a.wpk + (size_t)kc * kF2WBytes, (uint32_t)(min(64, a.Cin - kc * 64) / 8) * 144 * 16,
                     &w_full[0]);
          mbar_wait(&w_full[0], 0);
        }
      } else {
        if (total > 0) issue_w(0);
        if (total > 1) issue_w(1);
      }
      // The issuing thread is a serial instruction stream: everything per stage is kept to a handful of adds
      // (descriptors advance by constants, ring slots and parities are carried, no divisions).
      const uint64_t adesc0 = desc_sw128_kmajor(smem_u32(sA));
      const uint64_t wdesc0 = desc_kmajor(smem_u32(sW), 144);
      const int last_n16 = (a.Cin - (nkc - 1) * 64) / 16;
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0, g = 0;
      for (int tile = 0; tile < ntiles; ++tile) {
        const uint32_t buf = tile & 1;
        const uint32_t dacc = tmem + buf * 256;
        if (tile >= 2) {
          mbar_wait(&acc_empty[buf], ((tile >> 1) - 1) & 1);  // the epilogue drained this accumulator
          tc_fence_after();
        }
        for (int kc = 0; kc < nkc; ++kc, ++g) {
          if (!wres && g + 2 < total) issue_w(g + 2);
          const int n16 = kc == nkc - 1 ? last_n16 : 4;
          mbar_wait(&a_ready[sa], pa);
          if (!wres) mbar_wait(&w_full[sw], pw);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(sa * (kF2ABytes >> 4));
          const uint64_t wd = wdesc0 + (uint64_t)((wres ? (uint32_t)kc : sw) * (kF2WBytes >> 4));
          if (n16 > 0) umma_ss(dacc, ad, wd, idesc, kc != 0);
          if (n16 > 1) umma_ss(dacc, ad + 2, wd + 288, idesc, 1);
          if (n16 > 2) umma_ss(dacc, ad + 4, wd + 576, idesc, 1);
          if (n16 > 3) umma_ss(dacc, ad + 6, wd + 864, idesc, 1);
          umma_commit(&a_empty[sa]);
          if (!wres) umma_commit(&w_empty[sw]);
          if (kc == nkc - 1) umma_commit(&acc_full[buf]);
          if (++sa == NA) {
            sa = 0;
            pa ^= 1;
          }
          if (++sw == kF2W) {
            sw = 0;
            pw ^= 1;
          }
        }
      }
    }
  } else if (warp < 4) {
    // =============================================================== BN + ReLU in place
    // physical 16-byte slot q = i*128 + tid of a stage: row = q >> 3, slot-in-row = q & 7, holding the logical
    // 8-channel group j = slot ^ (row & 7) -- the same j for every i
    const int j = (tid & 7) ^ ((tid >> 3) & 7);
    uint32_t s = 0, ph = 0;
    for (int tile = 0; tile < ntiles; ++tile)
      for (int kc = 0; kc < nkc; ++kc) {
        const int ck = min(64, a.Cin - kc * 64);
        mbar_wait(&a_full[s], ph);
        if (8 * j < ck) {
          const float* sp = s_s + kc * 64 + 4 * j;
          const float* tp = s_t + kc * 64 + 4 * j;
          const float4 s0 = *reinterpret_cast<const float4*>(sp), s1 = *reinterpret_cast<const float4*>(sp + 32);
          const float4 t0 = *reinterpret_cast<const float4*>(tp), t1 = *reinterpret_cast<const float4*>(tp + 32);
          uint4* base = reinterpret_cast<uint4*>(sA + s * kF2ABytes) + tid;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = base[i * 128];
            const float2 p0 = unpack_bf16x2(v.x), p1 = unpack_bf16x2(v.y), p2 = unpack_bf16x2(v.z), p3 = unpack_bf16x2(v.w);
            uint4 o;
            o.x = pack_bf16x2_relu(fmaf(p0.x, s0.x, t0.x), fmaf(p0.y, s0.y, t0.y));
            o.y = pack_bf16x2_relu(fmaf(p1.x, s0.z, t0.z), fmaf(p1.y, s0.w, t0.w));
            o.z = pack_bf16x2_relu(fmaf(p2.x, s1.x, t1.x), fmaf(p2.y, s1.y, t1.y));
            o.w = pack_bf16x2_relu(fmaf(p3.x, s1.z, t1.z), fmaf(p3.y, s1.w, t1.w));
            base[i * 128] = o;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
        if (++s == NA) {
          s = 0;
          ph ^= 1;
        }
      }
  } else {
    // =============================================================== epilogue: two groups of four warps.  Group g (warps
    // 4-7: g = 0, warps 10-13: g = 1) owns output channels 8g .. 8g+7 of every pixel: its own half of the accumulator
    // columns of each tap and its own two planes of the col2im ring, so the groups never meet and the shift-add of a tile
    // (the latency-bound part of this kernel: TMEM read -> ring read-modify-write per tap) runs twice as wide.
    const int eg = warp >= 10 ? 1 : 0;
    const int ew = warp & 3;                       // TMEM lane quarter of this warp
    const int et = ew * 32 + lane;                 // accumulator row = pixel of the tile
    const int egw = eg ? warp - 10 : warp - 4;     // 0..3 inside the group (only used for the statistics scratch)
    float* zsg = zs + (size_t)(2 * eg) * kRing * 4;  // ring planes 2g, 2g+1: [plane][slot][4 channels]
    float st_sum[8], st_sq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st_sum[i] = st_sq[i] = 0.f;
    int emit_pos = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = tile * 128;
      const uint32_t buf = tile & 1;
      mbar_wait(&acc_full[buf], (tile >> 1) & 1);
      tc_fence_after();
      const int t = t0 + et;
      bool in_valid = t < n_in;
      int rr = 0, cc = 0;
      if (in_valid) {
        rr = t / SW;
        cc = t - rr * SW;
      }
      const int r = rin0 + rr, xx = xs + cc;
      in_valid = in_valid && xx >= 0 && xx < a.W;  // strip mode: columns of the window that lie outside the image
      const int ri = r % a.H;
      // Tap slots are packed dx-major, so one dx group = 48 consecutive accumulator columns = the three dy taps.
      // Within a tile (128 consecutive pixels) two threads can only meet in a ring slot through different dx when
      // SW >= 128 (a dy step moves the slot by SW): one barrier per dx group then suffices; narrower levels keep
      // one barrier per tap.
      const bool wide = SW >= 128;
      const uint32_t tacc = tmem + buf * 256 + ((uint32_t)(ew * 32) << 16) + 8 * eg;
      float v[24];
      tmem_ld8_nowait(tacc, v);
      tmem_ld8_nowait(tacc + 16, v + 8);
      tmem_ld8_nowait(tacc + 32, v + 16);
      tmem_wait_ld();
      reg_fence8(v);
      reg_fence8(v + 8);
      reg_fence8(v + 16);
#pragma unroll
      for (int dxi = 0; dxi < 3; ++dxi) {
        const int dx = dxi - 1;
        float vn[24];
        if (dxi < 2) {  // next group's accumulator columns stream in behind this group's shift-add
          tmem_ld8_nowait(tacc + (dxi + 1) * 48, vn);
          tmem_ld8_nowait(tacc + (dxi + 1) * 48 + 16, vn + 8);
          tmem_ld8_nowait(tacc + (dxi + 1) * 48 + 32, vn + 16);
        }
        const int qx = xx - dx;
        const bool okx = in_valid && qx >= x0 && qx < x1;
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi) {
          const int dy = dyi - 1;
          const int qr = r - dy, qi = ri - dy;
          if (okx && qi >= 0 && qi < a.H && qr >= R0 && qr < R1) {
            const int uidx = (t - dy * SW - dx + SW + 1) & (kRing - 1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float4* p = reinterpret_cast<float4*>(zsg + ((size_t)q * kRing + uidx) * 4);
              float4 o = *p;
              o.x += v[dyi * 8 + q * 4 + 0];
              o.y += v[dyi * 8 + q * 4 + 1];
              o.z += v[dyi * 8 + q * 4 + 2];
              o.w += v[dyi * 8 + q * 4 + 3];
              *p = o;
            }
          }
          if (!wide) epilogue_sync(eg);
        }
        if (dxi < 2) {
          tmem_wait_ld();
          reg_fence8(vn);
          reg_fence8(vn + 8);
          reg_fence8(vn + 16);
#pragma unroll
          for (int i = 0; i < 24; ++i) v[i] = vn[i];
          if (dxi == 1) {  // accumulator fully read by this thread: the MMAs two tiles ahead may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
        }
        if (wide) epilogue_sync(eg);
      }
      const int lim = (tile == ntiles - 1) ? n_in : max(0, min(t0 + 128 - SW - 1, n_in));
      for (int e0 = emit_pos; e0 < lim; e0 += 128) {
        const int e = e0 + et;
        if (e < lim) {
          const int er = e / SW, ec = e - er * SW;
          const int orow = rin0 + er, ox = xs + ec;
          if (orow >= R0 && orow < R1 && ox >= x0 && ox < x1) {
            const int uidx = (e + SW + 1) & (kRing - 1);
            const int img = orow / a.H;
            float o[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float4* p = reinterpret_cast<float4*>(zsg + ((size_t)q * kRing + uidx) * 4);
              float4 z = *p;
              *p = make_float4(0.f, 0.f, 0.f, 0.f);
              o[q * 4 + 0] = z.x;
              o[q * 4 + 1] = z.y;
              o[q * 4 + 2] = z.z;
              o[q * 4 + 3] = z.w;
            }
            uint32_t pk[4];
            float mk[8];
            if (a.mask != nullptr) {
              const float4* mp = reinterpret_cast<const float4*>(a.mask + (size_t)img * 16 + 8 * eg);
              const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
              mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w;
              mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) mk[c] = 1.f;
            }
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              float y0 = (o[c] + s_bias16[8 * eg + c]) * mk[c], y1 = (o[c + 1] + s_bias16[8 * eg + c + 1]) * mk[c + 1];
              pk[c >> 1] = pack_bf16x2(y0, y1);
              float2 rb = unpack_bf16x2(pk[c >> 1]);
              st_sum[c] += rb.x;
              st_sq[c] += rb.x * rb.x;
              st_sum[c + 1] += rb.y;
              st_sq[c + 1] += rb.y * rb.y;
            }
            const uint4 pv = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(a.act_out + ((size_t)orow * a.W + ox) * a.Ct + a.Cin + 8 * eg) = pv;
            if (a.x16 != nullptr) *reinterpret_cast<uint4*>(a.x16 + ((size_t)orow * a.W + ox) * 16 + 8 * eg) = pv;
          }
        }
      }
      emit_pos = max(emit_pos, lim);
      epilogue_sync(eg);
    }
    if (a.stats_out != nullptr) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float s1 = warp_sum(st_sum[c]), s2 = warp_sum(st_sq[c]);
        if (lane == 0) {
          red[eg][egw][c] = s1;
          red[eg][egw][8 + c] = s2;
        }
      }
      epilogue_sync(eg);
      if (et < 16) {
        float vsum = red[eg][0][et] + red[eg][1][et] + red[eg][2][et] + red[eg][3][et];
        const int c = et & 7;
        atomicAdd(a.stats_out + (et < 8 ? 0 : a.Ct) + a.Cin + 8 * eg + c, (double)vsum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_dense_fwd_tma(const DenseFwdArgs& a, const void* tmap, int grid, cudaStream_t st) {
  constexpr size_t kCap = kF2Cap;  // dynamic shared memory of the one resident CTA
  const int strip_mode = a.W > 160;
  if (!strip_mode && a.nstrips != 1) return rl_fail(RL_ERR_SHAPE, "dense_fwd_tma: a level up to 160 wide is one strip");
  if (strip_mode && (a.strip_w < 1 || a.strip_w > 126)) return rl_fail(RL_ERR_SHAPE, "dense_fwd_tma: strips are at most 126 wide");
  DenseFwdTmaArgs P;
  memcpy(&P.tmap, tmap, sizeof(CUtensorMap));
  P.a = a;
  P.strip_mode = strip_mode;
  const int nkc = (a.Cin + 63) / 64;
  const size_t fixed = 1024 + 4 * kRing * 16 + (size_t)((a.Cin + 63) & ~63) * 8;
  const size_t wall = (size_t)nkc * kF2WBytes;
  if (fixed + wall + 3 * kF2ABytes <= kCap) {
    P.w_resident = 1;
    P.na = (int)std::min<size_t>(kF2A, (kCap - fixed - wall) / kF2ABytes);
  } else {
    P.w_resident = 0;
    P.na = kF2A;
  }
  const size_t smem = fixed + (P.w_resident ? wall : (size_t)kF2W * kF2WBytes) + (size_t)P.na * kF2ABytes;
  if (smem > kCap) return rl_fail(RL_ERR_SHAPE, "dense_fwd_tma: smem %zu too large", smem);
  {
    ProfScope _ps(KC_DENSE_FWD, st);
    dense_fwd_tma_kernel<<<grid, kF2Threads, smem, st>>>(P);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int encode_arena_tmap(void* tmap_out, const void* arena, long npx, int Ct, int W) {
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
  // levels up to 160 wide: [pixels][Ct], box 128 px x 64 ch; wider levels: [rows][W][Ct], box 1 row x 128 columns x 64 ch
  const bool three = W > 160;
  const cuuint64_t dims[3] = {(cuuint64_t)Ct, (cuuint64_t)(three ? W : npx), (cuuint64_t)(three ? npx / W : 1)};
  const cuuint64_t strides[2] = {(cuuint64_t)Ct * 2, (cuuint64_t)Ct * 2 * (cuuint64_t)W};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, three ? 3 : 2, const_cast<void*>(arena), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rl_fail(RL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%ld][%d]", (int)r, npx, Ct);
  memcpy(tmap_out, &tm, sizeof(tm));
  return RL_OK;
}

size_t dense_fwd_smem(int Cin) { return (((size_t)Cin * 8 + 127) & ~(size_t)127) + 4 * kRing * 16 + 2 * (16384 + 18432); }

// ================================================================================================
// gemm_px_kernel
// ================================================================================================
__global__ void __launch_bounds__(128, 2) gemm_px_kernel(GemmPxArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_s = reinterpret_cast<float*>(smem);
  float* s_t = s_s + a.K;
  float* s_bias = reinterpret_cast<float*>(smem + (((size_t)a.K * 8 + 127) & ~(size_t)127));  // [256] this part's bias
  float* s_mask = s_bias + 256;                                                               // [256] dropout scale of the tile's image
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(s_mask + 256);
  const int part = blockIdx.y;
  const int n0 = a.part_n0[part], np = a.part_np[part];
  const int stageB = 8 * np * 16;           // bytes of one packed weight chunk slot for this part
  const int kStageBytes = 16384 + stageB;
  __shared__ uint64_t bar_full[2], bar_done[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = (a.K + 63) / 64;
  const uint8_t* wpart = a.wpk + (size_t)nkc * 8 * 16 * n0;  // parts are stored back to back

  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    fence_mbar_init();
  }
  const uint32_t tmem_cols = np <= 128 ? 128u : 256u;  // accumulator columns of this part: 3 CTAs per SM fit when <= 128
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (a.transform)
    for (int c = tid; c < a.K; c += 128) {
      BnCoef k = bn_coef(a.bn, c);
      s_s[c] = k.s;
      s_t[c] = k.t;
    }
  for (int c = tid; c < 256; c += 128) {
    s_bias[c] = (a.bias != nullptr && c < np) ? a.bias[n0 + c] : 0.f;
    s_mask[c] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16((uint32_t)np, 0, 0);

  uint32_t g = 0;
  bool w_issued = false;
  int mask_img = -1;
  TileLoader<128> ld;
  const __nv_bfloat16* rowptr[4];

  // row m of tile t -> arena pixel (or -1).  Within each 32-row quarter (= the TMEM lanes one warp can read) the four
  // positions of window w' sit in rows w', w' + 8, w' + 16, w' + 24: exactly the four lanes that two tcgen05.ld.16x256b
  // deliver to ONE thread, so the 2x2 max-pool is a per-thread operation (no shuffles, no redundant lanes)
  auto pixel_of = [&](long t, int m) -> long {
    const int pr_ = (int)(t / a.wblocks), wb_ = (int)(t - (long)pr_ * a.wblocks);
    const int win = ((m >> 5) << 3) | (m & 7), q = (m >> 3) & 3;  // m = 32 * (win / 8) + 8 * q + win % 8
    const int wi = wb_ * 32 + win;
    if (wi >= a.Wp) return -1;
    const int img = pr_ / a.Hp, py = pr_ - img * a.Hp;
    return ((long)img * a.H + 2 * py + (q >> 1)) * a.W + 2 * wi + (q & 1);
  };
  auto set_rowptrs = [&](long t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long P = pixel_of(t, TileLoader<128>::row_of(i));
      rowptr[i] = P >= 0 ? a.A + (size_t)P * a.lda + a.a_off : nullptr;
    }
  };
  // the first K chunk of a tile is fetched before the previous tile's epilogue, so its latency hides behind it
  if ((long)blockIdx.x < a.ntiles) {
    set_rowptrs(blockIdx.x);
    ld.load(rowptr, min(8, a.K / 8));
  }
  for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    // ---- row -> pixel mapping
    const int pr = (int)(tile / a.wblocks), wb = (int)(tile % a.wblocks);
    {
      const int img = pr / a.Hp;
      if (a.mask != nullptr && img != mask_img) {  // visible to the epilogue through the main loop's barriers
        for (int c = tid; c < np; c += 128) s_mask[c] = a.mask[(size_t)img * a.N + n0 + c];
        mask_img = img;
      }
    }
    for (int kc = 0; kc < nkc; ++kc) {
      const int ck = min(64, a.K - kc * 64);
      const uint32_t s = g & 1, u = g >> 1;
      uint8_t* sA = stage_base + s * kStageBytes;
      uint8_t* sW = sA + 16384;
      if (u > 0) mbar_wait(&bar_done[s], (u - 1) & 1);
      if (tid == 0 && !w_issued) {
        const uint32_t wbytes = (uint32_t)(ck / 8) * np * 16;
        mbar_arrive_expect_tx(&bar_full[s], wbytes);
        bulk_g2s(sW, wpart + (size_t)kc * stageB, wbytes, &bar_full[s]);
      }
      ld.store(sA, rowptr, ck / 8, a.transform ? s_s + kc * 64 : nullptr, s_t + kc * 64);
      fence_proxy_async();  // before the next loads are issued (the fence waits for outstanding memory operations)
      if (kc + 1 < nkc) {
        const __nv_bfloat16* rp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rp[i] = rowptr[i] ? rowptr[i] + (kc + 1) * 64 : nullptr;
        ld.load(rp, min(8, (a.K - (kc + 1) * 64) / 8));
      }
      __syncthreads();
      if (tid == 0) {
        mbar_wait(&bar_full[s], u & 1);
        tc_fence_after();
        const uint32_t aaddr = smem_u32(sA), waddr = smem_u32(sW);
        for (int kk = 0; kk < ck / 16; ++kk)
          umma_ss(tmem, desc_kmajor(aaddr + kk * 2 * 16 * 128, 128), desc_kmajor(waddr + kk * 2 * 16 * np, np), idesc,
                  (kc | kk) != 0);
        umma_commit(&bar_done[s]);
        const bool more = (kc + 1 < nkc) || (tile + gridDim.x < a.ntiles);
        if (more) {  // next chunk's weights, one iteration ahead
          const int nk = (kc + 1 < nkc) ? kc + 1 : 0;
          const int nck = min(64, a.K - nk * 64);
          const uint32_t ns = (g + 1) & 1, nu = (g + 1) >> 1;
          if (nu > 0) mbar_wait(&bar_done[ns], (nu - 1) & 1);
          const uint32_t wbytes = (uint32_t)(nck / 8) * np * 16;
          mbar_arrive_expect_tx(&bar_full[ns], wbytes);
          bulk_g2s(stage_base + ns * kStageBytes + 16384, wpart + (size_t)nk * stageB, wbytes, &bar_full[ns]);
        }
      }
      w_issued = true;
      ++g;
    }
    if (tile + gridDim.x < a.ntiles) {  // next tile's first chunk streams in behind this tile's epilogue
      set_rowptrs(tile + gridDim.x);
      ld.load(rowptr, min(8, a.K / 8));
    }
    {
      const uint32_t s = (g - 1) & 1, u = (g - 1) >> 1;
      mbar_wait(&bar_done[s], u & 1);
      tc_fence_after();
    }
    // ---- epilogue: thread (warp, lane) owns window 8 * warp + lane / 4 and, in every 16-column block, the four columns
    // 8k + 2(lane % 4) + e; the packed weights permute the output channels so that those are channels 4(lane % 4) + 2k + e
    // of the block (rl_misc.cu pack kind 2): 8 contiguous bytes per thread, 32 per quad.  +bias and the non-negative
    // Dropout2d scale are the same for the four pixels of a window, so they commute with the max.
    {
      const int win = warp * 8 + (lane >> 2), qc = lane & 3, wi = wb * 32 + win;
      const int img = pr / a.Hp, py = pr % a.Hp;
      const bool valid = wi < a.Wp;
      const size_t opx = ((size_t)img * a.Hp + py) * a.Wp + wi;
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < np; c0 += 16) {
        float r0[8], r1[8];
        tmem_ld_16x256b_x2(lane_base + c0, r0);                      // positions 0 (regs 4k+e) and 1 (regs 4k+2+e)
        tmem_ld_16x256b_x2(lane_base + (16u << 16) + c0, r1);        // positions 2 and 3
        tmem_wait_ld();
        reg_fence8(r0);
        reg_fence8(r1);
        const int cl = c0 + 4 * qc;  // first of this thread's four channels (part-local)
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cl), m4 = *reinterpret_cast<const float4*>(s_mask + cl);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
        float o[4];
        uint32_t am = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // channel cl + i = 2k + e
          const int k = i >> 1, e = i & 1;
          float mx = r0[4 * k + e];
          uint32_t arg = 0;
          const float v1 = r0[4 * k + 2 + e], v2 = r1[4 * k + e], v3 = r1[4 * k + 2 + e];
          if (v1 > mx) { mx = v1; arg = 1; }   // strict: the first position wins a tie, as torch's max_pool2d
          if (v2 > mx) { mx = v2; arg = 2; }
          if (v3 > mx) { mx = v3; arg = 3; }
          am |= arg << (8 * i);
          o[i] = (mx + bb[i]) * mm[i];
        }
        if (valid) {
          const uint2 pk = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
          const int ch = a.o_off + n0 + cl;
          *reinterpret_cast<uint2*>(a.out + opx * a.ldo + ch) = pk;
          if (a.out_xp != nullptr)
            *reinterpret_cast<uint2*>(a.out_xp + ((size_t)(ch >> 4) * (size_t)a.out_xp_npx + opx) * 16 + (ch & 15)) = pk;
          if (a.amax) *reinterpret_cast<uint32_t*>(a.amax + opx * a.N + n0 + cl) = am;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

size_t gemm_px_smem(int K, int np_max) { return (((size_t)K * 8 + 127) & ~(size_t)127) + 2048 + 2 * (16384 + (size_t)8 * np_max * 16); }

// ------------------------------------------------------------------------------------------------ launchers
int launch_dense_fwd(const DenseFwdArgs& a, int grid, cudaStream_t st) {
  size_t smem = dense_fwd_smem(a.Cin);
  if (smem > 110 * 1024) return rl_fail(RL_ERR_SHAPE, "dense_fwd: smem %zu too large", smem);
  {
    ProfScope _ps(KC_DENSE_FWD, st);
    dense_fwd_kernel<<<grid, kFwdThreads, smem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int launch_gemm_px(const GemmPxArgs& a, int grid_x, int nparts, cudaStream_t st) {
  int npmax = 0;
  for (int i = 0; i < nparts; ++i) npmax = a.part_np[i] > npmax ? a.part_np[i] : npmax;
  size_t smem = gemm_px_smem(a.K, npmax);
  if (smem > 110 * 1024) return rl_fail(RL_ERR_SHAPE, "gemm_px: smem %zu too large (K=%d np=%d)", smem, a.K, npmax);
  {
    ProfScope _ps(KC_GEMM_PX, st);
    gemm_px_kernel<<<dim3(grid_x, nparts), 128, smem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int configure_fwd_kernels() {
  RL_CUDA(cudaFuncSetAttribute(dense_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2Cap));
  RL_CUDA(cudaFuncSetAttribute(dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  RL_CUDA(cudaFuncSetAttribute(gemm_px_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  return RL_OK;
}

}  // namespace rl
