// rl_tu.cu — TransitionUp (reference layers.py:58-69): ConvTranspose2d(80 -> 80, kernel 3, stride 2, padding 0) + bias,
// center_crop (offset 0 for every shape in use, layers.py:82-86) and the concat with the skip connection.
//
// A stride-2 transposed convolution decomposes by output parity.  With (Y, X) = (2I + py, 2J + px):
//   y[2I  , 2J  ] = x[I,J]·W00 + x[I-1,J]·W20 + x[I,J-1]·W02 + x[I-1,J-1]·W22
//   y[2I  , 2J+1] = x[I,J]·W01 + x[I-1,J]·W21
//   y[2I+1, 2J  ] = x[I,J]·W10 + x[I,J-1]·W12
//   y[2I+1, 2J+1] = x[I,J]·W11                                   (x = 0 outside the input)
// so one tile of 128 low-resolution positions needs 9 x 5 MMAs (K = 80 input channels per tap, N = 80 output channels)
// into four accumulators, whose A operands are row-shifted views (start-address offsets) of ONE input halo tile, and the
// epilogue writes each parity's pixels straight into the cropped concat slice (+ bias, + the planar mirror).  The
// [pixels x 720] column matrix of the GEMM + col2im formulation never exists.
#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

constexpr int kTuThreads = 256;
constexpr int kTuWBytes = 9 * 10 * 80 * 16;          // packed weights: [tap][ci chunk][co][8 ci]
constexpr int kTuTileRows = 177;                     // (TR + 1) * TP + guard rows <= 5 * 32 + 2; odd: the loader's 16-byte
                                                     // stores of one pixel's ten chunks land in different banks
constexpr int kTuTileBytes = 10 * kTuTileRows * 16;  // one input halo tile in chunk layout T[j][row][8]
constexpr size_t kTuFwdSmem = kTuWBytes + 2 * (size_t)kTuTileBytes + 80 * sizeof(float);

__global__ void __launch_bounds__(kTuThreads, 1) tu_fwd_kernel(TuFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sX = smem + kTuWBytes;  // two tiles
  float* s_bias = reinterpret_cast<float*>(sX + 2 * kTuTileBytes);
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TP = a.TP, TR = 128 / TP, tp_shift = 31 - __clz(TP);
  const int hp = a.h + 1;
  const int nwin = (TR + 1) * TP;  // positions of a halo tile; smem row = 1 + window index (row 0: guard for the -1 view)

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid < 80) s_bias[tid] = a.bias[tid];
  for (int i = tid; i < 2 * kTuTileBytes / 16; i += kTuThreads) reinterpret_cast<uint4*>(sX)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_w, kTuWBytes);
    bulk_g2s(sW, a.wpk, kTuWBytes, &bar_w);
  }

  // window position rr of tile -> input pixel: row vr0 - 1 + rr / TP of the per-image (h + 1)-row grid (row h of an image
  // and everything outside are zero), column J0 - 1 + rr % TP
  auto issue_tile = [&](long tile, uint8_t* dst) {
    const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
    const int vr0 = tr * TR, J0 = tc * (TP - 1);
    for (int p = tid; p < nwin * 10; p += kTuThreads) {  // ten consecutive threads fetch one pixel's 160 contiguous bytes
      const int rr = p / 10, jc = p - rr * 10;
      const int rw = rr >> tp_shift, c = rr & (TP - 1);
      const int vr = vr0 - 1 + rw, j = J0 - 1 + c;
      const int b = vr >= 0 ? vr / hp : 0, i = vr - b * hp;
      const bool ok = vr >= 0 && b < a.B && i < a.h && j >= 0 && j < a.w;
      const __nv_bfloat16* src = ok ? a.A + ((size_t)((long)b * a.h + i) * a.w + j) * a.lda + a.a_off + jc * 8 : a.A;
      cp_async16(dst + ((size_t)jc * kTuTileRows + 1 + rr) * 16, src, ok);
    }
    cp_async_commit();
  };

  const uint32_t idesc = umma_idesc_bf16(80, 0, 0);
  long tile = blockIdx.x;
  int it = 0;
  if (tile < a.ntiles) issue_tile(tile, sX);
  uint32_t n_mma = 0;
  for (; tile < a.ntiles; tile += gridDim.x, ++it) {
    uint8_t* cur = sX + (it & 1) * kTuTileBytes;
    const long next = tile + gridDim.x;
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();  // tile `it` is in shared memory; every thread is past the epilogue of tile it-1 (TMEM reusable)
    // the next tile's copies start only now (their buffer's last readers, the MMAs of tile it-1, have retired): issued before
    // the fence they would be waited for by it
    if (next < a.ntiles) issue_tile(next, sX + ((it + 1) & 1) * kTuTileBytes);
    if (tid == 0) {
      if (it == 0) mbar_wait(&bar_w, 0);
      tc_fence_after();
      const uint64_t ad0 = desc_kmajor(smem_u32(cur) + 16, kTuTileRows);  // + guard row
      const uint64_t bd0 = desc_kmajor(smem_u32(sW), 80);
#pragma unroll
      for (int par = 0; par < 4; ++par) {
        const int py = par >> 1, px = par & 1;
        bool first = true;
#pragma unroll
        for (int ai = 0; ai < 2; ++ai) {
#pragma unroll
          for (int bi = 0; bi < 2; ++bi) {
            if ((ai == 1 && py == 1) || (bi == 1 && px == 1)) continue;
            const int tap = (py + 2 * ai) * 3 + (px + 2 * bi);
            const long roff = (long)(TP * (1 - ai) - bi);
            const uint64_t ad = ad0 + (uint64_t)roff, bd = bd0 + (uint64_t)(tap * (10 * 80 * 16 >> 4));
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
              umma_ss(tmem + par * 80, ad + (uint64_t)(ks * 2 * kTuTileRows), bd + (uint64_t)(ks * 2 * 80), idesc, first ? 0 : 1);
              first = false;
            }
          }
        }
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, n_mma & 1);
    ++n_mma;
    tc_fence_after();
    // ---- epilogue: warps 0-3 write parities (0,0) and (0,1), warps 4-7 (1,0) and (1,1); lane quarter = warp % 4
    {
      const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
      const int m = (warp & 3) * 32 + lane, r = m >> tp_shift, c = m & (TP - 1);
      const int vr = tr * TR + r, J = tc * (TP - 1) - 1 + c;
      const int b = vr / hp, I = vr - b * hp;
      const int py = warp >> 2;
      const int Y = 2 * I + py;
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int X = 2 * J + px;
        const bool ok = c >= 1 && b < a.B && J <= a.w && Y < a.H && X < a.W;
        const size_t P = ok ? ((size_t)((long)b * a.H + Y) * a.W + X) : 0;
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (py * 2 + px) * 80;
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          float v[16];
          tmem_ld16(tacc + 16 * g, v);
          if (ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += s_bias[16 * g + i];
            const uint4 lo = pack8(v), hi = pack8(v + 8);
            st256(a.out + P * a.ldo + a.o_off + 16 * g, lo, hi);
            if (a.out_xp != nullptr) st256(a.out_xp + dact_index(a.npx_out, (long)P, a.o_off + 16 * g), lo, hi);
          }
        }
      }
    }
    tc_fence_before();
  }
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ================================================================================================
// backward.  With dz split into parity planes dzP[py][px][I][J] = dz[2I + py, 2J + px] (zero outside the output):
//   dX[i, j, ci]      = Σ_{ky,kx} Σ_co dzP[ky&1][kx&1][i + (ky>>1), j + (kx>>1), co] · W[ci, co, ky, kx]
//   dW[ci, co, ky, kx] = Σ_{i,j}  x[i, j, ci] · dzP[ky&1][kx&1][i + (ky>>1), j + (kx>>1), co]
// Both are GEMMs over tiles of TR x TP low-resolution positions whose shifted operands are row-shifted views
// (start-address offsets 0, 1, TP, TP + 1) of (TR + 1) x TP tiles of the planes.
// ================================================================================================
constexpr int kTuPlaneRows = 169;                       // (TR + 1) * TP + 1 <= 161; odd (bank spread, as above)
constexpr int kTuPlaneBytes = 10 * kTuPlaneRows * 16;   // one plane tile, chunk layout T[co chunk][row][8]
constexpr size_t kTuDgradSmem = kTuWBytes + 4 * (size_t)kTuPlaneBytes;
constexpr int kTuXRows = 129;                           // 128 K rows (+1: odd pitch)
constexpr int kTuXBytes = 16 * kTuXRows * 16;           // x tile as the M = 128 operand: chunks 10..15 stay zero
constexpr int kTuWgBuf = kTuXBytes + 2 * kTuPlaneBytes;
constexpr size_t kTuWgradSmem = 2 * (size_t)kTuWgBuf;

// plane tile of `tile` -> dst (cp.async): window position rr = r' * TP + c <-> plane position (vr0 + r', J0 + c)
__device__ __forceinline__ void tu_issue_plane(const TuBwdArgs& a, int par, int vr0, int J0, int TP, int tp_shift, int nwin,
                                               uint8_t* dst, int tid, int nthreads) {
  const int hp = a.h + 1, wp = a.w + 1;
  const __nv_bfloat16* plane = a.dzp + (size_t)par * a.npos * 80;
  for (int p = tid; p < nwin * 10; p += nthreads) {
    const int rr = p / 10, jc = p - rr * 10;
    const int vr = vr0 + (rr >> tp_shift), J = J0 + (rr & (TP - 1));
    const bool ok = vr < a.B * hp && J < wp;
    cp_async16(dst + ((size_t)jc * kTuPlaneRows + rr) * 16, ok ? plane + ((size_t)vr * wp + J) * 80 + jc * 8 : plane, ok);
  }
}

__global__ void __launch_bounds__(256, 1) tu_dgrad_kernel(TuBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sP = smem + kTuWBytes;  // four plane tiles
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TP = a.TP, TR = 128 / TP, tp_shift = 31 - __clz(TP);
  const int hp = a.h + 1;
  const int nwin = (TR + 1) * TP;
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  for (int i = tid; i < 4 * kTuPlaneBytes / 16; i += 256) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_w, kTuWBytes);
    bulk_g2s(sW, a.wpk, kTuWBytes, &bar_w);
  }
  auto issue_tile = [&](long tile) {
    const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
#pragma unroll
    for (int par = 0; par < 4; ++par)
      tu_issue_plane(a, par, tr * TR, tc * (TP - 1), TP, tp_shift, nwin, sP + par * kTuPlaneBytes, tid, 256);
    cp_async_commit();
  };
  const uint32_t idesc = umma_idesc_bf16(80, 0, 0);
  long tile = blockIdx.x;
  if (tile < a.ntiles) issue_tile(tile);
  uint32_t n_mma = 0;
  for (; tile < a.ntiles; tile += gridDim.x) {
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();  // the tile is in shared memory; every thread is past the previous epilogue (TMEM reusable)
    if (tid == 0) {
      if (n_mma == 0) mbar_wait(&bar_w, 0);
      tc_fence_after();
      const uint64_t bd0 = desc_kmajor(smem_u32(sW), 80);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3, par = (ky & 1) * 2 + (kx & 1);
        const uint32_t roff = (uint32_t)((ky >> 1) * TP + (kx >> 1));
        const uint64_t ad = desc_kmajor(smem_u32(sP + par * kTuPlaneBytes) + roff * 16, kTuPlaneRows);
        const uint64_t bd = bd0 + (uint64_t)(tap * (10 * 80 * 16 >> 4));
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
          umma_ss(tmem, ad + (uint64_t)(ks * 2 * kTuPlaneRows), bd + (uint64_t)(ks * 2 * 80), idesc, (tap | ks) != 0);
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, n_mma & 1);
    ++n_mma;
    tc_fence_after();
    const long next = tile + gridDim.x;
    if (next < a.ntiles) issue_tile(next);  // the MMAs that read the plane tiles have retired
    {
      const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
      const int m = (warp & 3) * 32 + lane, r = m >> tp_shift, c = m & (TP - 1);
      const int vr = tr * TR + r, J = tc * (TP - 1) + c;
      const int b = vr / hp, I = vr - b * hp;
      const bool ok = c <= TP - 2 && b < a.B && I < a.h && J < a.w;
      const long P = ok ? ((long)b * a.h + I) * a.w + J : 0;
      const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      const int g0 = warp < 4 ? 0 : 3, g1 = warp < 4 ? 3 : 5;
      for (int g = g0; g < g1; ++g) {
        float v[16];
        tmem_ld16(tacc + 16 * g, v);
        if (ok) st256(a.dact + dact_index(a.npx_in, P, a.in_off + 16 * g), pack8(v), pack8(v + 8));
      }
    }
    tc_fence_before();
  }
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// taps of the three tap groups (blockIdx.y): {00: (0,0) (0,2) (2,0)}, {00: (2,2); 01: (0,1) (2,1)}, {10: (1,0) (1,2); 11: (1,1)}
__constant__ int kTuWgTaps[3][3] = {{0, 2, 6}, {8, 1, 7}, {3, 5, 4}};

__global__ void __launch_bounds__(256, 1) tu_wgrad_kernel(TuBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TP = a.TP, TR = 128 / TP, tp_shift = 31 - __clz(TP);
  const int hp = a.h + 1;
  const int nwin = (TR + 1) * TP;
  const int grp = blockIdx.y;
  int taps[3], pars[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    taps[t] = kTuWgTaps[grp][t];
    pars[t] = ((taps[t] / 3) & 1) * 2 + ((taps[t] % 3) & 1);
  }
  const int par0 = pars[0], par1 = pars[2];  // plane slots 0 and 1 (equal for group 0: one plane)
  if (tid == 0) {
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < (int)(kTuWgradSmem / 16); i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  auto issue_tile = [&](long tile, uint8_t* buf) {
    const int tr = (int)(tile / a.ncolt), tc = (int)(tile - (long)tr * a.ncolt);
    const int vr0 = tr * TR, J0 = tc * (TP - 1);
    for (int p = tid; p < 128 * 10; p += 256) {  // x: K row rr = r * TP + c <-> input pixel (vr0 + r, J0 + c); halo column is zero
      const int rr = p / 10, jc = p - rr * 10;
      const int c = rr & (TP - 1), vr = vr0 + (rr >> tp_shift), J = J0 + c;
      const int b = vr / hp, I = vr - b * hp;
      const bool ok = c <= TP - 2 && b < a.B && I < a.h && J < a.w;
      cp_async16(buf + ((size_t)jc * kTuXRows + rr) * 16, ok ? a.x + ((size_t)((long)b * a.h + I) * a.w + J) * a.ldx + a.x_off + jc * 8 : a.x, ok);
    }
    tu_issue_plane(a, par0, vr0, J0, TP, tp_shift, nwin, buf + kTuXBytes, tid, 256);
    if (par1 != par0) tu_issue_plane(a, par1, vr0, J0, TP, tp_shift, nwin, buf + kTuXBytes + kTuPlaneBytes, tid, 256);
    cp_async_commit();
  };

  const uint32_t idesc = umma_idesc_bf16(80, 1, 1);
  long tile = blockIdx.x;
  int it = 0;
  if (tile < a.ntiles) issue_tile(tile, smem);
  for (; tile < a.ntiles; tile += gridDim.x, ++it) {
    const long next = tile + gridDim.x;
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();
    if (next < a.ntiles) {  // after the fence (which would wait for these copies), into the buffer the MMAs of tile it-1 read
      if (it >= 1) mbar_wait(&bar_mma[(it + 1) & 1], ((it - 1) >> 1) & 1);
      issue_tile(next, smem + ((it + 1) & 1) * kTuWgBuf);
    }
    if (tid == 0) {
      tc_fence_after();
      uint8_t* buf = smem + (it & 1) * kTuWgBuf;
      const uint32_t xa = smem_u32(buf);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int ky = taps[t] / 3, kx = taps[t] % 3;
        const uint32_t roff = (uint32_t)((ky >> 1) * TP + (kx >> 1));
        const uint32_t pa = smem_u32(buf + kTuXBytes + (pars[t] == par0 ? 0 : kTuPlaneBytes)) + roff * 16;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem + t * 80, desc_mnmajor(xa + k * 256, kTuXRows), desc_mnmajor(pa + k * 256, kTuPlaneRows), idesc, (it | k) != 0);
      }
      umma_commit(&bar_mma[it & 1]);
    }
  }
  if (it > 0) {
    mbar_wait(&bar_mma[(it - 1) & 1], ((it - 1) >> 1) & 1);
    tc_fence_after();
    if (warp < 4) {  // accumulator lane = input channel
      const uint32_t tacc = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int t = 0; t < 3; ++t)
        for (int g = 0; g < 5; ++g) {
          float v[16];
          tmem_ld16(tacc + t * 80 + 16 * g, v);
          if (tid < 80) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(a.gw + ((size_t)tid * 80 + 16 * g + i) * 9 + taps[t], v[i]);
          }
        }
    }
  }
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int launch_tu_fwd(const TuFwdArgs& a, int grid, cudaStream_t st) {
  if (a.TP != 8 && a.TP != 16 && a.TP != 32) return rl_fail(RL_ERR_ARG, "tu_fwd: bad tile pitch");
  if ((a.o_off & 15) || (a.a_off & 7)) return rl_fail(RL_ERR_ARG, "tu_fwd: unaligned channel offsets");
  {
    ProfScope _ps(KC_GEMM_PX, st);
    tu_fwd_kernel<<<grid, kTuThreads, kTuFwdSmem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int launch_tu_dgrad(const TuBwdArgs& a, int grid, cudaStream_t st) {
  if (a.TP != 8 && a.TP != 16 && a.TP != 32) return rl_fail(RL_ERR_ARG, "tu_dgrad: bad tile pitch");
  if (a.in_off & 15) return rl_fail(RL_ERR_ARG, "tu_dgrad: unaligned channel offset");
  {
    ProfScope _ps(KC_GEMM_PX, st);
    tu_dgrad_kernel<<<grid, 256, kTuDgradSmem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int launch_tu_wgrad(const TuBwdArgs& a, int grid_x, cudaStream_t st) {
  if (a.TP != 8 && a.TP != 16 && a.TP != 32) return rl_fail(RL_ERR_ARG, "tu_wgrad: bad tile pitch");
  if (a.x_off & 7) return rl_fail(RL_ERR_ARG, "tu_wgrad: unaligned channel offset");
  {
    ProfScope _ps(KC_WGRAD, st);
    tu_wgrad_kernel<<<dim3(grid_x, 3), 256, kTuWgradSmem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int configure_tu_kernels() {
  RL_CUDA(cudaFuncSetAttribute(tu_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTuDgradSmem));
  RL_CUDA(cudaFuncSetAttribute(tu_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTuWgradSmem));
  RL_CUDA(cudaFuncSetAttribute(tu_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTuFwdSmem));
  return RL_OK;
}

}  // namespace rl
