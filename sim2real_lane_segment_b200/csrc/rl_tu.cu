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
constexpr int kTuTileRows = 176;                     // (TR + 1) * TP + guard rows <= 5 * 32 + 2
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
    for (int p = tid; p < nwin * 10; p += kTuThreads) {
      const int jc = p / nwin, rr = p - jc * nwin;
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
    if (next < a.ntiles) issue_tile(next, sX + ((it + 1) & 1) * kTuTileBytes);  // its last readers (MMAs of tile it-1) retired
    else cp_async_commit();
    cp_async_wait<1>();
    fence_proxy_async();
    __syncthreads();  // tile `it` is in shared memory; every thread is past the epilogue of tile it-1 (TMEM reusable)
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

int configure_tu_kernels() {
  RL_CUDA(cudaFuncSetAttribute(tu_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTuFwdSmem));
  return RL_OK;
}

}  // namespace rl
