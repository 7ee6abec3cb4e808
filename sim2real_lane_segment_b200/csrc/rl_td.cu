// rl_td.cu — TransitionDown forward (reference layers.py:43-55): BN -> ReLU -> Conv1x1 (+bias, Dropout2d scale) -> MaxPool2d(2)
// as one warp-specialised kernel, one CTA per SM:
//   warps 0-3  loaders: cp.async of the input pixels (rows in pooling order) into a 4-stage ring of K-major operand tiles
//   warps 4-7  BN + ReLU in place (thread = row)
//   warp  8    MMA issue: the part's whole packed weight matrix is resident in shared memory (one bulk TMA per CTA)
//   warps 9-16 epilogue (two per TMEM lane quarter, alternate 16-column blocks): the accumulator rows are ordered so that the four positions of a pooling window are the four TMEM
//              lanes that two tcgen05.ld.16x256b hand to ONE thread; max / arg-max / bias / dropout scale / stores per
//              thread; two accumulator buffers, so tile t+1 is computed while tile t is pooled
// gemm_px_kernel (rl_conv_fwd.cu), the single-role predecessor, remains for parts whose weights do not fit next to the ring.
#include "rl_common.cuh"
#include "rl_dev.cuh"
#include "rl_kernels.h"

namespace rl {

constexpr int kTdThreads = 17 * 32;
constexpr int kTdRows = 129;                       // odd row pitch: the producers' 16-byte stores of different chunks spread over the banks
constexpr int kTdStageBytes = 8 * kTdRows * 16;    // one 128-row x 64-channel operand tile, chunk layout T[j][row][8]
constexpr int kTdStages = 4;

size_t td_fwd_smem(int K, int np) {
  const int nkc = (K + 63) / 64;
  return (size_t)nkc * 8 * np * 16 + (size_t)kTdStages * kTdStageBytes + (size_t)((K + 63) & ~63) * 8 + 2 * 256 * 4 + 128;
}

__global__ void __launch_bounds__(kTdThreads, 1) td_fwd_kernel(GemmPxArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int part = blockIdx.y;
  const int n0 = a.part_n0[part], np = a.part_np[part];
  const int nkc = (a.K + 63) / 64;
  const int stageB = 8 * np * 16;
  uint8_t* sW = smem;
  uint8_t* sA = sW + (size_t)nkc * stageB;
  float* s_s = reinterpret_cast<float*>(sA + kTdStages * kTdStageBytes);
  float* s_t = s_s + ((a.K + 63) & ~63);
  float* s_bias = s_t + ((a.K + 63) & ~63);
  float* s_mask = s_bias + 256;  // Dropout2d scale of the image the epilogue is working on
  __shared__ uint64_t bar_w, bar_raw[kTdStages], bar_full[kTdStages], bar_empty[kTdStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < kTdStages; ++i) {
      mbar_init(&bar_raw[i], 128);  // one asynchronous arrival per loader thread (cp.async.mbarrier.arrive.noinc)
      mbar_init(&bar_full[i], 4);   // one arrival per transform warp
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  for (int c = tid; c < a.K; c += kTdThreads) {
    BnCoef k = bn_coef(a.bn, c);
    s_s[c] = k.s;
    s_t[c] = k.t;
  }
  for (int c = tid; c < 256; c += kTdThreads) {
    s_bias[c] = (a.bias != nullptr && c < np) ? a.bias[n0 + c] : 0.f;
    s_mask[c] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const long ntiles_cta = a.ntiles > (long)blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // row m of tile t -> arena pixel (or -1): m = 32 * (win / 8) + 8 * q + win % 8 (see the epilogue)
  auto pixel_of = [&](long t, int m) -> long {
    const int pr_ = (int)(t / a.wblocks), wb_ = (int)(t - (long)pr_ * a.wblocks);
    const int win = ((m >> 5) << 3) | (m & 7), q = (m >> 3) & 3;
    const int wi = wb_ * 32 + win;
    if (wi >= a.Wp) return -1;
    const int img = pr_ / a.Hp, py = pr_ - img * a.Hp;
    return ((long)img * a.H + 2 * py + (q >> 1)) * a.W + 2 * wi + (q & 1);
  };

  if (warp < 4) {
    // =============================================================== loaders: cp.async of the raw bf16 rows, several stages
    // ahead (asynchronous copies are not held back by anybody's fence.proxy.async; register-staged loads were: every
    // published stage waited for the loads of the stages prefetched behind it, and the loop ran at one DRAM latency per stage)
    const int j = tid & 7;  // 8-channel chunk of the 64-channel stage
    int roff[8], rwin[8];   // element offset of row r = (tid >> 3) + 16 i inside the tile's 2 x 64 pixel patch, and its window
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (tid >> 3) + 16 * i;
      const int win = ((r >> 5) << 3) | (r & 7), q = (r >> 3) & 3;
      rwin[i] = win;
      roff[i] = ((q >> 1) * a.W + 2 * win + (q & 1)) * a.lda + j * 8;
    }
    uint32_t s = 0, use = 0;
    for (long it = 0; it < ntiles_cta; ++it) {
      const long tile = blockIdx.x + it * gridDim.x;
      const int pr = (int)(tile / a.wblocks), wb = (int)(tile - (long)pr * a.wblocks);
      const int img = pr / a.Hp, py = pr - img * a.Hp;
      const __nv_bfloat16* base = a.A + ((size_t)((long)img * a.H + 2 * py) * a.W + 64 * wb) * a.lda + a.a_off;
      const int wleft = a.Wp - wb * 32;  // windows of this tile that exist
      for (int kc = 0; kc < nkc; ++kc) {
        if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1);  // the MMAs that read this stage have retired
        uint8_t* dst = sA + s * kTdStageBytes + ((size_t)j * kTdRows + (tid >> 3)) * 16;
        const bool jon = kc * 64 + 8 * j < a.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = jon && rwin[i] < wleft;
          cp_async16(dst + i * 16 * 16, ok ? base + roff[i] + kc * 64 : a.A, ok);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_raw[s])) : "memory");
        if (++s == kTdStages) {
          s = 0;
          ++use;
        }
      }
    }
  } else if (warp < 8) {
    // =============================================================== BN + ReLU in place: thread = row, chunk by chunk
    const int row = tid - 128;
    uint32_t s = 0, use = 0;
    const long total = ntiles_cta * nkc;
    int kc = 0;
    for (long g = 0; g < total; ++g) {
      mbar_wait(&bar_raw[s], use & 1);
      uint8_t* st = sA + s * kTdStageBytes + (size_t)row * 16;
      const int nj = min(8, (a.K - kc * 64) / 8);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (jj < nj) {
          uint4* p = reinterpret_cast<uint4*>(st + (size_t)jj * kTdRows * 16);
          *p = bn_relu8(*p, s_s + kc * 64 + 8 * jj, s_t + kc * 64 + 8 * jj);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
      if (++kc == nkc) kc = 0;
      if (++s == kTdStages) {
        s = 0;
        ++use;
      }
    }
  } else if (warp == 8) {
    // =============================================================== MMA issue
    if (lane == 0 && ntiles_cta > 0) {
      mbar_arrive_expect_tx(&bar_w, (uint32_t)(nkc * stageB));
      bulk_g2s(sW, a.wpk + (size_t)nkc * 8 * 16 * n0, (uint32_t)(nkc * stageB), &bar_w);
      const uint32_t idesc = umma_idesc_bf16((uint32_t)np, 0, 0);
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
      mbar_wait(&bar_w, 0);
      uint32_t s = 0, use = 0;
      for (long it = 0; it < ntiles_cta; ++it) {
        const uint32_t buf = (uint32_t)it & 1;
        if (it >= 2) {
          mbar_wait(&acc_empty[buf], ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int kc = 0; kc < nkc; ++kc) {
          const int n16 = min(4, (a.K - kc * 64) / 16);
          mbar_wait(&bar_full[s], use & 1);
          tc_fence_after();
          const uint32_t aa = a0 + s * kTdStageBytes, wa = w0 + kc * stageB;
          for (int kk = 0; kk < n16; ++kk)
            umma_ss(tmem + buf * 256, desc_kmajor(aa + kk * 2 * 16 * kTdRows, kTdRows), desc_kmajor(wa + kk * 2 * 16 * np, np), idesc,
                    (kc | kk) != 0);
          umma_commit(&bar_empty[s]);
          if (++s == kTdStages) {
            s = 0;
            ++use;
          }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // =============================================================== epilogue (warps 9-12; TMEM lane quarter = warp % 4)
    // thread (quarter, lane) owns window 8 * quarter + lane / 4 and, in every 16-column block, the four columns
    // 8k + 2(lane % 4) + e; the packed weights permute the output channels so that those are channels 4(lane % 4) + 2k + e
    // of the block (rl_misc.cu pack kind 2): 8 contiguous bytes per thread, 32 per quad.  +bias and the non-negative
    // Dropout2d scale are the same for the four pixels of a window, so they commute with the max.
    const int qt = warp & 3, sub = (warp - 9) >> 2;
    const int win = qt * 8 + (lane >> 2), qc = lane & 3;
    int mimg = -1;
    for (long it = 0; it < ntiles_cta; ++it) {
      const long tile = blockIdx.x + it * gridDim.x;
      const uint32_t buf = (uint32_t)it & 1;
      const int pr = (int)(tile / a.wblocks), wb = (int)(tile - (long)pr * a.wblocks);
      const int wi = wb * 32 + win;
      const int img = pr / a.Hp, py = pr - img * a.Hp;
      const bool valid = wi < a.Wp;
      const size_t opx = ((size_t)img * a.Hp + py) * a.Wp + wi;
      if (a.mask != nullptr && img != mimg) {  // the eight epilogue warps walk the same tile sequence: swap the scale row together
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = tid - 9 * 32; c < np; c += 256) s_mask[c] = a.mask[(size_t)img * a.N + n0 + c];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mimg = img;
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem + buf * 256 + ((uint32_t)(qt * 32) << 16);
      for (int c0 = 16 * sub; c0 < np; c0 += 32) {
        float r0[8], r1[8];
        tmem_ld_16x256b_x2(lane_base + c0, r0);                  // positions 0 (regs 4k+e) and 1 (regs 4k+2+e)
        tmem_ld_16x256b_x2(lane_base + (16u << 16) + c0, r1);    // positions 2 and 3
        tmem_wait_ld();
        reg_fence8(r0);
        reg_fence8(r1);
        const int cl = c0 + 4 * qc;  // first of this thread's four channels (part-local)
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cl);
        const float4 m4 = *reinterpret_cast<const float4*>(s_mask + cl);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// true when every part of this TransitionDown fits the resident-weight kernel
bool td_fwd_fits(int K, const int* part_np, int nparts) {
  for (int i = 0; i < nparts; ++i)
    if (td_fwd_smem(K, part_np[i]) > 226 * 1024 || part_np[i] > 256) return false;
  return true;
}

int launch_td_fwd(const GemmPxArgs& a, int grid_x, int nparts, cudaStream_t st) {
  int npmax = 0;
  for (int i = 0; i < nparts; ++i) npmax = a.part_np[i] > npmax ? a.part_np[i] : npmax;
  const size_t smem = td_fwd_smem(a.K, npmax);
  if (smem > 226 * 1024) return rl_fail(RL_ERR_SHAPE, "td_fwd: smem %zu too large (K=%d np=%d)", smem, a.K, npmax);
  {
    ProfScope _ps(KC_GEMM_PX, st);
    td_fwd_kernel<<<dim3(grid_x, nparts), kTdThreads, smem, st>>>(a);
  }
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

int configure_td_kernels() {
  RL_CUDA(cudaFuncSetAttribute(td_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  return RL_OK;
}

}  // namespace rl
