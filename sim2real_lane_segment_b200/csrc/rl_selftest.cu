// rl_selftest.cu — single-CTA tcgen05 GEMM used by tests to pin the descriptor conventions of rl_umma.cuh
// (chunk layout as K-major / MN-major operand, row-shifted start addresses, 1-D bulk TMA operand fill).
#include "rl_common.cuh"
#include "rl_umma.cuh"

namespace rl {

// Copies a row-major bf16 matrix [rows × cols] (cols % 8 == 0) into the chunk layout T[cols/8][rows][8].
__device__ void fill_chunk_layout(uint8_t* smem, const __nv_bfloat16* g, int rows, int cols) {
  const int nchunk = cols >> 3;
  for (int i = threadIdx.x; i < rows * nchunk; i += blockDim.x) {
    int r = i % rows, j = i / rows;
    uint4 v = *reinterpret_cast<const uint4*>(g + (size_t)r * cols + j * 8);
    *reinterpret_cast<uint4*>(smem + ((size_t)j * rows + r) * 16) = v;
  }
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __nv_bfloat16* A, int a_rows, int a_cols, int a_mn, int a_shift,
                     const __nv_bfloat16* B, int b_rows, int b_cols, int b_mn, int b_shift,
                     int N, int K, float* D, int bulk_b) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (((size_t)a_rows * a_cols * 2 + 127) & ~(size_t)127);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  fill_chunk_layout(sA, A, a_rows, a_cols);
  if (bulk_b) {
    if (threadIdx.x == 0) {
      uint32_t bytes = (uint32_t)b_rows * b_cols * 2;
      mbar_arrive_expect_tx(&bar_tma, bytes);
      bulk_g2s(sB, B, bytes, &bar_tma);
    }
  } else {
    fill_chunk_layout(sB, B, b_rows, b_cols);
  }
  fence_proxy_async();
  __syncthreads();

  if (threadIdx.x == 0) {
    if (bulk_b) mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16((uint32_t)N, (uint32_t)a_mn, (uint32_t)b_mn);
    const uint32_t a0 = smem_u32(sA) + 16u * a_shift, b0 = smem_u32(sB) + 16u * b_shift;
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t ad = a_mn ? desc_mnmajor(a0 + kk * 256u, a_rows) : desc_kmajor(a0 + kk * 2u * 16u * a_rows, a_rows);
      uint64_t bd = b_mn ? desc_mnmajor(b0 + kk * 256u, b_rows) : desc_kmajor(b0 + kk * 2u * 16u * b_rows, b_rows);
      umma_ss(tmem_base, ad, bd, idesc, kk > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  const int row = threadIdx.x;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace rl

extern "C" int rl_selftest_umma(const void* A, int a_rows, int a_cols, int a_mn_major, int a_row_shift,
                                const void* B, int b_rows, int b_cols, int b_mn_major, int b_row_shift, int N,
                                int K, float* D, int bulk_b, void* stream) {
  using namespace rl;
  if (N % 16 || N < 16 || N > 256 || K % 16 || a_cols % 8 || b_cols % 8) return rl_fail(RL_ERR_ARG, "selftest: bad shape");
  size_t sa = (((size_t)a_rows * a_cols * 2 + 127) & ~(size_t)127), sb = (size_t)b_rows * b_cols * 2;
  size_t smem = sa + sb;
  if (smem > 200 * 1024) return rl_fail(RL_ERR_ARG, "selftest: tile too large");
  RL_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)A, a_rows, a_cols, a_mn_major, a_row_shift, (const __nv_bfloat16*)B, b_rows, b_cols,
      b_mn_major, b_row_shift, N, K, D, bulk_b);
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ------------------------------------------------------------------------------------------------
// Measurement hook: stream a channel window [c0, c0+nch) of an NHWC bf16 arena ([npx][Ct]) with 32-byte accesses,
// the access pattern of every consumer of the concat arenas.  mode 0: read only, 1: read + write back (RMW).
// bench tools use it to measure what the memory system delivers for row-strided windows.
// ------------------------------------------------------------------------------------------------
namespace rl {
__global__ void __launch_bounds__(256) stream_window_kernel(__nv_bfloat16* arena, long npx, int Ct, int c0, int nch, int mode,
                                                            unsigned* sink) {
  const int per_row = nch / 16;  // 32-byte pieces per pixel row
  const long total = npx * per_row;
  unsigned acc = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long p = i / per_row;
    const int q = (int)(i - p * per_row);
    __nv_bfloat16* ptr = arena + p * Ct + c0 + q * 16;
    u32x8 v = ldg256_nc_nohint(ptr);
    acc ^= v.lo.x ^ v.lo.y ^ v.lo.z ^ v.lo.w ^ v.hi.x ^ v.hi.y ^ v.hi.z ^ v.hi.w;
    if (mode == 1) st256(ptr, v.lo, v.hi);
  }
  if (acc == 0x12345678u) *sink = acc;  // keeps the loads alive
}
}  // namespace rl

extern "C" int rl_selftest_stream(void* arena, long npx, int Ct, int c0, int nch, int mode, int grid, void* sink, void* stream) {
  using namespace rl;
  if (nch % 16 || c0 % 16 || c0 + nch > Ct) return rl_fail(RL_ERR_ARG, "selftest_stream: bad window");
  stream_window_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)arena, npx, Ct, c0, nch, mode, (unsigned*)sink);
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ------------------------------------------------------------------------------------------------
// Raw descriptor probe: the caller supplies the exact shared-memory images of A and B and the upper bits of both
// matrix descriptors (everything except the start address), so a layout / descriptor hypothesis (swizzle modes,
// LBO / SBO meaning, K-step advance) can be checked against a reference product without recompiling.
// ------------------------------------------------------------------------------------------------
namespace rl {
__global__ void __launch_bounds__(128, 1)
umma_raw_kernel(const uint8_t* a_img, int a_bytes, unsigned long long a_tmpl, int a_kstep, const uint8_t* b_img, int b_bytes,
                unsigned long long b_tmpl, int b_kstep, uint32_t idesc, int nk, int N, float* D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (((size_t)a_bytes + 1023) & ~(size_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    for (int kk = 0; kk < nk; ++kk) {
      const uint64_t ad = a_tmpl | (uint64_t)(((smem_u32(sA) + (uint32_t)(kk * a_kstep)) >> 4) & 0x3FFFu);
      const uint64_t bd = b_tmpl | (uint64_t)(((smem_u32(sB) + (uint32_t)(kk * b_kstep)) >> 4) & 0x3FFFu);
      umma_ss(tmem, ad, bd, idesc, kk > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)threadIdx.x * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace rl

extern "C" int rl_selftest_umma_raw(const void* a_img, int a_bytes, unsigned long long a_desc_tmpl, int a_kstep_bytes,
                                    const void* b_img, int b_bytes, unsigned long long b_desc_tmpl, int b_kstep_bytes,
                                    unsigned idesc, int nk, int N, float* D, void* stream) {
  using namespace rl;
  if (a_bytes % 16 || b_bytes % 16 || N % 16 || N < 16 || N > 256) return rl_fail(RL_ERR_ARG, "selftest_raw: bad sizes");
  const size_t smem = 1024 + (((size_t)a_bytes + 1023) & ~(size_t)1023) + (size_t)b_bytes;
  if (smem > 200 * 1024) return rl_fail(RL_ERR_ARG, "selftest_raw: images too large");
  RL_CUDA(cudaFuncSetAttribute(umma_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_raw_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)a_img, a_bytes, a_desc_tmpl, a_kstep_bytes,
                                                          (const uint8_t*)b_img, b_bytes, b_desc_tmpl, b_kstep_bytes, idesc, nk, N, D);
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}

// ------------------------------------------------------------------------------------------------
// Pins the register fragment layout of tcgen05.ld.16x256b (used by the pooling epilogue of TransitionDown): every TMEM
// cell (lane, column) of a 128 x 32 block is written with the tag (lane << 8 | column) through tcgen05.st.32x32b, then each
// warp reads its lane quarter back with two 16x256b.x4 loads (lanes +0..15 and +16..31).  out[tid][half][16] = the tags each
// thread received.
// ------------------------------------------------------------------------------------------------
namespace rl {
__global__ void __launch_bounds__(128, 1) tmem_shape_probe_kernel(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = ((uint32_t)tid << 8) | (uint32_t)(c0 + i);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(lane_base + c0),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t v[16];
    tmem_ld_16x256b_x4(lane_base + ((uint32_t)(16 * half) << 16), v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(tid * 2 + half) * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}
}  // namespace rl

extern "C" int rl_selftest_tmem_shapes(unsigned int* out, void* stream) {
  using namespace rl;
  if (!out) return rl_fail(RL_ERR_ARG, "rl_selftest_tmem_shapes: null argument");
  tmem_shape_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(out);
  RL_CUDA(cudaGetLastError());
  return RL_OK;
}
