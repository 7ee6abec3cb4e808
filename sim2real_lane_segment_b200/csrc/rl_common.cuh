// rl_common.cuh — error plumbing shared by every translation unit of librightlane.so.
// The C ABI never throws and never aborts: every entry point returns an int status and leaves a
// thread-local message retrievable through rl_last_error() (include/rightlane.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "rightlane.h"

namespace rl {
int rl_fail(int code, const char* fmt, ...);
}

namespace rl {
// kernel classes for launch counting and the optional per-class CUDA-event timing used by bench.py's roofline
enum KernelClass { KC_DENSE_FWD = 0, KC_GEMM_PX, KC_DGRAD_DENSE, KC_DGRAD_TD, KC_WGRAD, KC_HEAD, KC_DZ_PREPARE,
                   KC_STATS, KC_FIRST_FWD, KC_TU_COL2IM, KC_PACK, KC_OPT, KC_MISC, KC_DGRAD_BULK, KC_WGRAD_DENSE, KC_COUNT };
struct ProfScope {
  int cls;
  cudaStream_t st;
  int slot;
  ProfScope(int cls, cudaStream_t st);
  ~ProfScope();
};
}  // namespace rl

#define RL_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return rl::rl_fail(RL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
