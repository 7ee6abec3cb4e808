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

#define RL_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return rl::rl_fail(RL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
