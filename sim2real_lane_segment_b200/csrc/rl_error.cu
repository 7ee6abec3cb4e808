// rl_error.cu — thread-local last-error string behind rl_last_error().
#include <stdarg.h>
#include <string.h>

#include "rl_common.cuh"

namespace rl {
static thread_local char g_err[512] = "";
int rl_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace rl

// ------------------------------------------------------------------------------------------------
// launch accounting + optional event timing (single host thread drives a plan; see include/rightlane.h)
// ------------------------------------------------------------------------------------------------
namespace rl {
static long g_launches[KC_COUNT] = {0};
static bool g_prof_on = false;
struct EvPair { cudaEvent_t a, b; int cls; };
static EvPair g_ev[8192];
static int g_ev_created = 0, g_ev_used = 0;

ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), slot(-1) {
  ++g_launches[c];
  if (g_prof_on && g_ev_used < 8192) {
    if (g_ev_used >= g_ev_created) {
      cudaEventCreate(&g_ev[g_ev_created].a);
      cudaEventCreate(&g_ev[g_ev_created].b);
      ++g_ev_created;
    }
    slot = g_ev_used++;
    g_ev[slot].cls = c;
    cudaEventRecord(g_ev[slot].a, st);
  }
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_ev[slot].b, st);
}
}  // namespace rl

extern "C" long rl_launch_count(int kernel_class) {
  if (kernel_class < 0) {
    long t = 0;
    for (int i = 0; i < rl::KC_COUNT; ++i) t += rl::g_launches[i];
    return t;
  }
  return kernel_class < rl::KC_COUNT ? rl::g_launches[kernel_class] : 0;
}
extern "C" int rl_profile_begin(void) {
  rl::g_prof_on = true;
  rl::g_ev_used = 0;
  return 0;
}
// ms[KC_COUNT], launches[KC_COUNT]: summed CUDA-event time and launch count per kernel class since rl_profile_begin
extern "C" int rl_profile_end(double* ms, long* launches) {
  rl::g_prof_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return rl::rl_fail(RL_ERR_CUDA, "rl_profile_end: sync failed");
  for (int i = 0; i < rl::KC_COUNT; ++i) {
    ms[i] = 0;
    launches[i] = 0;
  }
  for (int i = 0; i < rl::g_ev_used; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, rl::g_ev[i].a, rl::g_ev[i].b);
    ms[rl::g_ev[i].cls] += t;
    launches[rl::g_ev[i].cls] += 1;
  }
  return 0;
}

// per-launch times of the last profiled region (call after rl_profile_end): fills up to max entries, returns the count
extern "C" int rl_profile_launches(float* ms, int* cls, int max) {
  int n = rl::g_ev_used < max ? rl::g_ev_used : max;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, rl::g_ev[i].a, rl::g_ev[i].b);
    ms[i] = t;
    cls[i] = rl::g_ev[i].cls;
  }
  return n;
}

extern "C" const char* rl_last_error(void) { return rl::g_err; }
extern "C" int rl_abi_version(void) { return 1; }
