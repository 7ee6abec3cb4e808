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

extern "C" const char* rl_last_error(void) { return rl::g_err; }
extern "C" int rl_abi_version(void) { return 1; }
