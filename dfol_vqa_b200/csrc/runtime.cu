// Error reporting and version of libdfol_b200.
#include <stdarg.h>

#include "dfol_common.cuh"

namespace dfol {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace dfol

extern "C" int dfol_version(void) { return DFOL_ABI_VERSION; }
extern "C" const char* dfol_last_error(void) { return dfol::g_error; }
