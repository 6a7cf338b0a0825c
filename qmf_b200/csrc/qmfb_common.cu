#include "qmfb_common.h"

namespace qmfb {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace qmfb

extern "C" {

const char* qmfb_last_error(void) { return qmfb::g_error; }

int qmfb_version(void) { return 100; }

int qmfb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return qmfb::set_error(QMFB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  return n;
}

}  // extern "C"
