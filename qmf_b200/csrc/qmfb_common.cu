#include "qmfb_common.h"

#include <cstring>

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

// ---- device buffers shareable between the one-process-per-GPU ranks of a box (CUDA IPC) ---------
int qmfb_ipc_alloc(int device, int64_t bytes, void** ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (bytes < 1 || !ptr || !handle64) return qmfb::set_error(QMFB_ERR_INVALID, "qmfb_ipc_alloc: bad argument");
  QMFB_CUDA(cudaSetDevice(device));
  QMFB_CUDA(cudaMalloc(ptr, size_t(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMemset(*ptr, 0, size_t(bytes));
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    return qmfb::set_error(QMFB_ERR_CUDA, "qmfb_ipc_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, sizeof h);
  return QMFB_OK;
}

int qmfb_ipc_open(int device, const void* handle64, void** ptr) {
  if (!ptr || !handle64) return qmfb::set_error(QMFB_ERR_INVALID, "qmfb_ipc_open: bad argument");
  QMFB_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  QMFB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QMFB_OK;
}

int qmfb_ipc_close(void* ptr) {
  if (ptr) QMFB_CUDA(cudaIpcCloseMemHandle(ptr));
  return QMFB_OK;
}

int qmfb_ipc_free(void* ptr) {
  if (ptr) QMFB_CUDA(cudaFree(ptr));
  return QMFB_OK;
}

}  // extern "C"
