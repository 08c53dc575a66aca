// Shared host-side helpers for libxmve (error reporting, argument checks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/xmve.h"

namespace xmve {

char* last_error_buf();   // thread-local, 512 bytes (api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int require_sm100();      // XMVE_OK or XMVE_ERR_DEVICE (cached per device; api.cu)
int sm_count();

// Function attributes (opt-in shared memory) and __device__ symbol addresses belong to ONE device: every cache of
// them is an array indexed by the current device ordinal, like the device gate in api.cu.
constexpr int MAX_DEVICES = 64;
inline int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) {
    cudaGetLastError();
    return -1;
  }
  return dev;
}

#define XMVE_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::xmve::fail(XMVE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

#define XMVE_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) return ::xmve::fail(XMVE_ERR_ARG, __VA_ARGS__);                           \
  } while (0)

#define XMVE_DEVICE_OR_RETURN()                                                            \
  do {                                                                                     \
    int _s = ::xmve::require_sm100();                                                      \
    if (_s != XMVE_OK) return _s;                                                          \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(XMVE_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return XMVE_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace xmve
