// Library-wide entry points: version, error string, device gate (sm_100 only, no fallback).
#include "common.cuh"

namespace xmve {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

namespace {
struct DeviceInfo {
  int checked = 0;   // 0 = not yet, 1 = ok, -1 = rejected
  int sms = 0;
};
DeviceInfo g_info[64];

int probe(int dev) {
  if (dev < 0 || dev >= 64) return fail(XMVE_ERR_DEVICE, "device ordinal %d out of range", dev);
  DeviceInfo& di = g_info[dev];
  if (di.checked == 0) {
    int major = 0, minor = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return fail(XMVE_ERR_DEVICE, "cannot query CUDA device %d (no GPU?) -- libxmve has no CPU path", dev);
    }
    di.sms = sms;
    di.checked = (major == 10) ? 1 : -1;
    if (di.checked < 0)
      return fail(XMVE_ERR_DEVICE, "device %d is sm_%d%d; libxmve is built for sm_100a only", dev, major, minor);
  }
  if (di.checked < 0) return fail(XMVE_ERR_DEVICE, "device %d is not sm_100", dev);
  return XMVE_OK;
}
}  // namespace

int require_sm100() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(XMVE_ERR_DEVICE, "no CUDA device available -- libxmve has no CPU path");
  }
  return probe(dev);
}

int sm_count() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || probe(dev) != XMVE_OK) return -1;
  return g_info[dev].sms;
}

}  // namespace xmve

extern "C" int xmve_version(void) { return 100; }   // 0.1.0

extern "C" const char* xmve_last_error(void) { return xmve::last_error_buf(); }

extern "C" int xmve_device_check(int device) {
  if (device < 0) return xmve::require_sm100();
  return xmve::probe(device);
}

extern "C" int xmve_sm_count(void) {
  int n = xmve::sm_count();
  return n > 0 ? n : XMVE_ERR_DEVICE;
}
