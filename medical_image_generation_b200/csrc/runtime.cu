// Error plumbing and the per-device attribute cache (the only global state in the library).
#include <mutex>

#include "common.cuh"

namespace mig {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

const DeviceInfo& device_info() {
  static DeviceInfo info[16];
  static bool init[16] = {};
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  if (!init[dev]) {
    std::lock_guard<std::mutex> lock(mu);
    if (!init[dev]) {
      DeviceInfo d{148, 0, 0, 0};
      cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
      cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
      cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      if (d.sm_count <= 0) d.sm_count = 148;
      info[dev] = d;
      init[dev] = true;
    }
  }
  return info[dev];
}

}  // namespace mig

extern "C" const char* mig_last_error(void) { return mig::g_err; }
extern "C" int mig_abi_version(void) { return MIG_ABI_VERSION; }
extern "C" int mig_has_tcgen05(void) {
  const mig::DeviceInfo& d = mig::device_info();
  return d.cc_major == 10 ? 1 : 0;
}
