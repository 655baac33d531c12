// api.cu - version, error string and device query of the C ABI (include/admmq.h).
#include <atomic>
#include <cstring>
#include "common.cuh"

namespace admmq {

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int device_props(DeviceProps* out) {
  static thread_local DeviceProps cache;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ADMMQ_E_CUDA, "no usable CUDA device: %s (libadmmq has no CPU fallback)", cudaGetErrorString(e));
  }
  if (cache.device != dev) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ADMMQ_E_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    }
    cache.device = dev;
    cache.sm_count = prop.multiProcessorCount;
    cache.cc_major = prop.major;
    cache.cc_minor = prop.minor;
    cache.coop = prop.cooperativeLaunch;
    cache.smem_optin = prop.sharedMemPerBlockOptin;
  }
  *out = cache;
  return ADMMQ_OK;
}

}  // namespace admmq

extern "C" int admmq_version(void) { return ADMMQ_VERSION; }

extern "C" uint64_t admmq_launch_count(void) { return (uint64_t)admmq::launches_so_far(); }

extern "C" const char* admmq_last_error(void) { return admmq::error_buffer(); }

extern "C" int admmq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  admmq::DeviceProps dp;
  if (int e = admmq::device_props(&dp)) return e;
  if (sm_count) *sm_count = dp.sm_count;
  if (cc_major) *cc_major = dp.cc_major;
  if (cc_minor) *cc_minor = dp.cc_minor;
  return ADMMQ_OK;
}
