// Error reporting, launch accounting and device queries shared by all entry points.
#include "common.cuh"

#include <mutex>

namespace arb {

static thread_local std::string g_last_error;
static thread_local int64_t g_launches = 0;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") at " +
                 file + ":" + std::to_string(line) + " in " + what;
  cudaGetLastError();   // clear the sticky-less error so later calls report their own
  return e == cudaErrorMemoryAllocation ? ARB_ENOMEM : ARB_ECUDA;
}

void count_launch(int n) { g_launches += n; }

int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

}  // namespace arb

extern "C" int arb_abi_version(void) { return ARB_ABI_VERSION; }

extern "C" const char* arb_last_error(void) { return arb::g_last_error.c_str(); }

extern "C" int arb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return arb::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
  return n;
}

extern "C" int64_t arb_launch_count(int reset) {
  int64_t v = arb::g_launches;
  if (reset) arb::g_launches = 0;
  return v;
}
