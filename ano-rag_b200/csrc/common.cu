// Error reporting, launch accounting and device queries shared by all entry points.
#include "common.cuh"

#include <mutex>
#include <vector>

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

struct ProfRec { int kind; cudaEvent_t e0, e1; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof;

ProfileScope::ProfileScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_), on(g_prof_on && g_prof.size() < 65536) {
  if (!on) return;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { on = false; return; }
  cudaEventRecord(e0, st);
}
ProfileScope::~ProfileScope() {
  if (!on) return;
  cudaEventRecord(e1, st);
  g_prof.push_back({kind, e0, e1});
}

// Keep freed scratch in the stream-ordered pool instead of returning it to the driver at
// every synchronisation (the default release threshold is 0): per-call scratch such as the
// dense candidate lists would otherwise be re-mapped on every search.
void configure_pool_once() {
  static std::once_flag once[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  std::call_once(once[dev], [dev] {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  });
}

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

extern "C" int arb_profile(int enable) {
  arb::g_prof_on = enable != 0;
  return ARB_OK;
}

extern "C" int arb_profile_collect(double* out_ms, int64_t* out_count) {
  for (int i = 0; i < arb::PROF_KINDS; ++i) { if (out_ms) out_ms[i] = 0.0; if (out_count) out_count[i] = 0; }
  int rc = ARB_OK;
  for (auto& r : arb::g_prof) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (e != cudaSuccess) rc = arb::cuda_fail(e, "arb_profile_collect", __FILE__, __LINE__);
    else { if (out_ms) out_ms[r.kind] += ms; if (out_count) out_count[r.kind] += 1; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  arb::g_prof.clear();
  return rc;
}
