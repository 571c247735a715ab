// ABI plumbing: version, error string, device probe.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>

namespace dpot {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
int g_pdl = 0;
int g_sm_budget = 0;     // dpot_set_sm_budget: > 0 caps the SM count the persistent kernels size their grids for
int sm_count_cur() {
  static int cache[64] = {0};
  const int d = cur_dev();
  if (cache[d] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
    cache[d] = n;
  }
  int n = cache[d];
  if (g_sm_budget > 0 && g_sm_budget < n) n = g_sm_budget & ~1;    // even: CTA pairs
  return n < 2 ? 2 : n;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}
}  // namespace dpot

extern "C" int dpot_abi_version(void) { return DPOT_ABI_VERSION; }
extern "C" const char* dpot_last_error_string(void) { return dpot::g_err; }

extern "C" long long dpot_launch_count(void) { return dpot::g_launches.load(); }

extern "C" int dpot_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

extern "C" void dpot_set_pdl(int32_t on) { dpot::g_pdl = on ? 1 : 0; }

// SM budget of the persistent kernels (contractions, fused mixer, tail / PatchEmbed backward): n > 0 makes them size
// their grids for n SMs instead of the whole device, leaving the rest to kernels that run CONCURRENTLY on other streams
// (NCCL's all-reduce during backward).  0 = the whole device (default).  Returns the previous value; n < 0 only queries.
extern "C" int dpot_set_sm_budget(int32_t n) {
  const int prev = dpot::g_sm_budget;
  if (n >= 0) dpot::g_sm_budget = n;
  return prev;
}
