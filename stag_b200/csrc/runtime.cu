// Library-wide runtime helpers: thread-local error string, device properties, version.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace stag {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace stag

extern "C" const char* stag_last_error(void) { return stag::g_err; }
extern "C" int stag_abi_version(void) { return STAG_ABI_VERSION; }
extern "C" long long stag_launch_count(void) { return stag::g_launches.load(std::memory_order_relaxed); }
extern "C" int stag_hub_threshold(void) { return stag::kHubThreshold; }
extern "C" int stag_hub_segment(void) { return stag::kHubSegment; }
