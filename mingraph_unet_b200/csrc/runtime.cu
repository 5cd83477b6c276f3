// Host-side runtime bits of the C ABI: error text, version, launch counter, device query.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace mg {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MG_ERR_CUDA;
  }
  return MG_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace mg

extern "C" {
int mg_version(void) { return MG_VERSION; }
const char* mg_last_error(void) { return mg::g_err; }
int64_t mg_launch_count(void) { return mg::g_launches.load(std::memory_order_relaxed); }
}
