// Library-level entry points: version, thread-local error string, device probe.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kfac {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
}  // namespace kfac

extern "C" int kfac_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char* kfac_last_error(void) { return kfac::g_err; }
extern "C" int kfac_device_arch(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    kfac::set_error("no CUDA device available");
    return KFAC_ERR_CUDA;
  }
  return p.major * 10 + p.minor;
}

extern "C" long long kfac_launch_count(void) { return __atomic_load_n(&kfac::g_launches, __ATOMIC_RELAXED); }
