// Library-wide C-ABI plumbing: error string, version, device query.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void ocrs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void ocrs_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

extern "C" {
// Kernels launched by this library since load (every entry point counts its own launches).
long long ocrs_launch_count(void) { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
const char* ocrs_last_error(void) { return g_err; }
int ocrs_version(void) { return 100; }
// Compute capability major*10+minor of the current device (100 on B200), or <0 on error.
int ocrs_device_arch(void) {
  int dev = 0, maj = 0, min = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return maj * 10 + min;
}
}
