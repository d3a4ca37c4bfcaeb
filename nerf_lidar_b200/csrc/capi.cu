// Error plumbing and library-level entry points of libnlb200.
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
#include "../../include/nlb200.h"

static thread_local char g_err[512] = "";

void nlb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int nlb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    nlb_set_error("%s: %s", what, cudaGetErrorString(e));
    return NLB_ECUDA;
  }
  return NLB_OK;
}

extern "C" const char* nlb_last_error(void) { return g_err; }

extern "C" int nlb_version(void) { return 100; }

extern "C" int nlb_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p.major == 10 ? 1 : 0;
}
