// Error plumbing and library-level entry points of libnlb200.
#include <cstdarg>
#include <cstdio>
#include <cmath>
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

// ---------------------------------------------------------------- CUDA-graph support
// Scalars that change every training step (see include/nlb200.h).  The pointer is host state of the CALLING
// THREAD (two trainers capturing on two threads must not see each other's buffer); kernels receive it as an
// argument at launch / capture time and read the values at execution time.
static thread_local const float* g_dyn = nullptr;
const float* nlb_dynamic_scalars() { return g_dyn; }

extern "C" int nlb_set_dynamic_scalars(const float* dev) {
  g_dyn = dev;
  return NLB_OK;
}

extern "C" int nlb_adam_bias_terms(float lr, float beta1, float beta2, int step, float* out2) {
  if (!out2 || step < 1) { nlb_set_error("adam_bias_terms: bad arguments"); return NLB_EINVAL; }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  out2[0] = (float)((double)lr / bc1);
  out2[1] = (float)(1.0 / sqrt(bc2));
  return NLB_OK;
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
int nlb_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}
