// Bandwidth-bound reductions of the bf16 pre-activation gradients the fused NerfMLP
// backward writes: bias gradients (column sums over all rows) and the per-ray sums that
// feed the view-direction columns of the weight gradients (the direction encoding is a
// per-ray constant broadcast over the samples, Z/internal/models.py:1192-1196).
// torch's generic reduce kernels took 0.75 ms per step on these; one pass at HBM speed is
// 0.1 ms (660 MB).
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

constexpr int kRedThreads = 256;

// x[M, ld] bf16 (first `cols` columns) -> out[cols] += column sums.  A thread owns one
// column pair and every (256 / pairs)-th row of the block's row range.
__global__ void __launch_bounds__(kRedThreads) k_colsum_bf16(const __nv_bfloat16* __restrict__ x, int64_t M, int cols,
                                                             int ld, float* __restrict__ out) {
  __shared__ float2 s_part[kRedThreads];
  const int pairs = cols >> 1;
  const int lanes = kRedThreads / pairs;  // row lanes per block
  const int p = threadIdx.x % pairs, rl = threadIdx.x / pairs;
  const int64_t per_block = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per_block;
  const int64_t r1 = r0 + per_block < M ? r0 + per_block : M;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int64_t r = r0 + rl; r < r1; r += lanes) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * ld + 2 * p));
    acc.x += v.x;
    acc.y += v.y;
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < pairs) {
    float2 t = make_float2(0.f, 0.f);
    for (int l = 0; l < lanes; ++l) {
      t.x += s_part[l * pairs + threadIdx.x].x;
      t.y += s_part[l * pairs + threadIdx.x].y;
    }
    atomicAdd(out + 2 * threadIdx.x, t.x);
    atomicAdd(out + 2 * threadIdx.x + 1, t.y);
  }
}

// x[groups * S, cols] bf16 -> out[groups, cols] fp32: sum over the S consecutive rows of a group.
__global__ void __launch_bounds__(kRedThreads) k_group_sum_bf16(const __nv_bfloat16* __restrict__ x, int64_t groups,
                                                                int S, int cols, int ld, float* __restrict__ out) {
  const int pairs = cols >> 1;
  const int per_block = kRedThreads / pairs;  // groups per block
  const int p = threadIdx.x % pairs;
  const int64_t g = (int64_t)blockIdx.x * per_block + threadIdx.x / pairs;
  if (g >= groups) return;
  const __nv_bfloat16* src = x + (g * S) * ld + 2 * p;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + (int64_t)s * ld));
    acc.x += v.x;
    acc.y += v.y;
  }
  *reinterpret_cast<float2*>(out + g * cols + 2 * p) = acc;
}

}  // namespace nlb

using namespace nlb;

static bool pow2_cols(int cols) { return cols >= 2 && cols <= 2 * kRedThreads && (cols & (cols - 1)) == 0; }

extern "C" int nlb_colsum_bf16(const void* x, int64_t M, int cols, int ld, float* out, void* stream) {
  if (!x || !out || M < 0 || ld < cols) { nlb_set_error("colsum_bf16: bad arguments"); return NLB_EINVAL; }
  if (!pow2_cols(cols) || (ld & 1)) { nlb_set_error("colsum_bf16: cols must be a power of two in [2,512] and ld even (cols=%d ld=%d)", cols, ld); return NLB_EUNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, sizeof(float) * cols, st) != cudaSuccess) return nlb_check_launch("colsum_bf16 memset");
  if (M == 0) return NLB_OK;
  int64_t want = (M + 255) / 256;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  k_colsum_bf16<<<blocks, kRedThreads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, cols, ld, out);
  return nlb_check_launch("colsum_bf16");
}

extern "C" int nlb_group_sum_bf16(const void* x, int64_t groups, int S, int cols, int ld, float* out, void* stream) {
  if (!x || !out || groups < 0 || S < 1 || ld < cols) { nlb_set_error("group_sum_bf16: bad arguments"); return NLB_EINVAL; }
  if (!pow2_cols(cols) || (ld & 1)) { nlb_set_error("group_sum_bf16: cols must be a power of two in [2,512] (cols=%d)", cols); return NLB_EUNSUPPORTED; }
  if (groups == 0) return NLB_OK;
  const int per_block = kRedThreads / (cols >> 1);
  const int64_t blocks = (groups + per_block - 1) / per_block;
  k_group_sum_bf16<<<(unsigned)blocks, kRedThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), groups, S, cols, ld, out);
  return nlb_check_launch("group_sum_bf16");
}
