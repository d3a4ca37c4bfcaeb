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


// ----------------------------------------------------------------------------- all sums of one backward pass, ONE launch
// The seven reductions of a NerfMLP backward (five bias gradients, two per-ray sums) as jobs of one grid: 16-byte
// loads (a thread owns 8 columns, unrolled four rows deep = 64 bytes in flight per thread), blocks of ~64-128 KB.
// Seven launches of 10-170 MB each had their tails and gaps exposed (0.22 ms for 661 MB = 3 TB/s).
constexpr int kMaxSumJobs = 8;
struct SumJobs {
  const __nv_bfloat16* x[kMaxSumJobs];
  float* out[kMaxSumJobs];
  int64_t rows[kMaxSumJobs];
  int cols[kMaxSumJobs], ld[kMaxSumJobs], group[kMaxSumJobs];
  int block0[kMaxSumJobs + 1];
  int n;
};

__device__ __forceinline__ void add8(float (&a)[8], const uint4 v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    a[2 * k] += f.x;
    a[2 * k + 1] += f.y;
  }
}

__global__ void __launch_bounds__(kRedThreads) k_bf16_sums(const __grid_constant__ SumJobs J) {
  __shared__ float s_part[kRedThreads][9];
  int j = 0;
  while (j + 1 < J.n && (int)blockIdx.x >= J.block0[j + 1]) ++j;
  const int b = blockIdx.x - J.block0[j], nb = J.block0[j + 1] - J.block0[j];
  const __nv_bfloat16* __restrict__ x = J.x[j];
  const int cols = J.cols[j], ld = J.ld[j], S = J.group[j];
  const int tpr = cols >> 3;                 // threads per row
  const int lanes = kRedThreads / tpr;       // rows (or groups) in flight per block
  const int p = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (S == 0) {  // column sums over this block's row range -> atomics into out[cols]
    const int64_t rows = J.rows[j];
    const int64_t per_block = (rows + nb - 1) / nb;
    const int64_t r0 = (int64_t)b * per_block;
    const int64_t r1 = r0 + per_block < rows ? r0 + per_block : rows;
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += lanes)
      add8(acc, __ldg(reinterpret_cast<const uint4*>(x + r * ld + 8 * p)));
#pragma unroll
    for (int k = 0; k < 8; ++k) s_part[threadIdx.x][k] = acc[k];
    __syncthreads();
    if ((int)threadIdx.x < cols) {
      const int pp = threadIdx.x >> 3, k = threadIdx.x & 7;
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += s_part[l * tpr + pp][k];
      atomicAdd(J.out[j] + threadIdx.x, t);
    }
  } else {       // sums over the S consecutive rows of a group -> out[groups, cols]
    const int64_t groups = J.rows[j] / S;
    const int64_t g = (int64_t)b * lanes + rl;
    if (g >= groups) return;
    const __nv_bfloat16* src = x + (g * S) * ld + 8 * p;
#pragma unroll 4
    for (int s = 0; s < S; ++s) add8(acc, __ldg(reinterpret_cast<const uint4*>(src + (int64_t)s * ld)));
    float4* dst = reinterpret_cast<float4*>(J.out[j] + g * cols + 8 * p);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
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

extern "C" int nlb_bf16_sums(const nlb_bf16_sum_job_t* jobs, int njobs, void* stream) {
  if (!jobs || njobs < 1 || njobs > kMaxSumJobs) { nlb_set_error("bf16_sums: 1..%d jobs", kMaxSumJobs); return NLB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  SumJobs J;
  J.n = 0;
  J.block0[0] = 0;
  for (int i = 0; i < njobs; ++i) {
    const nlb_bf16_sum_job_t& q = jobs[i];
    if (!q.x || !q.out || q.rows < 0 || q.group < 0 || q.ld < q.cols) { nlb_set_error("bf16_sums: bad arguments in job %d", i); return NLB_EINVAL; }
    if (q.cols < 8 || q.cols > kRedThreads || (q.cols & (q.cols - 1)) != 0 || (q.ld & 7) != 0 ||
        ((uintptr_t)q.x & 15) != 0 || ((uintptr_t)q.out & 15) != 0) {
      nlb_set_error("bf16_sums: job %d needs cols a power of two in [8,256], ld a multiple of 8 and 16-byte aligned pointers (cols=%d ld=%d)", i, q.cols, q.ld);
      return NLB_EUNSUPPORTED;
    }
    if (q.group > 0 && q.rows % q.group != 0) { nlb_set_error("bf16_sums: job %d: rows %lld not a multiple of the group size %d", i, (long long)q.rows, q.group); return NLB_EINVAL; }
    if (q.group == 0) {  // outputs are accumulated with atomics: clear them (adjacent outputs as one fill)
      size_t bytes = sizeof(float) * q.cols;
      int k = i;
      while (k + 1 < njobs && jobs[k + 1].group == 0 && jobs[k + 1].out == jobs[k].out + jobs[k].cols) bytes += sizeof(float) * jobs[++k].cols;
      const bool covered = i > 0 && jobs[i - 1].group == 0 && jobs[i - 1].out + jobs[i - 1].cols == q.out;
      if (!covered && cudaMemsetAsync(q.out, 0, bytes, st) != cudaSuccess) return nlb_check_launch("bf16_sums memset");
    }
    if (q.rows == 0) continue;
    const int lanes = kRedThreads / (q.cols >> 3);
    int64_t nb;
    if (q.group == 0) {
      nb = (q.rows * q.cols * 2 + 65535) / 65536;           // ~64 KB per block
      const int64_t cap = (q.rows + lanes - 1) / lanes;     // at least one row per row lane
      if (nb > cap) nb = cap;
      if (nb < 1) nb = 1;
    } else {
      nb = (q.rows / q.group + lanes - 1) / lanes;
    }
    if (nb > 0x3fffffff - J.block0[J.n]) { nlb_set_error("bf16_sums: grid too large"); return NLB_EINVAL; }
    const int n = J.n++;
    J.x[n] = reinterpret_cast<const __nv_bfloat16*>(q.x);
    J.out[n] = q.out;
    J.rows[n] = q.rows;
    J.cols[n] = q.cols;
    J.ld[n] = q.ld;
    J.group[n] = q.group;
    J.block0[n + 1] = J.block0[n] + (int)nb;
  }
  if (J.n == 0) return NLB_OK;
  k_bf16_sums<<<J.block0[J.n], kRedThreads, 0, st>>>(J);
  return nlb_check_launch("bf16_sums");
}
