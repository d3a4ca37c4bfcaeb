// Fused dense optimizer pass over a hash table (HBM-streaming, 16 B/lane vectors):
//   hash-decay gradient  Model.hash_decay_loss (Z/internal/models.py:203-223)
//   NaN scrub            train_utils.clip_gradients (Z/internal/train_utils.py:251-253)
//   Adam                 train_utils.create_optimizer (Z/internal/train_utils.py:256-275)
//   zero_grad            train.py:197
// The reference spends ~4.4 GB/step of HBM traffic in five separate passes
// (zero_grad, p^2 segment mean, its backward, nan_to_num, Adam); this is one pass
// of 16 B read + 16 B written per parameter.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

__device__ __forceinline__ float scrub(float g) {
  if (isnan(g)) return 0.f;
  if (isinf(g)) return g > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return g;
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float lr_c, float rsqrt_bc2,
                                            float beta1, float beta2, float eps) {
  m = m + (1.0f - beta1) * (g - m);            // exp_avg.lerp_(grad, 1-beta1)
  v = beta2 * v + (1.0f - beta2) * g * g;      // mul_(beta2).addcmul_(g, g, 1-beta2)
  const float denom = sqrtf(v) * rsqrt_bc2 + eps;
  p = p - lr_c * (m / denom);
}

constexpr int kMaxLevels = 32;
struct DecayTable {
  int64_t end[kMaxLevels];   // exclusive end (in floats) of every level
  float coef[kMaxLevels];    // 2*mult / (L*C*rows_l)
  int L;
};

template <bool kDecay>
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ param, float* __restrict__ grad,
                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, int64_t n,
                                              DecayTable dt, float lr_c, float rsqrt_bc2, float beta1, float beta2,
                                              float eps, float grad_scale, float* __restrict__ level_sumsq,
                                              const float* __restrict__ dyn, int64_t first) {
  __shared__ float s_sum[kMaxLevels];
  if (dyn) {  // CUDA-graph replay: this step's learning rate and bias corrections
    lr_c = __ldg(dyn + NLB_DYN_LR_C);
    rsqrt_bc2 = __ldg(dyn + NLB_DYN_RSQRT_BC2);
  }
  if (kDecay && level_sumsq) {
    if (threadIdx.x < kMaxLevels) s_sum[threadIdx.x] = 0.f;
    __syncthreads();
  }
  const int64_t n4 = n >> 2;
  // blocked partition: a block owns one contiguous range, so it sees one (rarely two) levels
  const int64_t per_block = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per_block;
  const int64_t hi = lo + per_block < n4 ? lo + per_block : n4;
  int cur_l = -1;
  float sq = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    // (evict-first __ldcs / __stcs accesses were measured: 4.8 TB/s against 5.2 TB/s with plain ones)
    float4 p = reinterpret_cast<float4*>(param)[i];
    float4 g = reinterpret_cast<float4*>(grad)[i];
    float4 m = reinterpret_cast<float4*>(exp_avg)[i];
    float4 v = reinterpret_cast<float4*>(exp_avg_sq)[i];
    float coef = 0.f;
    if (kDecay) {
      const int64_t e = first + (i << 2);  // level sizes are multiples of 8 rows, so a float4 never straddles levels
      int l = cur_l < 0 ? 0 : cur_l;
      while (l < dt.L - 1 && e >= dt.end[l]) ++l;
      coef = dt.coef[l];
      if (level_sumsq && l != cur_l) {
        if (cur_l >= 0) atomicAdd(&s_sum[cur_l], sq);
        sq = 0.f;
        cur_l = l;
      }
    }
    float gx = scrub(fmaf(p.x, coef, g.x * grad_scale));
    float gy = scrub(fmaf(p.y, coef, g.y * grad_scale));
    float gz = scrub(fmaf(p.z, coef, g.z * grad_scale));
    float gw = scrub(fmaf(p.w, coef, g.w * grad_scale));
    adam_update(p.x, gx, m.x, v.x, lr_c, rsqrt_bc2, beta1, beta2, eps);
    adam_update(p.y, gy, m.y, v.y, lr_c, rsqrt_bc2, beta1, beta2, eps);
    adam_update(p.z, gz, m.z, v.z, lr_c, rsqrt_bc2, beta1, beta2, eps);
    adam_update(p.w, gw, m.w, v.w, lr_c, rsqrt_bc2, beta1, beta2, eps);
    reinterpret_cast<float4*>(param)[i] = p;
    reinterpret_cast<float4*>(exp_avg)[i] = m;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
    reinterpret_cast<float4*>(grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kDecay && level_sumsq) sq += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;  // UPDATED parameters
  }
  if (kDecay && level_sumsq) {
    if (cur_l >= 0) atomicAdd(&s_sum[cur_l], sq);
    __syncthreads();
    if (threadIdx.x < dt.L && s_sum[threadIdx.x] != 0.f) atomicAdd(level_sumsq + threadIdx.x, s_sum[threadIdx.x]);
  }
  // tail (n not a multiple of 4): dense-layer tensors only
  if (!kDecay && blockIdx.x == 0) {
    for (int64_t e = (n4 << 2) + threadIdx.x; e < n; e += blockDim.x) {
      float p = param[e], m = exp_avg[e], v = exp_avg_sq[e];
      float g = scrub(grad[e] * grad_scale);
      adam_update(p, g, m, v, lr_c, rsqrt_bc2, beta1, beta2, eps);
      param[e] = p; exp_avg[e] = m; exp_avg_sq[e] = v; grad[e] = 0.f;
    }
  }
}

}  // namespace nlb

using namespace nlb;

// Grid of the streaming pass: every block owns one contiguous range.  Measured on the B200 (77.66 M parameters,
// three tables): 148 x 8 blocks 0.156 ms per table on average, 148 x 5 (= the resident blocks, one exact wave) 0.164,
// 148 x 4 0.160 -- the second, partial wave costs less than the longer per-block ranges do.
// NLB_ADAM_BLOCKS_PER_SM overrides (A/B timing).
static int adam_blocks(int64_t n, bool decay) {
  (void)decay;
  static const int per_sm = []() {
    const char* e = getenv("NLB_ADAM_BLOCKS_PER_SM");
    const int v = (e && *e) ? atoi(e) : 0;
    return v > 0 ? v : 8;
  }();
  const int64_t want = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)nlb_sm_count() * per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

static void bias_terms(float lr, float beta1, float beta2, int step, float& lr_c, float& rsqrt_bc2) {
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  lr_c = (float)((double)lr / bc1);
  rsqrt_bc2 = (float)(1.0 / sqrt(bc2));
}

extern "C" int nlb_adam_table_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                   const int32_t* offsets_host, int L, int C, float decay_mult, float lr, float beta1,
                                   float beta2, float eps, int step, float grad_scale, float* level_sumsq,
                                   void* stream) {
  if (!offsets_host || L < 1) { nlb_set_error("adam_table_step: null pointer"); return NLB_EINVAL; }
  return nlb_adam_table_step_range(param, grad, exp_avg, exp_avg_sq, offsets_host, L, C, decay_mult, lr, beta1, beta2, eps,
                                   step, grad_scale, level_sumsq, 0, (int64_t)offsets_host[L] * C, stream);
}

extern "C" int nlb_adam_table_step_range(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                         const int32_t* offsets_host, int L, int C, float decay_mult, float lr,
                                         float beta1, float beta2, float eps, int step, float grad_scale,
                                         float* level_sumsq, int64_t first, int64_t count, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !offsets_host) { nlb_set_error("adam_table_step: null pointer"); return NLB_EINVAL; }
  if (L < 1 || L > kMaxLevels) { nlb_set_error("adam_table_step: L=%d outside [1,%d]", L, kMaxLevels); return NLB_EINVAL; }
  if (step < 1) { nlb_set_error("adam_table_step: step counts from 1"); return NLB_EINVAL; }
  DecayTable dt;
  dt.L = L;
  for (int l = 0; l < L; ++l) {
    const int64_t rows = (int64_t)offsets_host[l + 1] - offsets_host[l];
    dt.end[l] = (int64_t)offsets_host[l + 1] * C;
    dt.coef[l] = (float)(2.0 * (double)decay_mult / ((double)L * C * (double)rows));
  }
  const int64_t total = (int64_t)offsets_host[L] * C;
  if (total % 4 != 0) { nlb_set_error("adam_table_step: table size must be a multiple of 4 floats"); return NLB_EINVAL; }
  if (first < 0 || count < 0 || first + count > total || (first | count) % 4 != 0) {
    nlb_set_error("adam_table_step: range [%lld, +%lld) must lie inside the table and be a multiple of 4 floats", (long long)first, (long long)count);
    return NLB_EINVAL;
  }
  if (count == 0) return NLB_OK;
  // this rank's slice of the table (data-parallel runs shard the optimizer pass; the moments are slice-local)
  float* p = param + first;
  float* g = grad + first;
  float lr_c, rs;
  bias_terms(lr, beta1, beta2, step, lr_c, rs);
  const int blocks = adam_blocks(count, decay_mult != 0.f);
  if (decay_mult != 0.f)
    k_adam<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, exp_avg, exp_avg_sq, count, dt, lr_c, rs, beta1, beta2, eps, grad_scale, level_sumsq, nlb_dynamic_scalars(), first);
  else
    k_adam<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, exp_avg, exp_avg_sq, count, dt, lr_c, rs, beta1, beta2, eps, grad_scale, nullptr, nlb_dynamic_scalars(), first);
  return nlb_check_launch("adam_table_step");
}

extern "C" int nlb_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) { nlb_set_error("adam_step: null pointer"); return NLB_EINVAL; }
  if (n == 0) return NLB_OK;
  if (step < 1) { nlb_set_error("adam_step: step counts from 1"); return NLB_EINVAL; }
  if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) {
    nlb_set_error("adam_step: buffers must be 16-byte aligned");
    return NLB_EINVAL;
  }
  DecayTable dt;
  dt.L = 0;
  float lr_c, rs;
  bias_terms(lr, beta1, beta2, step, lr_c, rs);
  const int blocks = adam_blocks(n, false);
  k_adam<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, dt, lr_c, rs, beta1, beta2, eps, grad_scale, nullptr, nlb_dynamic_scalars(), 0);
  return nlb_check_launch("adam_step");
}
