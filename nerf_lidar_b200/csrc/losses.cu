// Per-ray regularisers of the training step (SURVEY.md 8f "next #2"), one warp per ray,
// value AND gradient in one pass (both losses are differentiated only w.r.t. the
// per-sample weights; all distances are detached in the reference):
//   distortion      stepfun.lossfun_distortion   Z/internal/stepfun.py:297-307
//                   (the reference builds [N,S,S] tensors)
//   anti-interlevel train_utils.anti_interlevel_loss Z/internal/train_utils.py:134-172
//                   = stepfun.blur_stepfun (stepfun.py:425-433) + piecewise-quadratic CDF
//                   + math.sorted_interp_quad (math.py:111-131, [N,66,65] masks there)
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

constexpr int kLossWarps = 4;

// loss_ray = sum_ij w_i w_j |u_i - u_j| + (1/3) sum_i w_i^2 (t_{i+1} - t_i)
__global__ void __launch_bounds__(kLossWarps * 32) k_distortion(const float* __restrict__ sdist,
                                                                const float* __restrict__ weights, int N, int S,
                                                                float* __restrict__ loss_ray,
                                                                float* __restrict__ grad_w) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kLossWarps + warp;
  if (ray >= N) return;
  float* u = smem + (size_t)warp * 2 * S;
  float* w = u + S;
  const float* t = sdist + (size_t)ray * (S + 1);
  for (int i = lane; i < S; i += 32) {
    u[i] = (__ldg(t + i + 1) + __ldg(t + i)) / 2.0f;
    w[i] = __ldg(weights + (size_t)ray * S + i);
  }
  __syncwarp();
  float part = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float ui = u[i], wi = w[i];
    float inner = 0.f;
    for (int j = 0; j < S; ++j) inner = fmaf(w[j], fabsf(ui - u[j]), inner);
    const float dt = __ldg(t + i + 1) - __ldg(t + i);
    part += wi * inner + wi * wi * dt / 3.0f;
    if (grad_w) grad_w[(size_t)ray * S + i] = 2.0f * inner + 2.0f * wi * dt / 3.0f;
  }
  part = warp_sum(part);
  if (lane == 0) loss_ray[ray] = part;
}

__device__ __forceinline__ int lower_bound_l(const float* a, int n, float x) {  // #{a_j < x}
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int upper_bound_l(const float* a, int n, float x) {  // #{a_j <= x}
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}

__host__ __device__ inline int interlevel_smem_floats(int Sc, int Sp) {
  const int K = 2 * (Sc + 1);
  return (Sc + 1) + (Sc + 1) + 5 * K + (Sp + 1) + 8;
}

// loss_ray = sum_k max(w_s[k] - wp[k], 0)^2 / (wp[k] + 1e-5)
__global__ void __launch_bounds__(kLossWarps * 32) k_interlevel(const float* __restrict__ c_all,
                                                                const float* __restrict__ w_all, int Sc,
                                                                const float* __restrict__ cp_all,
                                                                const float* __restrict__ wp_all, int Sp, float r, int N,
                                                                float* __restrict__ loss_ray,
                                                                float* __restrict__ grad_wp) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kLossWarps + warp;
  if (ray >= N) return;
  const int n1 = Sc + 1, K = 2 * n1;
  float* base = smem + (size_t)warp * interlevel_smem_floats(Sc, Sp);
  float* c = base;            // n1 fenceposts of the NeRF level
  float* y1 = c + n1;         // n1 slope steps (w_norm differences / 2r)
  float* xr = y1 + n1;        // K sorted blurred knots
  float* sl = xr + K;         // K slope change carried by each sorted knot
  float* yr = sl + K;         // K blurred pdf at the knots
  float* cdf = yr + K;        // K piecewise-quadratic cdf
  float* pmx = cdf + K;       // K: prefix max of yr (sorted_interp_quad's "fpdf0")  | later reused
  float* ci = pmx + K;        // Sp+1 interpolated cdf at the proposal fenceposts

  for (int i = lane; i < n1; i += 32) c[i] = __ldg(c_all + (size_t)ray * n1 + i);
  __syncwarp();
  for (int k = lane; k < n1; k += 32) {
    // w_norm = min(w / (c[k+1]-c[k]), 10);  y1_k = (y_k - y_{k-1}) / (2r), y_{-1} = y_{Sc} = 0
    float yk = 0.f, ykm = 0.f;
    if (k < Sc) yk = fminf(__ldg(w_all + (size_t)ray * Sc + k) / (c[k + 1] - c[k]), 10.0f);
    if (k > 0) ykm = fminf(__ldg(w_all + (size_t)ray * Sc + k - 1) / (c[k] - c[k - 1]), 10.0f);
    y1[k] = (yk - ykm) / (2.0f * r);
  }
  __syncwarp();
  // merge {c - r} and {c + r} by rank; ties: minus-list first
  for (int k = lane; k < n1; k += 32) {
    const float a = c[k] - r, b = c[k] + r;
    // #{c_j + r < a} and #{c_j - r <= b}, searched on c with the shifted key
    int ra = k, rb = k;
    {
      int lo = 0, hi = n1;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] + r < a) lo = mid + 1; else hi = mid; }
      ra += lo;
      lo = 0; hi = n1;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] - r <= b) lo = mid + 1; else hi = mid; }
      rb += lo;
    }
    xr[ra] = a; sl[ra] = y1[k];
    xr[rb] = b; sl[rb] = -y1[k];
  }
  __syncwarp();
  if (lane == 0) {
    // yr = [0, clamp_min(cumsum(dx * cumsum(slope)), 0)], cdf = [0, cumsum(trapezoids)], running max
    float slope = 0.f, acc = 0.f, area = 0.f, prev = 0.f, mx;
    yr[0] = 0.f; cdf[0] = 0.f; pmx[0] = 0.f; mx = 0.f;
    for (int m = 0; m < K - 1; ++m) {
      slope += sl[m];
      const float dx = xr[m + 1] - xr[m];
      acc += dx * slope;
      const float cur = fmaxf(acc, 0.f);
      area += 0.5f * (cur + prev) * dx;
      yr[m + 1] = cur;
      cdf[m + 1] = area;
      mx = fmaxf(mx, cur);
      pmx[m + 1] = mx;
      prev = cur;
    }
    // suffix min of yr -> sl (slopes are dead)
    float mn = yr[K - 1];
    for (int m = K - 1; m >= 0; --m) { mn = fminf(mn, yr[m]); sl[m] = mn; }
  }
  __syncwarp();
  const float* cp = cp_all + (size_t)ray * (Sp + 1);
  for (int q = lane; q <= Sp; q += 32) {
    const float x = __ldg(cp + q);
    const int cnt = upper_bound_l(xr, K, x);
    const int i0 = min(max(cnt - 1, 0), K - 1), i1 = min(cnt, K - 1);
    // max of the values under the mask / min of the values outside it (defaults: first / last)
    const float f0 = cnt == 0 ? yr[0] : pmx[i0];
    const float f1 = cnt == K ? yr[K - 1] : sl[i1];
    const float F0 = cnt == 0 ? cdf[0] : cdf[i0];
    const float x0 = cnt == 0 ? xr[0] : xr[i0];
    const float x1 = cnt == K ? xr[K - 1] : xr[i1];
    float off = (x - x0) / (x1 - x0);
    if (isnan(off)) off = 0.f;
    off = fminf(fmaxf(off, 0.f), 1.f);
    ci[q] = F0 + (x - x0) * (f0 + f1 * off + f0 * (1.0f - off)) / 2.0f;
  }
  __syncwarp();
  float part = 0.f;
  for (int k = lane; k < Sp; k += 32) {
    const float ws = ci[k + 1] - ci[k];
    const float wpk = __ldg(wp_all + (size_t)ray * Sp + k);
    const float d = fmaxf(ws - wpk, 0.f);
    const float den = wpk + 1e-5f;
    part += d * d / den;
    if (grad_wp) grad_wp[(size_t)ray * Sp + k] = -2.0f * d / den - d * d / (den * den);
  }
  part = warp_sum(part);
  if (lane == 0) loss_ray[ray] = part;
}


// ----------------------------------------------------------------------------- loss assembly
// Z/train.py:283-462 sums its loss dictionary and back-propagates the sum; issued as torch scalar arithmetic that
// is ~45 tiny launches per step (means, multiplier products, the python sum, their backward nodes).  Two generic
// launches instead: k_weighted_sums forms every scalar of the dictionary in one block, k_scale_tensors seeds all
// gradients of a backward pass.
constexpr int kMaxSumTerms = 24, kMaxSumOut = 16, kMaxScaleJobs = 8;
struct SumTerms {
  nlb_sum_term_t t[kMaxSumTerms];
  int n;
};

__global__ void __launch_bounds__(1024) k_weighted_sums(const __grid_constant__ SumTerms T, float* __restrict__ out,
                                                        int nout) {
  __shared__ float s_warp[32];
  __shared__ float s_val[kMaxSumTerms];   // sum_i x[i] w[i] of every term that reads memory
  __shared__ float s_out[kMaxSumOut];
  __shared__ int s_named[kMaxSumOut];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kMaxSumOut) { s_out[tid] = 0.f; s_named[tid] = 0; }
  // scalar terms (most of a loss dictionary): one thread each, all loads in flight together -- walking them one
  // after the other paid a global round trip per term.  A zero coefficient reads nothing (0 * NaN must stay 0).
  if (tid < T.n) {
    const nlb_sum_term_t& q = T.t[tid];
    if (q.x != nullptr && q.n == 1 && q.coef != 0.f) s_val[tid] = __ldg(q.x) * (q.w ? __ldg(q.w) : 1.0f);
  }
  // array terms: the whole block, one term after the other
  for (int k = 0; k < T.n; ++k) {
    const nlb_sum_term_t& q = T.t[k];
    if (q.x == nullptr || q.n <= 1 || q.coef == 0.f) continue;   // (uniform over the block)
    float a = 0.f;
    if (q.w) {
      for (int64_t i = tid; i < q.n; i += 1024) a = fmaf(__ldg(q.x + i), __ldg(q.w + i), a);
    } else {
      for (int64_t i = tid; i < q.n; i += 1024) a += __ldg(q.x + i);
    }
    a = warp_sum(a);
    __syncthreads();               // s_warp of the previous term has been read
    if (lane == 0) s_warp[warp] = a;
    __syncthreads();
    if (warp == 0) {
      a = warp_sum(s_warp[lane]);
      if (lane == 0) s_val[k] = a;
    }
  }
  __syncthreads();
  if (tid == 0) {                  // the terms in order: a term without x adds an output formed so far
    for (int k = 0; k < T.n; ++k) {
      const nlb_sum_term_t& q = T.t[k];
      s_named[q.out_index] = 1;
      if (q.coef == 0.f) continue;
      const float v = q.x == nullptr ? s_out[(int)q.n] : (q.n == 0 ? 0.f : s_val[k]);
      s_out[q.out_index] += q.coef * v;
    }
  }
  __syncthreads();
  if (tid < nout && s_named[tid]) out[tid] = s_out[tid];
}

struct ScaleJobs {
  nlb_scale_job_t j[kMaxScaleJobs];
};

__global__ void __launch_bounds__(256) k_scale_tensors(const __grid_constant__ ScaleJobs J) {
  const nlb_scale_job_t& q = J.j[blockIdx.y];
  const float g = q.g ? __ldg(q.g) : 1.0f;
  const float a = q.coef * g * (q.s ? __ldg(q.s) : 1.0f);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (q.src2) {
    const float b = q.coef2 * g * (q.s2 ? __ldg(q.s2) : 1.0f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += stride)
      q.dst[i] = fmaf(__ldg(q.src2 + i), b, __ldg(q.src + i) * a);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += stride) q.dst[i] = __ldg(q.src + i) * a;
  }
}

}  // namespace nlb

using namespace nlb;

extern "C" int nlb_distortion_loss(const float* sdist, const float* weights, int N, int S, float* loss_ray,
                                   float* grad_w, void* stream) {
  if (N == 0) return NLB_OK;
  if (!sdist || !weights || !loss_ray || S < 1) { nlb_set_error("distortion_loss: bad arguments"); return NLB_EINVAL; }
  const size_t smem = (size_t)kLossWarps * 2 * S * sizeof(float);
  if (smem > 48 * 1024) { nlb_set_error("distortion_loss: S=%d too large", S); return NLB_EUNSUPPORTED; }
  k_distortion<<<div_up(N, kLossWarps), kLossWarps * 32, smem, (cudaStream_t)stream>>>(sdist, weights, N, S, loss_ray, grad_w);
  return nlb_check_launch("distortion_loss");
}

extern "C" int nlb_interlevel_loss(const float* c, const float* w, int Sc, const float* cp, const float* wp, int Sp,
                                   float pulse_width, int N, float* loss_ray, float* grad_wp, void* stream) {
  if (N == 0) return NLB_OK;
  if (!c || !w || !cp || !wp || !loss_ray || Sc < 1 || Sp < 1 || !(pulse_width > 0.f)) {
    nlb_set_error("interlevel_loss: bad arguments");
    return NLB_EINVAL;
  }
  const size_t smem = (size_t)kLossWarps * interlevel_smem_floats(Sc, Sp) * sizeof(float);
  if (smem > 48 * 1024) { nlb_set_error("interlevel_loss: Sc=%d / Sp=%d too large", Sc, Sp); return NLB_EUNSUPPORTED; }
  k_interlevel<<<div_up(N, kLossWarps), kLossWarps * 32, smem, (cudaStream_t)stream>>>(c, w, Sc, cp, wp, Sp, pulse_width, N,
                                                                                     loss_ray, grad_wp);
  return nlb_check_launch("interlevel_loss");
}

extern "C" int nlb_weighted_sums(const nlb_sum_term_t* terms, int nterms, float* out, int nout, void* stream) {
  if (!terms || !out || nterms < 1 || nterms > kMaxSumTerms || nout < 1 || nout > kMaxSumOut) {
    nlb_set_error("weighted_sums: 1..%d terms, 1..%d outputs", kMaxSumTerms, kMaxSumOut);
    return NLB_EINVAL;
  }
  SumTerms T;
  T.n = nterms;
  for (int k = 0; k < nterms; ++k) {
    const nlb_sum_term_t& q = terms[k];
    if (q.out_index < 0 || q.out_index >= nout || q.n < 0 || (!q.x && q.n >= nout)) {
      nlb_set_error("weighted_sums: term %d names an output outside [0,%d)", k, nout);
      return NLB_EINVAL;
    }
    T.t[k] = q;
  }
  k_weighted_sums<<<1, 1024, 0, (cudaStream_t)stream>>>(T, out, nout);
  return nlb_check_launch("weighted_sums");
}

extern "C" int nlb_scale_tensors(const nlb_scale_job_t* jobs, int njobs, void* stream) {
  if (!jobs || njobs < 1 || njobs > kMaxScaleJobs) { nlb_set_error("scale_tensors: 1..%d jobs", kMaxScaleJobs); return NLB_EINVAL; }
  ScaleJobs J;
  int64_t longest = 0;
  for (int k = 0; k < njobs; ++k) {
    const nlb_scale_job_t& q = jobs[k];
    if (q.n < 0 || (q.n > 0 && (!q.src || !q.dst))) { nlb_set_error("scale_tensors: bad arguments in job %d", k); return NLB_EINVAL; }
    J.j[k] = q;
    if (q.n > longest) longest = q.n;
  }
  if (longest == 0) return NLB_OK;
  int64_t bx = (longest + 1023) / 1024;   // four elements per thread
  if (bx > 148 * 8) bx = 148 * 8;
  k_scale_tensors<<<dim3((unsigned)bx, njobs), 256, 0, (cudaStream_t)stream>>>(J);
  return nlb_check_launch("scale_tensors");
}
