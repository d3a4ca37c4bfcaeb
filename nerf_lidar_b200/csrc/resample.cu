// Step-function resampling of one sampling level, one warp per ray:
//   stepfun.max_dilate_weights  (Z/internal/stepfun.py:64-105)  [levels > 0]
//   [1:-1] trim, anneal * log(w) logits (Z/internal/models.py:339-355)
//   softmax + integrate_weights + sorted_interp (stepfun.py:108-161, math.py:89-108)
//   sample_intervals midpoints / reflected ends (stepfun.py:251-294)
//   s_to_t power transformation (coord.py:103-162)
// The reference materialises [N,193,64] and [N,191,64] masks for the dilation and
// the interpolation; here a ray's 64-256 intervals live in shared memory, the three
// sorted fencepost lists are merged by rank (binary searches), the max-pool is a
// range scan, the CDF is a warp scan and the inversion a binary search.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

// #{j < n : a[j] < x}
__device__ __forceinline__ int lower_bound(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// #{j < n : a[j] <= x}
__device__ __forceinline__ int upper_bound(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// math.sorted_interp for one query: lower knot = last j with xp[j] <= x (default 0),
// upper knot = first j with xp[j] > x (default n-1); offset = clip(nan_to_num(.,0),0,1).
__device__ __forceinline__ float interp_sorted(float x, const float* xp, const float* fp, int n, int& i0_out) {
  int cnt = upper_bound(xp, n, x);
  int i0 = min(max(cnt - 1, 0), n - 1);
  int i1 = min(cnt, n - 1);
  float x0 = xp[i0], x1 = xp[i1], f0 = fp[i0], f1 = fp[i1];
  float off = __fdiv_rn(__fsub_rn(x, x0), __fsub_rn(x1, x0));
  if (isnan(off)) off = 0.f;           // nan_to_num(nan -> 0); +-inf fall to the clip
  off = fminf(fmaxf(off, 0.f), 1.f);
  i0_out = i0;
  return __fadd_rn(f0, __fmul_rn(off, __fsub_rn(f1, f0)));
}

constexpr int kWarpsPerBlock = 4;

__host__ __device__ inline int resample_smem_floats(int n_in, int S) {
  // t(n+1) p(n) a(n) b(n) td(3n+1) wd(3n) cw(3n+1) centers(S)
  return (n_in + 1) + 3 * n_in + (3 * n_in + 1) + 3 * n_in + (3 * n_in + 1) + S + 8;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_resample(
    const float* __restrict__ sdist_in, const float* __restrict__ weights_in, int n, int dilate, float dilation,
    float anneal, float pad, const float* __restrict__ u_base, const float* __restrict__ jitter, float max_jitter,
    const float* __restrict__ near, const float* __restrict__ far, float lam, int S, int N,
    float* __restrict__ sdist_out, float* __restrict__ tdist_out, int32_t* __restrict__ sample_idx,
    const float* __restrict__ dyn) {
  extern __shared__ float smem[];
  if (dyn) anneal = __ldg(dyn + NLB_DYN_ANNEAL);  // CUDA-graph replay: this step's value
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= N) return;
  float* base = smem + (size_t)warp * resample_smem_floats(n, S);
  float* t = base;                  // n+1
  float* p = t + (n + 1);           // n   (weights, then pdf)
  float* a = p + n;                 // n
  float* b = a + n;                 // n
  float* td = b + n;                // 3n+1
  float* wd = td + (3 * n + 1);     // 3n
  float* cw = wd + 3 * n;           // 3n+1
  float* centers = cw + (3 * n + 1);// S

  for (int i = lane; i <= n; i += 32) t[i] = sdist_in ? __ldg(sdist_in + (size_t)ray * (n + 1) + i) : (float)i;
  for (int i = lane; i < n; i += 32) p[i] = weights_in ? __ldg(weights_in + (size_t)ray * n + i) : 1.0f;
  __syncwarp();

  const float* knots;
  float* wts;
  int m;
  if (dilate) {
    for (int j = lane; j < n; j += 32) {
      float t0 = t[j], t1 = t[j + 1];
      a[j] = __fsub_rn(t0, dilation);
      b[j] = __fadd_rn(t1, dilation);
    }
    __syncwarp();
    for (int j = lane; j < n; j += 32) p[j] = __fdiv_rn(p[j], fmaxf(__fsub_rn(t[j + 1], t[j]), kEps));
    // rank-merge of the three sorted lists; ties ordered t < a < b
    for (int i = lane; i <= n; i += 32) {
      float x = t[i];
      int r = i + lower_bound(a, n, x) + lower_bound(b, n, x);
      td[r] = fminf(fmaxf(x, 0.f), 1.f);
    }
    for (int j = lane; j < n; j += 32) {
      float x = a[j];
      int r = j + upper_bound(t, n + 1, x) + lower_bound(b, n, x);
      td[r] = fminf(fmaxf(x, 0.f), 1.f);
      x = b[j];
      r = j + upper_bound(t, n + 1, x) + upper_bound(a, n, x);
      td[r] = fminf(fmaxf(x, 0.f), 1.f);
    }
    __syncwarp();
    // max-pool the pdf over the dilated support, back to weights
    float part = 0.f;
    for (int k = lane; k < 3 * n; k += 32) {
      float x = td[k];
      int jlo = upper_bound(b, n, x);      // first j with b[j] > x
      int jhi = upper_bound(a, n, x) - 1;  // last j with a[j] <= x
      float mx = 0.f;
      for (int j = jlo; j <= jhi; ++j) mx = fmaxf(mx, p[j]);
      float w = __fmul_rn(mx, __fsub_rn(td[k + 1], x));
      wd[k] = w;
      part += w;
    }
    float tot = fmaxf(warp_sum(part), kEps);
    __syncwarp();
    for (int k = lane; k < 3 * n; k += 32) wd[k] = __fdiv_rn(wd[k], tot);
    __syncwarp();
    knots = td + 1;
    wts = wd + 1;
    m = 3 * n - 2;
  } else {
    knots = t;
    wts = p;
    m = n;
  }

  // logits -> softmax (in place in wts)
  float mx = -INFINITY;
  for (int k = lane; k < m; k += 32) {
    float lg = (knots[k + 1] > knots[k]) ? __fmul_rn(anneal, logf(__fadd_rn(wts[k], pad))) : -INFINITY;
    wts[k] = lg;
    mx = fmaxf(mx, lg);
  }
  mx = warp_max(mx);
  float part = 0.f;
  for (int k = lane; k < m; k += 32) {
    float e = expf(__fsub_rn(wts[k], mx));
    wts[k] = e;
    part += e;
  }
  const float denom = warp_sum(part);
  __syncwarp();
  // CDF: cw[0]=0, cw[k]=min(1, sum_{i<k} w_i), cw[m]=1  (blocked warp scan)
  const int chunk = (m + 31) / 32;
  const int k0 = lane * chunk, k1 = min(m, k0 + chunk);
  float run = 0.f;
  for (int k = k0; k < k1; ++k) {
    run += __fdiv_rn(wts[k], denom);
    cw[k + 1] = run;  // local inclusive sums
  }
  float incl = warp_scan_incl(run, lane);
  float excl = incl - run;
  for (int k = k0; k < k1; ++k) cw[k + 1] = fminf(__fadd_rn(cw[k + 1], excl), 1.0f);
  if (lane == 0) cw[0] = 0.f;
  __syncwarp();
  if (lane == 0) cw[m] = 1.0f;
  __syncwarp();

  const float jit = jitter ? __fmul_rn(__ldg(jitter + ray), max_jitter) : 0.f;
  for (int s = lane; s < S; s += 32) {
    float u = __ldg(u_base + s);
    if (jitter) u = __fadd_rn(u, jit);
    int i0;
    centers[s] = interp_sorted(u, cw, knots, m + 1, i0);
    if (sample_idx) sample_idx[(size_t)ray * S + s] = i0;
  }
  __syncwarp();
  const RayWarp rw = make_warp(__ldg(near + ray), __ldg(far + ray), lam);
  for (int s = lane; s <= S; s += 32) {
    float v;
    if (s == 0) {
      float mid = __fdiv_rn(__fadd_rn(centers[1], centers[0]), 2.0f);
      v = fmaxf(__fsub_rn(__fmul_rn(2.0f, centers[0]), mid), 0.0f);
    } else if (s == S) {
      float mid = __fdiv_rn(__fadd_rn(centers[S - 1], centers[S - 2]), 2.0f);
      v = fminf(__fsub_rn(__fmul_rn(2.0f, centers[S - 1]), mid), 1.0f);
    } else {
      v = __fdiv_rn(__fadd_rn(centers[s], centers[s - 1]), 2.0f);
    }
    sdist_out[(size_t)ray * (S + 1) + s] = v;
    if (tdist_out) tdist_out[(size_t)ray * (S + 1) + s] = s_to_t(rw, v);
  }
}

__global__ void __launch_bounds__(128) k_sorted_interp(const float* __restrict__ x, const float* __restrict__ xp,
                                                       const float* __restrict__ fp, int N, int nx, int np,
                                                       float* __restrict__ out, int32_t* __restrict__ idx) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 4 + warp;
  if (ray >= N) return;
  float* sx = smem + (size_t)warp * 2 * np;
  float* sf = sx + np;
  for (int i = lane; i < np; i += 32) {
    sx[i] = __ldg(xp + (size_t)ray * np + i);
    sf[i] = __ldg(fp + (size_t)ray * np + i);
  }
  __syncwarp();
  for (int s = lane; s < nx; s += 32) {
    int i0;
    out[(size_t)ray * nx + s] = interp_sorted(__ldg(x + (size_t)ray * nx + s), sx, sf, np, i0);
    if (idx) idx[(size_t)ray * nx + s] = i0;
  }
}

}  // namespace nlb

using namespace nlb;

extern "C" int nlb_resample(const float* sdist_in, const float* weights_in, int n_in, int dilate, float dilation,
                            float anneal, float resample_padding, const float* u_base, const float* jitter,
                            float max_jitter, const float* near, const float* far, float lam, int S, int N,
                            float* sdist_out, float* tdist_out, int32_t* sample_idx, void* stream) {
  if (N == 0) return NLB_OK;
  if (S <= 1) { nlb_set_error("num_samples must be > 1, is %d.", S); return NLB_EINVAL; }
  if (n_in < 1 || !u_base || !near || !far || !sdist_out) { nlb_set_error("resample: bad arguments"); return NLB_EINVAL; }
  if (dilate && (!sdist_in || !weights_in)) { nlb_set_error("resample: dilation needs an input step function"); return NLB_EINVAL; }
  if (!sdist_in && n_in != 1) { nlb_set_error("resample: sdist_in may be NULL only for the initial [0,1] interval"); return NLB_EINVAL; }
  size_t smem = (size_t)kWarpsPerBlock * resample_smem_floats(n_in, S) * sizeof(float);
  if (smem > 200 * 1024) { nlb_set_error("resample: n_in=%d too large for shared memory", n_in); return NLB_EUNSUPPORTED; }
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_resample<<<div_up(N, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      sdist_in, weights_in, n_in, dilate, dilation, anneal, resample_padding, u_base, jitter, max_jitter, near, far,
      lam, S, N, sdist_out, tdist_out, sample_idx, nlb_dynamic_scalars());
  return nlb_check_launch("resample");
}

extern "C" int nlb_sorted_interp(const float* x, const float* xp, const float* fp, int N, int nx, int np,
                                 float* out, int32_t* idx, void* stream) {
  if (N == 0 || nx == 0) return NLB_OK;
  if (np < 1) { nlb_set_error("sorted_interp: np must be >= 1"); return NLB_EINVAL; }
  size_t smem = (size_t)4 * 2 * np * sizeof(float);
  if (smem > 200 * 1024) { nlb_set_error("sorted_interp: np too large"); return NLB_EUNSUPPORTED; }
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_sorted_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_sorted_interp<<<div_up(N, 4), 128, smem, (cudaStream_t)stream>>>(x, xp, fp, N, nx, np, out, idx);
  return nlb_check_launch("sorted_interp");
}
