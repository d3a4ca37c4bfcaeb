// Alpha compositing, one warp per ray:
//   render.compute_alpha_weights  (Z/internal/render.py:170-189)
//   render.volumetric_rendering   (Z/internal/render.py:192-284) incl.
//   stepfun.weighted_percentile   (Z/internal/stepfun.py:329-339)
// The reference runs ~60 eager kernels per level; here one kernel reads each
// per-sample tensor exactly once (warp scan for the transmittance, warp
// reductions for the K-channel weighted sums) and writes the per-ray outputs.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

constexpr int kCompWarps = 4;
constexpr int kMaxK = 32;  // semantic classes handled by one lane each in the column sums

__device__ __forceinline__ int upper_bound_f(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float interp_sorted_c(float x, const float* xp, const float* fp, int n) {
  int cnt = upper_bound_f(xp, n, x);
  int i0 = min(max(cnt - 1, 0), n - 1);
  int i1 = min(cnt, n - 1);
  float x0 = xp[i0], x1 = xp[i1], f0 = fp[i0], f1 = fp[i1];
  float off = __fdiv_rn(__fsub_rn(x, x0), __fsub_rn(x1, x0));
  if (isnan(off)) off = 0.f;
  off = fminf(fmaxf(off, 0.f), 1.f);
  return __fadd_rn(f0, __fmul_rn(off, __fsub_rn(f1, f0)));
}

// smem per warp: cw[S+2], taug[S+2], wchunk[32], prod[32*K]
__host__ __device__ inline int comp_smem_floats(int S, int K) { return 2 * (S + 2) + 32 + 32 * (K > 0 ? K : 1); }

__global__ void __launch_bounds__(kCompWarps * 32) k_composite_fwd(nlb_composite_in_t in, nlb_composite_out_t out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCompWarps + warp;
  if (ray >= in.N) return;
  const int S = in.S, K = in.K;
  float* cw = smem + (size_t)warp * comp_smem_floats(S, K);
  float* taug = cw + (S + 2);
  float* wchunk = taug + (S + 2);
  float* prod = wchunk + 32;

  const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
              dz = __ldg(in.directions + 3 * ray + 2);
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float* td = in.tdist + (size_t)ray * (S + 1);

  float carry = 0.f;   // sum of density*delta over previous chunks
  float wcarry = 0.f;  // sum of weights over previous chunks
  float a_acc = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_dep = 0.f, a_log = 0.f, a_int = 0.f;
  float a_sem = 0.f;  // lane k < K accumulates class k
  for (int c0 = 0; c0 < S; c0 += 32) {
    const int s = c0 + lane;
    const bool ok = s < S;
    float t0 = 0.f, t1 = 0.f, dd = 0.f;
    if (ok) {
      t0 = __ldg(td + s);
      t1 = __ldg(td + s + 1);
      dd = __fmul_rn(__ldg(in.density + (size_t)ray * S + s), __fmul_rn(__fsub_rn(t1, t0), dnorm));
    }
    const float dd_scan = (ok && !(in.opaque_background && s == S - 1)) ? dd : 0.f;
    if (ok && in.opaque_background && s == S - 1) dd = INFINITY;
    const float incl = warp_scan_incl(dd_scan, lane);
    float excl = __shfl_up_sync(NLB_FULL_MASK, incl, 1);
    if (lane == 0) excl = 0.f;
    const float trans = expf(-__fadd_rn(carry, excl));
    const float alpha = __fsub_rn(1.0f, expf(-dd));
    const float w = ok ? __fmul_rn(alpha, trans) : 0.f;
    carry = __fadd_rn(carry, __shfl_sync(NLB_FULL_MASK, incl, 31));
    if (ok) {
      if (out.weights) out.weights[(size_t)ray * S + s] = w;
      taug[s] = t0;
      if (s == S - 1) taug[S] = t1;
      const float tm = __fmul_rn(0.5f, __fadd_rn(t0, t1));
      a_acc += w;
      a_dep = fmaf(w, tm, a_dep);
      if (in.compute_extras) a_log = fmaf(w, logf(tm), a_log);
      if (in.rgb) {
        const float* c = in.rgb + ((size_t)ray * S + s) * 3;
        a_r = fmaf(w, __ldg(c), a_r);
        a_g = fmaf(w, __ldg(c + 1), a_g);
        a_b = fmaf(w, __ldg(c + 2), a_b);
      }
      if (in.intensity) a_int = fmaf(w, __ldg(in.intensity + (size_t)ray * S + s), a_int);
    }
    wchunk[lane] = w;
    // running (unclamped) sum of the weights for the percentile CDF
    {
      const float wi = warp_scan_incl(w, lane);
      if (ok) cw[s + 1] = __fadd_rn(wcarry, wi);
      wcarry = __fadd_rn(wcarry, __shfl_sync(NLB_FULL_MASK, wi, 31));
    }
    __syncwarp();
    if (in.semantic && out.semantic) {
      const int cnt = min(32, S - c0) * K;
      const float* sp = in.semantic + ((size_t)ray * S + c0) * K;
      for (int i = lane; i < cnt; i += 32) prod[i] = wchunk[i / K] * __ldg(sp + i);
      __syncwarp();
      if (lane < K) {
        const int rows = min(32, S - c0);
        for (int r = 0; r < rows; ++r) a_sem += prod[r * K + lane];
      }
      __syncwarp();
    }
  }
  __syncwarp();
  const float acc = warp_sum(a_acc);
  const float bg_w = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  const float den = fmaxf(acc, kEps);
  if (out.rgb) {
    float r = warp_sum(a_r), g = warp_sum(a_g), b = warp_sum(a_b);
    if (lane == 0) {
      out.rgb[3 * (size_t)ray] = fmaf(bg_w, in.bg, r);
      out.rgb[3 * (size_t)ray + 1] = fmaf(bg_w, in.bg, g);
      out.rgb[3 * (size_t)ray + 2] = fmaf(bg_w, in.bg, b);
    }
  }
  const float dep = warp_sum(a_dep);
  if (out.depth && lane == 0) out.depth[ray] = __fdiv_rn(dep, den);
  if (out.acc && lane == 0) out.acc[ray] = acc;
  if (in.intensity && out.intensity) {
    float v = warp_sum(a_int);
    if (lane == 0) out.intensity[ray] = v;
  }
  if (in.semantic && out.semantic && lane < K) out.semantic[(size_t)ray * K + lane] = a_sem;
  if (in.compute_extras) {
    const float lg = warp_sum(a_log);
    if (out.distance_mean && lane == 0) {
      float v = expf(__fdiv_rn(lg, den));
      if (isnan(v)) v = INFINITY;
      out.distance_mean[ray] = fminf(fmaxf(v, taug[0]), taug[S]);
    }
    if (out.distance_percentiles) {
      // cw = [0, min(1,cumsum(w)), 1] (S+2 knots), t_aug = [tdist, far]
      for (int s = lane; s < S; s += 32) cw[s + 1] = fminf(cw[s + 1], 1.0f);
      if (lane == 0) { cw[0] = 0.f; cw[S + 1] = 1.0f; taug[S + 1] = __ldg(in.far + ray); }
      __syncwarp();
      if (lane < 3) {
        const float q = (lane == 0) ? 0.05f : (lane == 1 ? 0.5f : 0.95f);
        out.distance_percentiles[(size_t)ray * 3 + lane] = interp_sorted_c(q, cw, taug, S + 2);
      }
    }
  }
}

// Backward.  With w_s = alpha_s T_s:  dL/d(dd_s) = gw_s T_{s+1} - sum_{k>s} gw_k w_k,
// dL/d density_s = dL/d(dd_s) * delta_s; the opaque last interval has no gradient.
// semantic / intensity are composited with detached weights (render.py:240-249).
__global__ void __launch_bounds__(kCompWarps * 32) k_composite_bwd(nlb_composite_in_t in,
                                                                   const float* __restrict__ weights,
                                                                   nlb_composite_grad_t g, float* __restrict__ g_density,
                                                                   float* __restrict__ g_rgb_s,
                                                                   float* __restrict__ g_sem_s,
                                                                   float* __restrict__ g_int_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCompWarps + warp;
  if (ray >= in.N) return;
  const int S = in.S, K = in.K;
  const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
              dz = __ldg(in.directions + 3 * ray + 2);
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float* td = in.tdist + (size_t)ray * (S + 1);
  const float* wr = weights + (size_t)ray * S;

  // pass 1: acc and depth numerator
  float a_acc = 0.f, a_dep = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float w = __ldg(wr + s);
    a_acc += w;
    a_dep = fmaf(w, 0.5f * (__ldg(td + s) + __ldg(td + s + 1)), a_dep);
  }
  const float acc = warp_sum(a_acc), dep = warp_sum(a_dep);
  const float den = fmaxf(acc, kEps);
  const float gr = g.g_rgb ? __ldg(g.g_rgb + 3 * (size_t)ray) : 0.f;
  const float gg = g.g_rgb ? __ldg(g.g_rgb + 3 * (size_t)ray + 1) : 0.f;
  const float gb = g.g_rgb ? __ldg(g.g_rgb + 3 * (size_t)ray + 2) : 0.f;
  const float gdep = g.g_depth ? __ldg(g.g_depth + ray) : 0.f;
  const float gacc = g.g_acc ? __ldg(g.g_acc + ray) : 0.f;
  const float gint = g.g_intensity ? __ldg(g.g_intensity + ray) : 0.f;
  // d/dw of the background term max(1-acc,0)*bg and of 1/max(acc,eps)
  const float bg_term = ((1.0f - acc) >= 0.f) ? -in.bg * (gr + gg + gb) : 0.f;
  const float den_term = (acc >= kEps) ? -gdep * dep / (den * den) : 0.f;

  // pass 2: reverse over chunks with a suffix carry
  float suffix = 0.f;  // sum_{k > chunk} gw_k w_k
  const int nchunks = (S + 31) / 32;
  // transmittance needs the forward prefix; recompute it chunk by chunk from the front
  // by first accumulating the per-chunk totals of dd.
  float chunk_before[8];  // S <= 256
  {
    float run = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      chunk_before[c] = run;
      const int s = c * 32 + lane;
      float dd = 0.f;
      if (s < S && !(in.opaque_background && s == S - 1))
        dd = __fmul_rn(__ldg(in.density + (size_t)ray * S + s), __fmul_rn(__fsub_rn(__ldg(td + s + 1), __ldg(td + s)), dnorm));
      run = __fadd_rn(run, warp_sum(dd));
    }
  }
  for (int c = nchunks - 1; c >= 0; --c) {
    const int s = c * 32 + lane;
    const bool ok = s < S;
    float t0 = 0.f, t1 = 0.f, delta = 0.f, dd = 0.f, w = 0.f;
    if (ok) {
      t0 = __ldg(td + s);
      t1 = __ldg(td + s + 1);
      delta = __fmul_rn(__fsub_rn(t1, t0), dnorm);
      dd = __fmul_rn(__ldg(in.density + (size_t)ray * S + s), delta);
      w = __ldg(wr + s);
    }
    const bool last_opaque = ok && in.opaque_background && s == S - 1;
    const float dd_scan = (ok && !last_opaque) ? dd : 0.f;
    const float incl = warp_scan_incl(dd_scan, lane);
    const float t_next = expf(-__fadd_rn(chunk_before[c], incl));  // T_{s+1}
    float gw = 0.f;
    if (ok) {
      const float tm = 0.5f * (t0 + t1);
      gw = (g.g_weights ? __ldg(g.g_weights + (size_t)ray * S + s) : 0.f) + gacc + bg_term + gdep * tm / den + den_term;
      if (in.rgb) {
        const float* cc = in.rgb + ((size_t)ray * S + s) * 3;
        gw += gr * __ldg(cc) + gg * __ldg(cc + 1) + gb * __ldg(cc + 2);
      }
      if (g_rgb_s) {
        float* o = g_rgb_s + ((size_t)ray * S + s) * 3;
        o[0] = w * gr; o[1] = w * gg; o[2] = w * gb;
      }
      if (g_int_s) g_int_s[(size_t)ray * S + s] = w * gint;
    }
    // suffix sums of gw*w within the chunk (exclusive, from the right)
    const float v = gw * w;
    float sfx = v;  // inclusive from the right
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_down_sync(NLB_FULL_MASK, sfx, o);
      if (lane + o < 32) sfx += n;
    }
    const float excl_right = sfx - v + suffix;
    if (ok) g_density[(size_t)ray * S + s] = last_opaque ? 0.f : (gw * t_next - excl_right) * delta;
    suffix += __shfl_sync(NLB_FULL_MASK, sfx, 0);
  }
  if (g_sem_s && g.g_semantic) {
    for (int i = lane; i < S * K; i += 32) {
      const int s = i / K, k = i - s * K;
      g_sem_s[(size_t)ray * S * K + i] = __ldg(wr + s) * __ldg(g.g_semantic + (size_t)ray * K + k);
    }
  }
}

}  // namespace nlb

using namespace nlb;

static int check_comp(const nlb_composite_in_t* in, const char* who) {
  if (!in) { nlb_set_error("%s: null descriptor", who); return NLB_EINVAL; }
  if (in->S < 1 || in->S > 256) { nlb_set_error("%s: S=%d outside [1,256]", who, in->S); return NLB_EUNSUPPORTED; }
  if (in->K < 0 || in->K > kMaxK) { nlb_set_error("%s: K=%d classes > %d", who, in->K, kMaxK); return NLB_EUNSUPPORTED; }
  if (in->N > 0 && (!in->density || !in->tdist || !in->directions)) { nlb_set_error("%s: null pointer", who); return NLB_EINVAL; }
  return NLB_OK;
}

extern "C" int nlb_composite_forward(const nlb_composite_in_t* in, const nlb_composite_out_t* out, void* stream) {
  if (int e = check_comp(in, "composite_forward")) return e;
  if (!out) { nlb_set_error("composite_forward: null outputs"); return NLB_EINVAL; }
  if (in->N == 0) return NLB_OK;
  if (in->compute_extras && out->distance_percentiles && !in->far) { nlb_set_error("composite_forward: far is required for the percentiles"); return NLB_EINVAL; }
  size_t smem = (size_t)kCompWarps * comp_smem_floats(in->S, in->K) * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_composite_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_composite_fwd<<<div_up(in->N, kCompWarps), kCompWarps * 32, smem, (cudaStream_t)stream>>>(*in, *out);
  return nlb_check_launch("composite_forward");
}

extern "C" int nlb_composite_backward(const nlb_composite_in_t* in, const float* weights, const nlb_composite_grad_t* g,
                                      float* g_density, float* g_rgb, float* g_semantic, float* g_intensity,
                                      void* stream) {
  if (int e = check_comp(in, "composite_backward")) return e;
  if (in->N == 0) return NLB_OK;
  if (!weights || !g || !g_density) { nlb_set_error("composite_backward: null pointer"); return NLB_EINVAL; }
  k_composite_bwd<<<div_up(in->N, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(*in, weights, *g, g_density,
                                                                                         g_rgb, g_semantic, g_intensity);
  return nlb_check_launch("composite_backward");
}
