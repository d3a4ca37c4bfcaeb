// Alpha compositing, one warp per ray:
//   render.compute_alpha_weights  (Z/internal/render.py:170-189)
//   render.volumetric_rendering   (Z/internal/render.py:192-284) incl.
//   stepfun.weighted_percentile   (Z/internal/stepfun.py:329-339)
// The reference runs ~60 eager kernels per level; here one kernel reads each
// per-sample tensor exactly once (warp scan for the transmittance, warp
// reductions for the K-channel weighted sums) and writes the per-ray outputs.
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/nlb200.h"

namespace nlb {

constexpr int kCompWarps = 4;
constexpr int kMaxK = 32;  // semantic classes handled by one lane each in the column sums
constexpr int kSemPre4 = 5;  // float4 per lane prefetched from the chunk's [32, K] class probabilities (K <= 20)

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
    // every global load of the chunk is issued here, before the first dependent instruction: a ray is one
    // short dependent chain per warp, and with the loads spread along it (after the scans and the warp
    // barriers) their latencies added up instead of overlapping
    float t0 = 0.f, t1 = 0.f, dd = 0.f, den_s = 0.f, c_r = 0.f, c_g = 0.f, c_b = 0.f, c_i = 0.f;
    // class probabilities of the chunk: [rows, K] contiguous, read as float4 (16-byte aligned when S*K % 4 == 0)
    float4 semv[kSemPre4];
    const int cnt = (in.semantic && out.semantic) ? min(32, S - c0) * K : 0;
    const float* sp = in.semantic ? in.semantic + ((size_t)ray * S + c0) * K : nullptr;
    const bool vec4 = cnt > 0 && K >= 4 && (S & 1) == 0 && ((S * K) & 3) == 0 && (cnt & 3) == 0 && cnt <= 128 * kSemPre4;
    if (vec4) {
#pragma unroll
      for (int j = 0; j < kSemPre4; ++j)
        semv[j] = (4 * (lane + 32 * j) < cnt) ? __ldg(reinterpret_cast<const float4*>(sp) + lane + 32 * j)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (ok) {
      t0 = __ldg(td + s);
      t1 = __ldg(td + s + 1);
      den_s = __ldg(in.density + (size_t)ray * S + s);
      if (in.rgb) {
        const float* c = in.rgb + ((size_t)ray * S + s) * 3;
        c_r = __ldg(c); c_g = __ldg(c + 1); c_b = __ldg(c + 2);
      }
      if (in.intensity) c_i = __ldg(in.intensity + (size_t)ray * S + s);
      dd = __fmul_rn(den_s, __fmul_rn(__fsub_rn(t1, t0), dnorm));
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
        a_r = fmaf(w, c_r, a_r);
        a_g = fmaf(w, c_g, a_g);
        a_b = fmaf(w, c_b, a_b);
      }
      if (in.intensity) a_int = fmaf(w, c_i, a_int);
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
      // element e = row * K + class; row / class advance incrementally (no division by the run-time K)
      if (vec4) {
        int row = (4 * lane) / K, cls = 4 * lane - row * K;
        const int drow = 128 / K, dcls = 128 - drow * K;
#pragma unroll
        for (int j = 0; j < kSemPre4; ++j) {
          const int e = 4 * (lane + 32 * j);
          if (e < cnt) {
            const float w0 = wchunk[row];
            const float w1 = wchunk[row + (cls + 1 >= K)], w2 = wchunk[row + (cls + 2 >= K)], w3 = wchunk[row + (cls + 3 >= K)];
            *reinterpret_cast<float4*>(prod + e) = make_float4(w0 * semv[j].x, w1 * semv[j].y, w2 * semv[j].z, w3 * semv[j].w);
          }
          row += drow;
          cls += dcls;
          if (cls >= K) { cls -= K; ++row; }
        }
      } else {
        int row = lane / K, cls = lane - row * K;
        const int drow = 32 / K, dcls = 32 - drow * K;
        for (int i = lane; i < cnt; i += 32) {
          prod[i] = wchunk[row] * __ldg(sp + i);
          row += drow;
          cls += dcls;
          if (cls >= K) { cls -= K; ++row; }
        }
      }
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

// Proposal levels (no colour / class / intensity channels): one THREAD per ray.  With a warp per ray the
// fixed per-ray cost (two warp scans per chunk, eight warp reductions, the percentile search) is ~1000 warp
// instructions for 820 algorithmic bytes and the kernel is issue-bound at 12 % of the HBM roofline (ncu at
// 1 M rays); a thread walks its ray sequentially in ~60 warp instructions per ray.  Coalescing comes from
// staging: a warp loads 32 rays x 32 samples with lane = sample into padded shared-memory rows, then every
// thread reads its own row; the weights leave the same way.
constexpr int kPropRays = 128;

__global__ void __launch_bounds__(kPropRays) k_composite_prop_fwd(nlb_composite_in_t in, nlb_composite_out_t out) {
  __shared__ float s_den[kPropRays][33];  // densities of the chunk, overwritten by the weights
  __shared__ float s_t[kPropRays][34];    // fenceposts c0 .. c0 + 32
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = in.S;
  const int ray0 = blockIdx.x * kPropRays;
  const int ray = ray0 + tid;
  const bool rok = ray < in.N;
  float dnorm = 0.f;
  if (rok) {
    const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
                dz = __ldg(in.directions + 3 * ray + 2);
    dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  }
  float carry = 0.f, acc = 0.f, dep = 0.f, lg = 0.f, wsum = 0.f;
  float prev_cw = 0.f, prev_tk = 0.f, t_first = 0.f, t_last = 0.f;
  const float qs[3] = {0.05f, 0.5f, 0.95f};
  float pct[3] = {0.f, 0.f, 0.f};
  bool found[3] = {false, false, false};
  auto knot = [&](int k, float x1, float f1) {  // interp_sorted_c between the previous knot and (x1, f1)
    float off = __fdiv_rn(__fsub_rn(qs[k], prev_cw), __fsub_rn(x1, prev_cw));
    if (isnan(off)) off = 0.f;
    off = fminf(fmaxf(off, 0.f), 1.f);
    pct[k] = __fadd_rn(prev_tk, __fmul_rn(off, __fsub_rn(f1, prev_tk)));
    found[k] = true;
  };
  for (int c0 = 0; c0 < S; c0 += 32) {
    const int n = min(32, S - c0);
    // stage: this warp's 32 rays, lane = sample; asynchronous 4-byte copies, so the 64+ loads of a lane are
    // all in flight together (a load -> store loop paid the global latency once per ray row)
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int row = warp * 32 + rr, gr = ray0 + row;
      if (gr < in.N) {
        if (lane < n) umma::cp_async4(&s_den[row][lane], in.density + (size_t)gr * S + c0 + lane);
        if (lane <= n) umma::cp_async4(&s_t[row][lane], in.tdist + (size_t)gr * (S + 1) + c0 + lane);
        if (lane == 0 && n == 32) umma::cp_async4(&s_t[row][32], in.tdist + (size_t)gr * (S + 1) + c0 + 32);
      }
    }
    umma::cp_async_wait_all();
    __syncwarp();
    if (rok) {
      if (c0 == 0) { t_first = s_t[tid][0]; prev_tk = t_first; }
      for (int j = 0; j < n; ++j) {
        const float t0 = s_t[tid][j], t1 = s_t[tid][j + 1];
        float dd = __fmul_rn(s_den[tid][j], __fmul_rn(__fsub_rn(t1, t0), dnorm));
        const bool last = in.opaque_background && (c0 + j == S - 1);
        const float trans = expf(-carry);
        if (last) dd = INFINITY; else carry = __fadd_rn(carry, dd);
        const float alpha = __fsub_rn(1.0f, expf(-dd));
        const float w = __fmul_rn(alpha, trans);
        s_den[tid][j] = w;
        const float tm = __fmul_rn(0.5f, __fadd_rn(t0, t1));
        acc += w;
        dep = fmaf(w, tm, dep);
        if (in.compute_extras) {
          lg = fmaf(w, logf(tm), lg);
          wsum = __fadd_rn(wsum, w);
          const float cwj = fminf(wsum, 1.0f);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (!found[k] && cwj > qs[k]) knot(k, cwj, t1);
          prev_cw = cwj;
          prev_tk = t1;
        }
        t_last = t1;
      }
    }
    __syncwarp();
    if (out.weights) {
      for (int rr = 0; rr < 32; ++rr) {
        const int row = warp * 32 + rr, gr = ray0 + row;
        if (gr < in.N && lane < n) out.weights[(size_t)gr * S + c0 + lane] = s_den[row][lane];
      }
    }
    __syncwarp();
  }
  if (!rok) return;
  const float bg_w = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  const float den = fmaxf(acc, kEps);
  if (out.rgb) {
    const float v = fmaf(bg_w, in.bg, 0.f);
    out.rgb[3 * (size_t)ray] = v;
    out.rgb[3 * (size_t)ray + 1] = v;
    out.rgb[3 * (size_t)ray + 2] = v;
  }
  if (out.depth) out.depth[ray] = __fdiv_rn(dep, den);
  if (out.acc) out.acc[ray] = acc;
  if (in.compute_extras) {
    if (out.distance_mean) {
      float v = expf(__fdiv_rn(lg, den));
      if (isnan(v)) v = INFINITY;
      out.distance_mean[ray] = fminf(fmaxf(v, t_first), t_last);
    }
    if (out.distance_percentiles) {
      const float far = __ldg(in.far + ray);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!found[k]) knot(k, 1.0f, far);  // the closing knot (cw = 1, t = far)
        out.distance_percentiles[(size_t)ray * 3 + k] = pct[k];
      }
    }
  }
}

// Proposal levels with S % 4 == 0: the same walk with direct 16-byte loads of the densities and stores of
// the weights (no staging instructions); only the fenceposts ([N, S+1] rows are not 16-byte aligned) are
// scalar loads.
// CS = samples per iteration: with CS = 8 a thread consumes (and writes) one whole 32-byte sector of its density /
// weight row per iteration instead of relying on L1 to keep the second half of the sector until the next one
// (the working set of a resident SM, 2048 rays x 772 B, is six times L1).
template <int CS>
__global__ void __launch_bounds__(128) k_composite_prop4_fwd(nlb_composite_in_t in, nlb_composite_out_t out) {
  const int S = in.S;
  const int ray = blockIdx.x * 128 + threadIdx.x;
  if (ray >= in.N) return;
  const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
              dz = __ldg(in.directions + 3 * ray + 2);
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const float* td = in.tdist + (size_t)ray * (S + 1);
  float carry = 0.f, acc = 0.f, dep = 0.f, lg = 0.f, wsum = 0.f;
  float t_prev = __ldg(td);
  const float t_first = t_prev;
  float prev_cw = 0.f, prev_tk = t_first, t_last = t_first;
  const float qs[3] = {0.05f, 0.5f, 0.95f};
  float pct[3] = {0.f, 0.f, 0.f};
  bool found[3] = {false, false, false};
  auto knot = [&](int k, float x1, float f1) {
    float off = __fdiv_rn(__fsub_rn(qs[k], prev_cw), __fsub_rn(x1, prev_cw));
    if (isnan(off)) off = 0.f;
    off = fminf(fmaxf(off, 0.f), 1.f);
    pct[k] = __fadd_rn(prev_tk, __fmul_rn(off, __fsub_rn(f1, prev_tk)));
    found[k] = true;
  };
#pragma unroll(CS == 4 ? 2 : 1)
  for (int c0 = 0; c0 < S; c0 += CS) {
    float dens[CS];
#pragma unroll
    for (int q = 0; q < CS / 4; ++q) {
      const float4 dv = __ldg(reinterpret_cast<const float4*>(in.density + (size_t)ray * S + c0 + 4 * q));
      dens[4 * q] = dv.x; dens[4 * q + 1] = dv.y; dens[4 * q + 2] = dv.z; dens[4 * q + 3] = dv.w;
    }
    float tt[CS + 1];
    tt[0] = t_prev;
#pragma unroll
    for (int j = 1; j <= CS; ++j) tt[j] = __ldg(td + c0 + j);
    t_prev = tt[CS];
    float w[CS];
#pragma unroll
    for (int j = 0; j < CS; ++j) {
      const float t0 = tt[j], t1 = tt[j + 1];
      float dd = __fmul_rn(dens[j], __fmul_rn(__fsub_rn(t1, t0), dnorm));
      const bool last = in.opaque_background && (c0 + j == S - 1);
      const float trans = expf(-carry);
      if (last) dd = INFINITY; else carry = __fadd_rn(carry, dd);
      w[j] = __fmul_rn(__fsub_rn(1.0f, expf(-dd)), trans);
      const float tm = __fmul_rn(0.5f, __fadd_rn(t0, t1));
      acc += w[j];
      dep = fmaf(w[j], tm, dep);
      if (in.compute_extras) {
        lg = fmaf(w[j], logf(tm), lg);
        wsum = __fadd_rn(wsum, w[j]);
        const float cwj = fminf(wsum, 1.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (!found[k] && cwj > qs[k]) knot(k, cwj, t1);
        prev_cw = cwj;
        prev_tk = t1;
      }
      t_last = t1;
    }
    if (out.weights) {
#pragma unroll
      for (int q = 0; q < CS / 4; ++q)
        *reinterpret_cast<float4*>(out.weights + (size_t)ray * S + c0 + 4 * q) = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  }
  const float bg_w = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  const float den = fmaxf(acc, kEps);
  if (out.rgb) {
    const float v = fmaf(bg_w, in.bg, 0.f);
    out.rgb[3 * (size_t)ray] = v;
    out.rgb[3 * (size_t)ray + 1] = v;
    out.rgb[3 * (size_t)ray + 2] = v;
  }
  if (out.depth) out.depth[ray] = __fdiv_rn(dep, den);
  if (out.acc) out.acc[ray] = acc;
  if (in.compute_extras) {
    if (out.distance_mean) {
      float v = expf(__fdiv_rn(lg, den));
      if (isnan(v)) v = INFINITY;
      out.distance_mean[ray] = fminf(fmaxf(v, t_first), t_last);
    }
    if (out.distance_percentiles) {
      const float far = __ldg(in.far + ray);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!found[k]) knot(k, 1.0f, far);
        out.distance_percentiles[(size_t)ray * 3 + k] = pct[k];
      }
    }
  }
}

// Proposal levels, S % 8 == 0, 16-byte loads for the fenceposts as well.  The [N, S+1] fencepost rows are not
// 16-byte aligned (a row starts 4 * (ray mod 4) bytes past a 16-byte boundary), so the kernel above reads them
// with S scalar loads per ray -- and with a thread per ray every load instruction touches 32 different lines,
// which L1 retires at one per clock whatever the access width (tools/gather_probe.cu): 81 load + 16 store
// instructions x 32 cycles per 32 rays = 0.29 ms per 1 M rays, the measured time.  Here a row is covered by 17
// ALIGNED float4 loads (two new ones per 8-sample chunk, the third carried) and the lane's misalignment
// A = ray mod 4 is resolved in registers with selects: 33 loads + 16 stores per ray.  (Making A a compile-time
// constant per warp -- warp w of a block takes the rays with ray mod 4 == w -- was measured first: 0.50 ms, the
// interleaved rays of a warp cost more in DRAM / L2 locality than the selects cost in issue slots.)
__device__ __forceinline__ void composite_prop_aligned(const nlb_composite_in_t& in, const nlb_composite_out_t& out, int ray) {
  const int S = in.S;
  const int A = ray & 3;
  const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
              dz = __ldg(in.directions + 3 * ray + 2);
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  // float4 view of the fencepost matrix: this row's element j sits at float index row0 + j = 4 * g0 + A + j
  const size_t row0 = (size_t)ray * (S + 1);
  const size_t g0 = (row0 - A) >> 2;
  const size_t total = (size_t)in.N * (S + 1);
  const float4* t4 = reinterpret_cast<const float4*>(in.tdist);
  auto load_group = [&](size_t g) -> float4 {
    if ((g + 1) * 4 <= total) return __ldg(t4 + g);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);   // the very last group of the matrix: element-wise, bounded
    if (g * 4 + 0 < total) r.x = __ldg(in.tdist + g * 4 + 0);
    if (g * 4 + 1 < total) r.y = __ldg(in.tdist + g * 4 + 1);
    if (g * 4 + 2 < total) r.z = __ldg(in.tdist + g * 4 + 2);
    return r;
  };
  float carry = 0.f, acc = 0.f, dep = 0.f, lg = 0.f, wsum = 0.f;
  float4 q0 = load_group(g0);
  const float t_first = A == 0 ? q0.x : (A == 1 ? q0.y : (A == 2 ? q0.z : q0.w));
#define NLB_SEL(k) (A == 0 ? f[(k)] : (A == 1 ? f[(k) + 1] : (A == 2 ? f[(k) + 2] : f[(k) + 3])))
  float prev_cw = 0.f, prev_tk = t_first, t_last = t_first;
  const float qs[3] = {0.05f, 0.5f, 0.95f};
  float pct[3] = {0.f, 0.f, 0.f};
  bool found[3] = {false, false, false};
  auto knot = [&](int k, float x1, float f1) {
    float off = __fdiv_rn(__fsub_rn(qs[k], prev_cw), __fsub_rn(x1, prev_cw));
    if (isnan(off)) off = 0.f;
    off = fminf(fmaxf(off, 0.f), 1.f);
    pct[k] = __fadd_rn(prev_tk, __fmul_rn(off, __fsub_rn(f1, prev_tk)));
    found[k] = true;
  };
#pragma unroll 1
  for (int c0 = 0; c0 < S; c0 += 8) {
    const float4 q1 = load_group(g0 + (c0 >> 2) + 1), q2 = load_group(g0 + (c0 >> 2) + 2);
    const float f[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    q0 = q2;
    float dens[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 dv = __ldg(reinterpret_cast<const float4*>(in.density + (size_t)ray * S + c0 + 4 * q));
      dens[4 * q] = dv.x; dens[4 * q + 1] = dv.y; dens[4 * q + 2] = dv.z; dens[4 * q + 3] = dv.w;
    }
    float tt[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) tt[j] = NLB_SEL(j);
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t0 = tt[j], t1 = tt[j + 1];
      float dd = __fmul_rn(dens[j], __fmul_rn(__fsub_rn(t1, t0), dnorm));
      const bool last = in.opaque_background && (c0 + j == S - 1);
      const float trans = expf(-carry);
      if (last) dd = INFINITY; else carry = __fadd_rn(carry, dd);
      w[j] = __fmul_rn(__fsub_rn(1.0f, expf(-dd)), trans);
      const float tm = __fmul_rn(0.5f, __fadd_rn(t0, t1));
      acc += w[j];
      dep = fmaf(w[j], tm, dep);
      if (in.compute_extras) {
        lg = fmaf(w[j], logf(tm), lg);
        wsum = __fadd_rn(wsum, w[j]);
        const float cwj = fminf(wsum, 1.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (!found[k] && cwj > qs[k]) knot(k, cwj, t1);
        prev_cw = cwj;
        prev_tk = t1;
      }
      t_last = t1;
    }
    if (out.weights) {
      *reinterpret_cast<float4*>(out.weights + (size_t)ray * S + c0) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(out.weights + (size_t)ray * S + c0 + 4) = make_float4(w[4], w[5], w[6], w[7]);
    }
  }
  const float bg_w = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  const float den = fmaxf(acc, kEps);
  if (out.rgb) {
    const float v = fmaf(bg_w, in.bg, 0.f);
    out.rgb[3 * (size_t)ray] = v;
    out.rgb[3 * (size_t)ray + 1] = v;
    out.rgb[3 * (size_t)ray + 2] = v;
  }
  if (out.depth) out.depth[ray] = __fdiv_rn(dep, den);
  if (out.acc) out.acc[ray] = acc;
  if (in.compute_extras) {
    if (out.distance_mean) {
      float v = expf(__fdiv_rn(lg, den));
      if (isnan(v)) v = INFINITY;
      out.distance_mean[ray] = fminf(fmaxf(v, t_first), t_last);
    }
    if (out.distance_percentiles) {
      const float far = __ldg(in.far + ray);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!found[k]) knot(k, 1.0f, far);
        out.distance_percentiles[(size_t)ray * 3 + k] = pct[k];
      }
    }
  }
}

#undef NLB_SEL
__global__ void __launch_bounds__(128) k_composite_prop_aligned_fwd(nlb_composite_in_t in, nlb_composite_out_t out) {
  const int ray = blockIdx.x * 128 + threadIdx.x;
  if (ray >= in.N) return;
  composite_prop_aligned(in, out, ray);
}

// NeRF level with the nuScenes head layout (rgb + K = 19 class probabilities + intensity): one thread per
// ray as well, reading its own rows straight from global memory with 16-byte loads: a ray's 8-sample chunk
// is 608 contiguous bytes of class probabilities, 96 of colour, 32 of density and of intensity, so every
// 32-byte sector a thread touches is fully used (its second half is an L1 hit) and a chunk's ~50 independent
// loads are in flight together.  With K fixed at compile time the (sample, class) of every element of the
// 152-float class block is known and the 19 class sums live in registers.  (The warp-per-ray kernel above
// needs ~900 warp instructions per ray -- index arithmetic, a 32-step column sum in shared memory, eight warp
// reductions -- and is issue-bound at 48 % of the HBM roofline at 1 M rays; staging the rows through shared
// memory with cp.async cost 560 warp instructions per ray in address arithmetic and was slower still.  The
// warp kernel remains the path for other layouts.)
#ifndef NLB_RAYCS
#define NLB_RAYCS 4
#endif
constexpr int kRayK = 19, kRayCS = NLB_RAYCS, kRayThreads = 128;
constexpr int kSem4 = kRayCS * kRayK / 4, kRgb4 = kRayCS * 3 / 4, kVec4 = kRayCS / 4;  // float4 per chunk
static_assert(kRayCS % 4 == 0, "chunk = whole float4 of every per-sample tensor");

__global__ void __launch_bounds__(kRayThreads) k_composite_ray19_fwd(nlb_composite_in_t in, nlb_composite_out_t out) {
  const int S = in.S;
  const int ray = blockIdx.x * kRayThreads + threadIdx.x;
  if (ray >= in.N) return;
  const float dx = __ldg(in.directions + 3 * ray), dy = __ldg(in.directions + 3 * ray + 1),
              dz = __ldg(in.directions + 3 * ray + 2);
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  float carry = 0.f, acc = 0.f, dep = 0.f, lg = 0.f, wsum = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_int = 0.f;
  float a_sem[kRayK];
#pragma unroll
  for (int k = 0; k < kRayK; ++k) a_sem[k] = 0.f;
  const float* td = in.tdist + (size_t)ray * (S + 1);
  float t_prev = __ldg(td);
  const float t_first = t_prev;
  float prev_cw = 0.f, prev_tk = t_first, t_last = t_first;
  const float qs[3] = {0.05f, 0.5f, 0.95f};
  float pct[3] = {0.f, 0.f, 0.f};
  bool found[3] = {false, false, false};
  auto knot = [&](int k, float x1, float f1) {
    float off = __fdiv_rn(__fsub_rn(qs[k], prev_cw), __fsub_rn(x1, prev_cw));
    if (isnan(off)) off = 0.f;
    off = fminf(fmaxf(off, 0.f), 1.f);
    pct[k] = __fadd_rn(prev_tk, __fmul_rn(off, __fsub_rn(f1, prev_tk)));
    found[k] = true;
  };
#pragma unroll 1
  for (int c0 = 0; c0 < S; c0 += kRayCS) {
    const size_t e0 = (size_t)ray * S + c0;  // first sample of the chunk
    // ---- all loads of the chunk first
    const float4* sem4 = reinterpret_cast<const float4*>(in.semantic + e0 * kRayK);
    const float4* rgb4 = reinterpret_cast<const float4*>(in.rgb + e0 * 3);
    float4 sv[kSem4], cv[kRgb4], dv[kVec4], iv[kVec4];
#pragma unroll
    for (int m = 0; m < kSem4; ++m) sv[m] = __ldg(sem4 + m);
#pragma unroll
    for (int m = 0; m < kRgb4; ++m) cv[m] = __ldg(rgb4 + m);
#pragma unroll
    for (int m = 0; m < kVec4; ++m) {
      dv[m] = __ldg(reinterpret_cast<const float4*>(in.density + e0) + m);
      iv[m] = __ldg(reinterpret_cast<const float4*>(in.intensity + e0) + m);
    }
    float tt[kRayCS + 1];
    tt[0] = t_prev;
#pragma unroll
    for (int j = 1; j <= kRayCS; ++j) tt[j] = __ldg(td + c0 + j);
    t_prev = tt[kRayCS];
    float dens[kRayCS], ints[kRayCS], w[kRayCS];
#pragma unroll
    for (int m = 0; m < kVec4; ++m) {
      dens[4 * m] = dv[m].x; dens[4 * m + 1] = dv[m].y; dens[4 * m + 2] = dv[m].z; dens[4 * m + 3] = dv[m].w;
      ints[4 * m] = iv[m].x; ints[4 * m + 1] = iv[m].y; ints[4 * m + 2] = iv[m].z; ints[4 * m + 3] = iv[m].w;
    }
#pragma unroll
    for (int j = 0; j < kRayCS; ++j) {
      const float t0 = tt[j], t1 = tt[j + 1];
      float dd = __fmul_rn(dens[j], __fmul_rn(__fsub_rn(t1, t0), dnorm));
      const bool last = in.opaque_background && (c0 + j == S - 1);
      const float trans = expf(-carry);
      if (last) dd = INFINITY; else carry = __fadd_rn(carry, dd);
      w[j] = __fmul_rn(__fsub_rn(1.0f, expf(-dd)), trans);
      const float tm = __fmul_rn(0.5f, __fadd_rn(t0, t1));
      acc += w[j];
      dep = fmaf(w[j], tm, dep);
      a_int = fmaf(w[j], ints[j], a_int);
      if (in.compute_extras) {
        lg = fmaf(w[j], logf(tm), lg);
        wsum = __fadd_rn(wsum, w[j]);
        const float cwj = fminf(wsum, 1.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (!found[k] && cwj > qs[k]) knot(k, cwj, t1);
        prev_cw = cwj;
        prev_tk = t1;
      }
      t_last = t1;
    }
    // colours: 24 floats = samples x (r, g, b)
#pragma unroll
    for (int m = 0; m < kRgb4; ++m) {
      const float e[4] = {cv[m].x, cv[m].y, cv[m].z, cv[m].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = 4 * m + q, j = idx / 3, ch = idx - 3 * j;
        if (ch == 0) a_r = fmaf(w[j], e[q], a_r);
        else if (ch == 1) a_g = fmaf(w[j], e[q], a_g);
        else a_b = fmaf(w[j], e[q], a_b);
      }
    }
    // class probabilities: 152 floats = samples x 19 classes (weights detached: render.py:240-249)
#pragma unroll
    for (int m = 0; m < kSem4; ++m) {
      const float e[4] = {sv[m].x, sv[m].y, sv[m].z, sv[m].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = 4 * m + q, j = idx / kRayK, k = idx - kRayK * j;
        a_sem[k] = fmaf(w[j], e[q], a_sem[k]);
      }
    }
    if (out.weights) {
      float4* dst = reinterpret_cast<float4*>(out.weights + e0);
#pragma unroll
      for (int m = 0; m < kVec4; ++m) dst[m] = make_float4(w[4 * m], w[4 * m + 1], w[4 * m + 2], w[4 * m + 3]);
    }
  }
  const float bg_w = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  const float den = fmaxf(acc, kEps);
  if (out.rgb) {
    out.rgb[3 * (size_t)ray] = fmaf(bg_w, in.bg, a_r);
    out.rgb[3 * (size_t)ray + 1] = fmaf(bg_w, in.bg, a_g);
    out.rgb[3 * (size_t)ray + 2] = fmaf(bg_w, in.bg, a_b);
  }
  if (out.depth) out.depth[ray] = __fdiv_rn(dep, den);
  if (out.acc) out.acc[ray] = acc;
  if (out.intensity) out.intensity[ray] = a_int;
  if (out.semantic) {
#pragma unroll
    for (int k = 0; k < kRayK; ++k) out.semantic[(size_t)ray * kRayK + k] = a_sem[k];
  }
  if (in.compute_extras) {
    if (out.distance_mean) {
      float v = expf(__fdiv_rn(lg, den));
      if (isnan(v)) v = INFINITY;
      out.distance_mean[ray] = fminf(fmaxf(v, t_first), t_last);
    }
    if (out.distance_percentiles) {
      const float far = __ldg(in.far + ray);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!found[k]) knot(k, 1.0f, far);
        out.distance_percentiles[(size_t)ray * 3 + k] = pct[k];
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
  // thread-per-ray kernels need >= ~32 K rays to fill the machine (148 SMs x 2048 threads / a few); a training
  // batch (10 240 rays) is served faster by the warp-per-ray kernel (0.016 against 0.026 ms at S = 64)
  const bool many_rays = in->N >= 32768;
  if (many_rays && !in->rgb && !in->semantic && !in->intensity) {  // proposal levels: thread per ray
    static const bool kDirect = getenv("NLB_PROP_COMPOSITE_STAGED") == nullptr;
    if (kDirect && in->S % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(in->density) | reinterpret_cast<uintptr_t>(out->weights)) % 16 == 0)
    {
      static const int kChunk = [] { const char* e = getenv("NLB_PROP_COMPOSITE_CHUNK"); return e ? atoi(e) : 8; }();  // A/B timing
      if (kChunk == 8 && in->S % 8 == 0 && reinterpret_cast<uintptr_t>(in->tdist) % 16 == 0)
        k_composite_prop_aligned_fwd<<<div_up(in->N, 128), 128, 0, (cudaStream_t)stream>>>(*in, *out);
      else if (kChunk == 16 && in->S % 8 == 0)
        k_composite_prop4_fwd<8><<<div_up(in->N, 128), 128, 0, (cudaStream_t)stream>>>(*in, *out);
      else if (kChunk == 116 && in->S % 16 == 0)
        k_composite_prop4_fwd<16><<<div_up(in->N, 128), 128, 0, (cudaStream_t)stream>>>(*in, *out);
      else if (kChunk == 132 && in->S % 32 == 0)
        k_composite_prop4_fwd<32><<<div_up(in->N, 128), 128, 0, (cudaStream_t)stream>>>(*in, *out);
      else
        k_composite_prop4_fwd<4><<<div_up(in->N, 128), 128, 0, (cudaStream_t)stream>>>(*in, *out);
    }
    else
      k_composite_prop_fwd<<<div_up(in->N, kPropRays), kPropRays, 0, (cudaStream_t)stream>>>(*in, *out);
    return nlb_check_launch("composite_forward");
  }
  if (many_rays && in->rgb && in->semantic && in->intensity && in->K == kRayK && in->S % kRayCS == 0 && out->semantic && out->intensity &&
      (reinterpret_cast<uintptr_t>(in->rgb) | reinterpret_cast<uintptr_t>(in->semantic) | reinterpret_cast<uintptr_t>(in->intensity) |
       reinterpret_cast<uintptr_t>(in->density) | reinterpret_cast<uintptr_t>(out->weights)) % 16 == 0) {
    k_composite_ray19_fwd<<<div_up(in->N, kRayThreads), kRayThreads, 0, (cudaStream_t)stream>>>(*in, *out);
    return nlb_check_launch("composite_forward");
  }
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
