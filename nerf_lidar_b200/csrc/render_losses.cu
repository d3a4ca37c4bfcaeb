// Supervision losses of the training step on the final rendering, value AND gradient in
// four launches (Z/train.py:283-455, Z/internal/train_utils.py:55-123,329-345,412-431):
//   data    Charbonnier / MSE on rgb over the camera rays
//   depth   log(|d| + 1) over the depth-supervised rays below the 0.9 quantile of |d|
//   sem     -log(p[label] + 1e-6) over the labelled camera rays
//   int     squared intensity error over the LiDAR rays
//   d_smo / s_smo  edge-aware smoothness of depth / semantics on the 32x32 patches
// The reference (and this repository's plain-torch restatement, kept for cross-checking)
// spends ~350 elementwise / reduction launches on these per step.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

enum { S_DATA_NUM = 0, S_DATA_DEN, S_DEPTH_NUM, S_DEPTH_DEN, S_SEM_NUM, S_SEM_DEN, S_INT_NUM, S_INT_DEN,
       S_DSMO_X, S_DSMO_Y, S_SSMO_X, S_SSMO_Y, S_THRE, S_CNT_X, S_CNT_Y, S_LAYOUT_ERR, kNumSums = 16 };

struct RayMasks {
  bool rgb, depth, sem, lidar;
};

__device__ __forceinline__ RayMasks ray_masks(const nlb_losses_in_t& in, int i) {
  const bool patch = __ldg(in.patch_mask + i) == 1.0f;
  const bool lidar = __ldg(in.lidar_mask + i) == 1.0f;
  const bool valid = in.ray_valid == nullptr || __ldg(in.ray_valid + i) != 0.0f;
  RayMasks m;
  m.lidar = lidar;
  m.rgb = valid && !patch;
  m.depth = (__ldg(in.t_depth + i) > 0.f) && m.rgb;
  m.sem = in.semantic != nullptr && (__ldg(in.t_semantic + i) != 255.0f) && m.rgb;
  if (in.lidar_supervision) {
    m.rgb = m.rgb && !lidar;
    m.depth = m.depth || lidar;
    m.sem = m.sem && !lidar;
    if (in.only_lidar_supervision) m.depth = m.depth && lidar;
  }
  return m;
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {  // all threads get the sum
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += s_red[w];
  return t;
}

// ---- 0.9 quantile (linear interpolation, torch.quantile) of |pred - target| over the
// depth-supervised rays: one block, exact k-th order statistics by a bitwise search on the
// float bit patterns (non-negative floats order like unsigned integers).
__global__ void __launch_bounds__(1024) k_depth_quantile(nlb_losses_in_t in, float* __restrict__ sums, int cached) {
  extern __shared__ uint32_t s_bits[];  // [N] |d| bit patterns (0xffffffff: not supervised) when they fit
  __shared__ float s_red[32];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_pick[2];  // chosen digit, elements below it
  const int N = in.N;
  auto bits_of = [&](int i) -> uint32_t {
    if (!ray_masks(in, i).depth) return 0xffffffffu;
    return __float_as_uint(fabsf(__ldg(in.depth + i) - __ldg(in.t_depth + i)));
  };
  float cnt = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t b = bits_of(i);
    if (cached) s_bits[i] = b;
    cnt += b != 0xffffffffu ? 1.f : 0.f;
  }
  const int n = (int)block_sum(cnt, s_red);
  if (n == 0) {
    if (threadIdx.x == 0) sums[S_THRE] = INFINITY;
    return;
  }
  const float pos = 0.9f * fmaxf((float)n - 1.0f, 0.f);
  const int k_lo = min(max((int)floorf(pos), 0), N - 1), k_hi = min(max((int)ceilf(pos), 0), N - 1);
  float sel[2];
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    if (which && k_hi == k_lo) { sel[1] = sel[0]; break; }
    // radix select, 8 bits per pass: histogram of the next digit among the elements that share the
    // prefix found so far, then the digit in which the k-th smallest falls
    uint32_t k = which ? k_hi : k_lo, prefix = 0, prefix_mask = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const uint32_t b = cached ? s_bits[i] : bits_of(i);
        if (b != 0xffffffffu && (b & prefix_mask) == prefix) atomicAdd(&s_hist[(b >> shift) & 255u], 1u);
      }
      __syncthreads();
      // the digit d in which the k-th smallest falls (first d < 255 with below(d) + hist[d] > k, else 255) by warp 0:
      // a lane owns 8 bins, a shuffle scan gives the elements below them (one thread walking the 256 bins was
      // 4 us per pass, 32 of the kernel's 42 us)
      if (threadIdx.x < 32) {
        const int l = threadIdx.x;
        uint32_t h[8], loc = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { h[q] = s_hist[l * 8 + q]; loc += h[q]; }
        uint32_t inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(NLB_FULL_MASK, inc, o);
          if (l >= o) inc += t;
        }
        const uint32_t exc = inc - loc;
        const unsigned owners = __ballot_sync(NLB_FULL_MASK, k >= exc && k < inc);
        const int owner = owners ? __ffs(owners) - 1 : 31;
        if (l == owner) {
          uint32_t below = exc;
          int q = 0;
          for (; q < 8; ++q) {
            if (l * 8 + q >= 255 || below + h[q] > k) break;
            below += h[q];
          }
          s_pick[0] = (uint32_t)(l * 8 + q);
          s_pick[1] = below;
        }
      }
      __syncthreads();
      prefix |= s_pick[0] << shift;
      prefix_mask |= 255u << shift;
      k -= s_pick[1];
      __syncthreads();
    }
    sel[which] = __uint_as_float(prefix);
  }
  if (threadIdx.x == 0) sums[S_THRE] = sel[0] + (sel[1] - sel[0]) * (pos - floorf(pos));
}

// ---- per-ray terms: partial sums -> sums[], unnormalised gradients -> g_*
__global__ void __launch_bounds__(256) k_ray_losses(nlb_losses_in_t in, float* __restrict__ sums,
                                                   float* __restrict__ g_rgb, float* __restrict__ g_depth,
                                                   float* __restrict__ g_sem, float* __restrict__ g_int) {
  __shared__ float s_red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  float cnt_x = 0.f, cnt_y = 0.f, layout_err = 0.f;
  if (i < in.N) {
    const RayMasks m = ray_masks(in, i);
    // smoothness edges whose two pixels are valid (edge_aware_loss_v2's mask_x / mask_y), and the layout the
    // patch kernel relies on: patch rays == the leading num_patch * P * P rows
    const int P = in.patch_size, n_patch_rays = in.num_patch * P * P;
    if ((i < n_patch_rays) != (__ldg(in.patch_mask + i) == 1.0f)) layout_err = 1.f;
    if (i < n_patch_rays) {
      auto ok = [&](int r) { return in.ray_valid == nullptr || __ldg(in.ray_valid + r) != 0.0f; };
      const int t = i % (P * P), py = t / P, px = t - py * P;
      if (ok(i)) {
        if (px + 1 < P && ok(i + 1)) cnt_x = 1.f;
        if (py + 1 < P && ok(i + P)) cnt_y = 1.f;
      }
    }
    // data
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float resid = __ldg(in.rgb + 3 * i + c) - __ldg(in.t_rgb + 3 * i + c);
      float per, dper;
      if (in.charb) {
        per = sqrtf(resid * resid + in.charb_padding * in.charb_padding);
        dper = resid / per;
      } else {
        per = resid * resid;
        dper = 2.0f * resid;
      }
      if (m.rgb) { acc[S_DATA_NUM] += per; acc[S_DATA_DEN] += 1.f; }
      g_rgb[3 * i + c] = m.rgb ? dper : 0.f;
    }
    // depth
    {
      const float d = __ldg(in.depth + i) - __ldg(in.t_depth + i);
      const bool keep = m.depth && (d < sums[S_THRE]);
      const float ad = fabsf(d);
      if (keep) { acc[S_DEPTH_NUM] += logf(ad + 1.0f); acc[S_DEPTH_DEN] += 1.f; }
      g_depth[i] = keep ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / (ad + 1.0f) : 0.f;
    }
    // semantic cross-entropy on the rendered class probabilities
    if (in.semantic) {
      const int label = m.sem ? (int)__ldg(in.t_semantic + i) : 0;
      for (int c = 0; c < in.K; ++c) {
        float g = 0.f;
        if (m.sem && c == label) {
          const float p = __ldg(in.semantic + (size_t)i * in.K + c) + 1e-6f;
          acc[S_SEM_NUM] += -logf(p);
          g = -1.0f / p;
        }
        g_sem[(size_t)i * in.K + c] = g;
      }
      if (m.sem) acc[S_SEM_DEN] += 1.f;
    }
    // intensity
    if (in.intensity) {
      const float diff = __ldg(in.intensity + i) - __ldg(in.t_intensity + i);
      if (m.lidar) { acc[S_INT_NUM] += diff * diff; acc[S_INT_DEN] += 1.f; }
      g_int[i] = m.lidar ? 2.0f * diff : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float t = block_sum(acc[k], s_red);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(sums + k, t);
  }
  if (in.num_patch > 0) {
    const float tx = block_sum(cnt_x, s_red), ty = block_sum(cnt_y, s_red), te = block_sum(layout_err, s_red);
    if (threadIdx.x == 0) {
      if (tx != 0.f) atomicAdd(sums + S_CNT_X, tx);
      if (ty != 0.f) atomicAdd(sums + S_CNT_Y, ty);
      if (te != 0.f) atomicAdd(sums + S_LAYOUT_ERR, te);
    }
  }
}

// ---- edge-aware smoothness on the patches (train_utils.edge_aware_loss_v2 /
// edge_aware_loss_for_semantic with mask = ray_valid; launched after k_ray_losses, which counted the valid
// edges into sums[S_CNT_X / S_CNT_Y]): block = (patch, channel), thread =
// pixel.  channel 0 = depth (eps 1e-7), channel 1 + c = semantic class c (eps 1e-5, the
// per-channel terms add up).  x_n = x / (mean + eps);  L = sum_edges exp(-mean_c |d rgb|)
// |d x_n|;  dL/dx_k = q_k / (mean + eps) - (sum_i q_i x_i) / (n (mean + eps)^2).
__global__ void __launch_bounds__(1024) k_patch_smooth(nlb_losses_in_t in, float* __restrict__ sums,
                                                      float* __restrict__ g_depth_smo, float* __restrict__ g_sem_smo) {
  extern __shared__ float s_xn[];  // [P*P]
  __shared__ float s_red[32];
  const int P = in.patch_size, n = P * P;
  const int patch = blockIdx.x, ch = blockIdx.y;
  const int t = threadIdx.x, py = t / P, px = t - py * P;
  const int ray = patch * n + t;  // patch rays lead the batch (Z/internal/datasets.py:356-366)
  const bool is_depth = ch == 0;
  const float eps = is_depth ? 1e-7f : 1e-5f;
  const float scale_x = 1.0f / fmaxf(sums[S_CNT_X], 1.0f), scale_y = 1.0f / fmaxf(sums[S_CNT_Y], 1.0f);
  auto ok = [&](int r) { return in.ray_valid == nullptr || __ldg(in.ray_valid + r) != 0.0f; };
  const bool v_c = ok(ray);
  const float x = is_depth ? __ldg(in.depth + ray) : __ldg(in.semantic + (size_t)ray * in.K + (ch - 1));
  const float mean = block_sum(x, s_red) / (float)n;
  const float inv = 1.0f / (mean + eps);
  const float xn = x * inv;
  s_xn[t] = xn;
  __syncthreads();
  const float r0 = __ldg(in.t_rgb + 3 * ray), r1 = __ldg(in.t_rgb + 3 * ray + 1), r2 = __ldg(in.t_rgb + 3 * ray + 2);
  auto edge_w = [&](int other) {  // exp(-mean_c |rgb - rgb_other|)
    const float a = fabsf(r0 - __ldg(in.t_rgb + 3 * other)) + fabsf(r1 - __ldg(in.t_rgb + 3 * other + 1)) +
                    fabsf(r2 - __ldg(in.t_rgb + 3 * other + 2));
    return expf(-(a / 3.0f));
  };
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  float q = 0.f, lx = 0.f, ly = 0.f;
  if (px + 1 < P && v_c && ok(ray + 1)) {  // edge to the right neighbour: counted once, here
    const float w = edge_w(ray + 1), d = xn - s_xn[t + 1];
    lx += w * fabsf(d);
    q += w * sgn(d) * scale_x;
  }
  if (px > 0 && v_c && ok(ray - 1)) {
    const float w = edge_w(ray - 1), d = s_xn[t - 1] - xn;
    q -= w * sgn(d) * scale_x;
  }
  if (py + 1 < P && v_c && ok(ray + P)) {
    const float w = edge_w(ray + P), d = xn - s_xn[t + P];
    ly += w * fabsf(d);
    q += w * sgn(d) * scale_y;
  }
  if (py > 0 && v_c && ok(ray - P)) {
    const float w = edge_w(ray - P), d = s_xn[t - P] - xn;
    q -= w * sgn(d) * scale_y;
  }
  const float sqx = block_sum(q * x, s_red);
  const float g = q * inv - sqx * inv * inv / (float)n;
  if (is_depth) g_depth_smo[ray] = g;
  else g_sem_smo[(size_t)ray * in.K + (ch - 1)] = g;
  const float tx = block_sum(lx, s_red), ty = block_sum(ly, s_red);
  if (t == 0) {
    atomicAdd(sums + (is_depth ? S_DSMO_X : S_SSMO_X), tx);
    atomicAdd(sums + (is_depth ? S_DSMO_Y : S_SSMO_Y), ty);
  }
}

// ---- loss values and the per-term gradient scales (multiplier / denominator)
__global__ void k_finalize_losses(nlb_losses_in_t in, const float* __restrict__ sums, float* __restrict__ losses,
                                  float* __restrict__ scales) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto fin = [](float v) { return isfinite(v); };
  // data: sum / denom if denom > 0 else 0
  {
    const float den = sums[S_DATA_DEN];
    losses[0] = den > 0.f ? sums[S_DATA_NUM] / den : 0.f;
    scales[0] = den > 0.f ? 1.0f / den : 0.f;
  }
  {  // masked means: sum / max(count, 1)
    const float den = fmaxf(sums[S_DEPTH_DEN], 1.0f);
    losses[1] = in.depth_mult * (sums[S_DEPTH_NUM] / den);
    scales[1] = in.depth_mult / den;
  }
  {
    const float den = fmaxf(sums[S_SEM_DEN], 1.0f);
    losses[2] = in.sem_mult * (sums[S_SEM_NUM] / den);
    scales[2] = in.sem_mult / den;
  }
  {
    const float den = fmaxf(sums[S_INT_DEN], 1.0f);
    losses[3] = in.int_mult * (sums[S_INT_NUM] / den);
    scales[3] = in.int_mult / den;
  }
  // smoothness: nan_to_num(0.01 * (sum_x / count_x + sum_y / count_y)) over the valid edges (a direction
  // without valid edges is the reference's NaN -> 0); the per-edge normalisation is already folded into the
  // gradients.  A patch_mask that does not lead the batch poisons both terms.
  {
    const float sx = 1.0f / fmaxf(sums[S_CNT_X], 1.0f), sy = 1.0f / fmaxf(sums[S_CNT_Y], 1.0f);
    const bool empty = sums[S_CNT_X] == 0.f || sums[S_CNT_Y] == 0.f;
    const bool bad = sums[S_LAYOUT_ERR] != 0.f;
    const float v = in.smooth_mult * (sums[S_DSMO_X] * sx + sums[S_DSMO_Y] * sy);
    losses[4] = bad ? NAN : ((fin(v) && !empty) ? v : 0.f);
    scales[4] = (fin(v) && !empty && !bad) ? in.smooth_mult : 0.f;
    const float s = in.smooth_mult * (sums[S_SSMO_X] * sx + sums[S_SSMO_Y] * sy);
    losses[5] = bad ? NAN : ((fin(s) && !empty) ? s : 0.f);
    scales[5] = (fin(s) && !empty && !bad) ? in.smooth_mult : 0.f;
  }
}

}  // namespace nlb

using namespace nlb;

extern "C" size_t nlb_render_losses_workspace_bytes(void) { return kNumSums * sizeof(float); }

extern "C" int nlb_render_losses(const nlb_losses_in_t* in_, float* losses, float* scales, float* g_rgb, float* g_depth,
                                 float* g_sem, float* g_int, float* g_depth_smo, float* g_sem_smo, float* workspace,
                                 void* stream) {
  if (!in_ || !losses || !scales || !workspace) { nlb_set_error("render_losses: null pointer"); return NLB_EINVAL; }
  nlb_losses_in_t in = *in_;
  if (!in.rgb || !in.depth || !in.t_rgb || !in.t_depth || !in.patch_mask || !in.lidar_mask || !g_rgb || !g_depth) {
    nlb_set_error("render_losses: rgb / depth inputs, targets, masks and gradient buffers are required");
    return NLB_EINVAL;
  }
  if (in.semantic && (!in.t_semantic || !g_sem || in.K < 1)) { nlb_set_error("render_losses: semantic needs labels, K and a gradient buffer"); return NLB_EINVAL; }
  if (in.intensity && (!in.t_intensity || !g_int)) { nlb_set_error("render_losses: intensity needs targets and a gradient buffer"); return NLB_EINVAL; }
  if (in.N <= 0) { nlb_set_error("render_losses: empty batch"); return NLB_EINVAL; }
  const int n_patch_rays = in.num_patch * in.patch_size * in.patch_size;
  if (in.num_patch > 0) {
    if (in.patch_size < 2 || in.patch_size > 32 || (in.patch_size * in.patch_size) % 32 != 0) {
      nlb_set_error("render_losses: patch_size %d unsupported (need patch_size^2 a multiple of 32, at most 32)", in.patch_size);
      return NLB_EUNSUPPORTED;
    }
    if (n_patch_rays > in.N) { nlb_set_error("render_losses: %d patch rays but N=%d", n_patch_rays, in.N); return NLB_EINVAL; }
    if (!g_depth_smo || (in.semantic && !g_sem_smo)) { nlb_set_error("render_losses: smoothness gradient buffers required"); return NLB_EINVAL; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, kNumSums * sizeof(float), st) != cudaSuccess) return nlb_check_launch("render_losses memset");
  {
    const size_t qsmem = (size_t)in.N * sizeof(uint32_t);
    const int cached = qsmem <= 160 * 1024;
    static bool attr_set[64] = {false};   // per device
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (dev_ < 0 || dev_ >= 64) dev_ = 0;
    if (!attr_set[dev_]) {
      cudaFuncSetAttribute(k_depth_quantile, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_set[dev_] = true;
    }
    k_depth_quantile<<<1, 1024, cached ? qsmem : 0, st>>>(in, workspace, cached);
  }
  k_ray_losses<<<div_up(in.N, 256), 256, 0, st>>>(in, workspace, g_rgb, g_depth, g_sem, g_int);
  if (in.num_patch > 0) {
    const int pp = in.patch_size * in.patch_size;
    dim3 grid(in.num_patch, 1 + (in.semantic ? in.K : 0));
    k_patch_smooth<<<grid, pp, pp * sizeof(float), st>>>(in, workspace, g_depth_smo, g_sem_smo);
  }
  k_finalize_losses<<<1, 32, 0, st>>>(in, workspace, losses, scales);
  return nlb_check_launch("render_losses");
}
