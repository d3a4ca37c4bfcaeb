// Stage-3 ray-drop front and back end: what happens to a rendered LiDAR sweep after the zipnerf
// path -- `R/` = NeRF_LiDAR/NeRF_Lidar_code/ of the reference:
//   depth_filter               R/src/depth_filter.py:4-31       (per-beam neighbourhood test of the rendered points)
//   LaserScan.do_range_projection   R/src/lidar_utils.py:209-275 (spherical projection to the H x W range image, the
//                                                                 nearest point wins a pixel)
//   the drop selection         R/src/drop_simulation_rays.py:88-140 (U-Net drop probability > threshold, projected
//                                                                 mask, depth filter, sky / road-outlier removal)
// The reference does these on the host with numpy (argsort of all depths + fancy assignment per scan).  Here a
// pixel's winner is ONE 64-bit atomicMin on (depth bits << 32 | point index), the images are filled by a second
// pass over the pixels, and the surviving points are compacted in order on the device.  The U-Net itself
// (R/src/unet/) is not part of this library: its per-pixel logits are an input.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {
namespace raydrop {

// R/src/depth_filter.py: points [H, W, 3] (beam-major sweep); count the +-width neighbours along the beam that lie
// within `radius`; keep if count > threshold, or -- with labels -- on a semantic edge or on a car (class 13).
__global__ void k_depth_filter(const float* __restrict__ points, const float* __restrict__ semantic, int H, int W, int width,
                               float radius, int threshold, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int b = i / W, j = i - b * W;
  const float px = __ldg(points + 3 * i), py = __ldg(points + 3 * i + 1), pz = __ldg(points + 3 * i + 2);
  int count = 0;
  for (int d = -width; d <= width; ++d) {
    if (d == 0) continue;
    int k = (j - d) % W;               // np.roll(points_, d, axis=1)[b, j] = points_[b, (j - d) mod W]
    if (k < 0) k += W;
    const float* q = points + 3 * ((size_t)b * W + k);
    const float dx = __fsub_rn(px, __ldg(q)), dy = __fsub_rn(py, __ldg(q + 1)), dz = __fsub_rn(pz, __ldg(q + 2));
    const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    count += dist < radius;
  }
  bool keep = count > threshold;
  if (semantic) {
    const float s = __ldg(semantic + i);
    const float sl = __ldg(semantic + (size_t)b * W + (j + 1) % W), sr = __ldg(semantic + (size_t)b * W + (j + W - 1) % W);
    keep = keep || sl != s || sr != s || s == 13.f;
  }
  mask[i] = keep ? 1 : 0;
}

// do_range_projection, first pass: pixel of every point (float32 arithmetic as numpy does it on a float32 cloud)
// and the contest for the pixel.  key = depth bits (non-negative floats order like unsigned integers) in the
// high word, the point index in the low one: atomicMin keeps the nearest point, the reference's
// "sort by decreasing depth, assign in order, last one wins".
__global__ void k_range_project(const float* __restrict__ points, int n, int H, int W, float fov, float fov_down,
                                int32_t* __restrict__ proj_x, int32_t* __restrict__ proj_y, float* __restrict__ unproj_range,
                                unsigned long long* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __ldg(points + 3 * i), y = __ldg(points + 3 * i + 1), z = __ldg(points + 3 * i + 2);
  const float depth = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  const float yaw = -atan2f(y, x);
  const float pitch = asinf(__fdiv_rn(z, depth));
  float fx = __fmul_rn(0.5f, __fadd_rn(__fdiv_rn(yaw, 3.14159265358979323846f), 1.0f));
  float fy = __fsub_rn(1.0f, __fdiv_rn(__fadd_rn(pitch, fabsf(fov_down)), fov));
  fx = floorf(__fmul_rn(fx, (float)W));
  fy = floorf(__fmul_rn(fy, (float)H));
  const int ix = (int)fmaxf(0.f, fminf((float)(W - 1), fx));
  const int iy = (int)fmaxf(0.f, fminf((float)(H - 1), fy));
  proj_x[i] = ix;
  proj_y[i] = iy;
  unproj_range[i] = depth;
  if (depth == depth)   // NaN depths (a point at the origin) never win a pixel
    atomicMin(key + (size_t)iy * W + ix, ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned)i);
}

// second pass: one thread per pixel fills the projected images from the winner
__global__ void k_range_fill(const unsigned long long* __restrict__ key, const float* __restrict__ points,
                             const float* __restrict__ semantic, const float* __restrict__ rgb, int H, int W,
                             nlb_range_image_t o) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const unsigned long long k = key[p];
  const bool hit = k != ~0ull;
  const int idx = hit ? (int)(unsigned)(k & 0xffffffffull) : -1;
  if (o.proj_idx) o.proj_idx[p] = idx;
  // `proj_mask = proj_idx > 0` (lidar_utils.py:274): point 0 never counts -- kept
  if (o.proj_mask) o.proj_mask[p] = idx > 0 ? 1.f : 0.f;
  if (o.proj_range) o.proj_range[p] = hit ? __uint_as_float((unsigned)(k >> 32)) : -1.f;
  if (o.proj_xyz) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o.proj_xyz[3 * (size_t)p + c] = hit ? __ldg(points + 3 * (size_t)idx + c) : -1.f;
  }
  if (o.proj_semantic) o.proj_semantic[p] = hit ? (semantic ? __ldg(semantic + idx) : 0.f) : -1.f;
  if (o.proj_rgb) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o.proj_rgb[3 * (size_t)p + c] = (hit && rgb) ? __ldg(rgb + 3 * (size_t)idx + c) : 0.f;
  }
}

// drop selection (drop_simulation_rays.py:104-140, save_near=False): keep[i] of every point
__global__ void k_drop_keep(const float* __restrict__ logits /*[2,H,W]*/, float thre, const float* __restrict__ proj_mask,
                            const int32_t* __restrict__ proj_x, const int32_t* __restrict__ proj_y,
                            const uint8_t* __restrict__ filter_mask, const float* __restrict__ points,
                            const float* __restrict__ labels, int n, int H, int W, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = proj_y[i] * W + proj_x[i];
  // softmax over the two classes, probability of "kept" (class 1)
  const float l0 = __ldg(logits + p), l1 = __ldg(logits + (size_t)H * W + p);
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  const float prob = __fdiv_rn(e1, __fadd_rn(e0, e1));
  bool k = prob > thre && __ldg(proj_mask + p) == 1.f;
  if (filter_mask) k = k && filter_mask[i] == 1;
  const float lab = __ldg(labels + i);
  k = k && lab != 10.f;                                           // sky
  k = k && !(lab == 0.f && __ldg(points + 3 * (size_t)i + 2) < -3.f);   // road outliers
  keep[i] = k ? 1 : 0;
}

// order-preserving compaction of the kept points: per-block counts, scan of the counts, scatter
constexpr int kCompactThreads = 256;
__global__ void __launch_bounds__(kCompactThreads) k_block_counts(const uint8_t* __restrict__ keep, int n, int* __restrict__ counts) {
  const int i = blockIdx.x * kCompactThreads + threadIdx.x;
  const int c = __syncthreads_count(i < n && keep[i]);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) k_scan_counts(int* __restrict__ counts, int nblocks, int* __restrict__ total) {
  __shared__ int s[1024];
  int run = 0;
  for (int base = 0; base < nblocks; base += 1024) {   // one block walks the (few thousand) block counts
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) counts[i] = run + s[threadIdx.x] - v;   // exclusive
    run += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = run;
}
__global__ void __launch_bounds__(kCompactThreads) k_scatter_kept(const uint8_t* __restrict__ keep, const int* __restrict__ offsets,
                                                                  const float* __restrict__ points,
                                                                  const float* __restrict__ labels, int n,
                                                                  float* __restrict__ out_points,
                                                                  float* __restrict__ out_labels) {
  __shared__ int warp_base[kCompactThreads / 32];
  const int i = blockIdx.x * kCompactThreads + threadIdx.x;
  const bool k = i < n && keep[i];
  const unsigned b = __ballot_sync(NLB_FULL_MASK, k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_base[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) { const int c = warp_base[w]; warp_base[w] = run; run += c; }
  }
  __syncthreads();
  if (k) {
    const int dst = offsets[blockIdx.x] + warp_base[warp] + __popc(b & ((1u << lane) - 1u));
    out_points[3 * (size_t)dst] = points[3 * (size_t)i];
    out_points[3 * (size_t)dst + 1] = points[3 * (size_t)i + 1];
    out_points[3 * (size_t)dst + 2] = points[3 * (size_t)i + 2];
    out_labels[dst] = labels[i];
  }
}

}  // namespace raydrop
}  // namespace nlb

using namespace nlb;

extern "C" int nlb_depth_filter(const float* points, const float* semantic, int H, int W, int width, float radius,
                                int threshold, uint8_t* mask, void* stream) {
  if (H == 0 || W == 0) return NLB_OK;
  if (H < 0 || W < 0 || width < 0 || !points || !mask) { nlb_set_error("depth_filter: bad argument"); return NLB_EINVAL; }
  const int n = H * W;
  raydrop::k_depth_filter<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, semantic, H, W, width, radius, threshold, mask);
  return nlb_check_launch("depth_filter");
}

extern "C" size_t nlb_range_projection_workspace_bytes(int H, int W) { return (size_t)H * W * sizeof(unsigned long long); }

extern "C" int nlb_range_projection(const float* points, const float* semantic, const float* rgb, int n, int H, int W,
                                    float fov_up_deg, float fov_down_deg, int32_t* proj_x, int32_t* proj_y,
                                    float* unproj_range, const nlb_range_image_t* image, void* workspace, void* stream) {
  if (H <= 0 || W <= 0 || n < 0 || !image || !workspace || (n > 0 && (!points || !proj_x || !proj_y || !unproj_range))) {
    nlb_set_error("range_projection: bad argument");
    return NLB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(workspace);
  cudaMemsetAsync(key, 0xff, (size_t)H * W * sizeof(unsigned long long), st);
  // laser parameters in radians, rounded to float32 as a float32 numpy expression does with python scalars
  const double up = (double)fov_up_deg / 180.0 * 3.141592653589793, down = (double)fov_down_deg / 180.0 * 3.141592653589793;
  if (n > 0)
    raydrop::k_range_project<<<(n + 255) / 256, 256, 0, st>>>(points, n, H, W, (float)(fabs(down) + fabs(up)), (float)down,
                                                              proj_x, proj_y, unproj_range, key);
  raydrop::k_range_fill<<<(H * W + 255) / 256, 256, 0, st>>>(key, points, semantic, rgb, H, W, *image);
  return nlb_check_launch("range_projection");
}

extern "C" size_t nlb_raydrop_select_workspace_bytes(int n) {
  const int nblocks = (n + raydrop::kCompactThreads - 1) / raydrop::kCompactThreads;
  return (size_t)n + (size_t)(nblocks + 1) * sizeof(int) + 16;
}

extern "C" int nlb_raydrop_select(const float* logits, float mask_thre, const float* proj_mask, const int32_t* proj_x,
                                  const int32_t* proj_y, const uint8_t* filter_mask, const float* points,
                                  const float* labels, int n, int H, int W, float* remain_points, float* remain_labels,
                                  int* remain_count, void* workspace, void* stream) {
  if (!remain_count) { nlb_set_error("raydrop_select: remain_count is required"); return NLB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaMemsetAsync(remain_count, 0, sizeof(int), st); return NLB_OK; }
  if (n < 0 || H <= 0 || W <= 0 || !logits || !proj_mask || !proj_x || !proj_y || !points || !labels || !remain_points ||
      !remain_labels || !workspace) {
    nlb_set_error("raydrop_select: bad argument");
    return NLB_EINVAL;
  }
  const int nblocks = (n + raydrop::kCompactThreads - 1) / raydrop::kCompactThreads;
  uint8_t* keep = reinterpret_cast<uint8_t*>(workspace);
  int* counts = reinterpret_cast<int*>(keep + (((size_t)n + 15) & ~(size_t)15));
  raydrop::k_drop_keep<<<(n + 255) / 256, 256, 0, st>>>(logits, mask_thre, proj_mask, proj_x, proj_y, filter_mask, points, labels,
                                                        n, H, W, keep);
  raydrop::k_block_counts<<<nblocks, raydrop::kCompactThreads, 0, st>>>(keep, n, counts);
  raydrop::k_scan_counts<<<1, 1024, 0, st>>>(counts, nblocks, remain_count);
  raydrop::k_scatter_kept<<<nblocks, raydrop::kCompactThreads, 0, st>>>(keep, counts, points, labels, n, remain_points, remain_labels);
  return nlb_check_launch("raydrop_select");
}
