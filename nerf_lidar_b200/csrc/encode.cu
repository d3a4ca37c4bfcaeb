// Fused sample-point generation + hash-grid encoding for the zipnerf path.
//
// Reference chain (per interval, 7 multisamples), materialised as separate tensors
// there: render.cast_rays (Z/internal/render.py:129-168) -> means[N,S,7,3], stds[N,S,7]
// -> coord.contract_mean_std and /2 (coord.py:51-63, models.py:968-973) ->
// GridEncoder (gridencoder.cu kernel_grid, outputs [L,B,C] + permute copy) ->
// erf re-weighting and mean over the 7 samples (models.py:974-977).
// Here a thread owns one (interval, level) [NeRF level] or one interval with all
// levels [proposal levels] and keeps everything in registers: the only HBM traffic
// is tdist + ray parameters in, table gathers, and features[N*S, L*C] (or the
// proposal density) out.
//
// Launch order is level-major (blockIdx.y = level) for the NeRF table so one 33.5 MB
// level is L2-resident at a time; lanes of a warp are consecutive intervals of the
// same ray, so coarse-level gathers coalesce in L1.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

template <int C>
__device__ __forceinline__ void gather_row(const float* __restrict__ p, float (&v)[C]) {
  if constexpr (C == 1) {
    v[0] = __ldg(p);
  } else if constexpr (C == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (C == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
      v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    }
  }
}

template <int C>
__device__ __forceinline__ void red_add_row(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C == 1) {
    atomicAdd(p, v[0]);
  } else if constexpr (C == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
#pragma unroll
    for (int c = 0; c < C; c += 4)
      atomicAdd(reinterpret_cast<float4*>(p + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
  }
}

// trilinear corner weights in the reference's order: bit d of the corner id selects
// +1 on dimension d, weight product accumulated x, then y, then z.
__device__ __forceinline__ void corner_weights(float fx, float fy, float fz, float (&w)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float wi = 1.f;
    wi *= (i & 1) ? fx : 1.f - fx;
    wi *= (i & 2) ? fy : 1.f - fy;
    wi *= (i & 4) ? fz : 1.f - fz;
    w[i] = wi;
  }
}

// Interpolated feature of one point at one level (same accumulation order as
// kernel_grid, gridencoder.cu:166-191).
template <int C>
__device__ __forceinline__ void lookup(const float* __restrict__ table, const Level3& lv, float x, float y, float z,
                                       float (&out)[C]) {
  uint32_t cx, cy, cz;
  float fx, fy, fz;
  cell_of(x, lv.scale, cx, fx);
  cell_of(y, lv.scale, cy, fy);
  cell_of(z, lv.scale, cz, fz);
  float w[8];
  corner_weights(fx, fy, fz, w);
  float rows[8][C];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t idx = vertex_index3(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
    gather_row<C>(table + ((size_t)lv.offset + idx) * C, rows[i]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = fmaf(w[i], rows[i][c], out[c]);
}

// ----------------------------------------------------------------------------- shared structure
// Every kernel below has one thread per interval.  Phase A computes the interval's 7
// multisample points ONCE into shared memory ([7][128] float4, 14 KB per block);
// phase B walks the levels with compact, non-unrolled (level, sample) loops.  A first
// version recomputed the points per level with everything unrolled: ncu showed it
// issue-bound at 743 instructions per point-level with `no_instruction` (I-cache
// miss) the top stall.  Blocks of one wave sweep the levels together, so the L2
// working set is still about one level of the table at a time.
constexpr int kEncThreads = 128;

__device__ __forceinline__ void stage_points(const nlb_rays_t& rays, int row, float4 (*s_pts)[kEncThreads]) {
  const int ray = row / rays.S, s = row - ray * rays.S;
  const RayGeom rg = load_ray(rays.origins, rays.directions, rays.base_x, rays.base_y, rays.radii, ray);
  const float t0 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s);
  const float t1 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s + 1);
  const bool has_noise = rays.deg_noise != nullptr;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const float noise = has_noise ? __ldg(rays.deg_noise + (size_t)row * 7 + j) : 0.f;
    const SamplePoint p = sample_point(rg, t0, t1, j, noise, has_noise, rays.std_scale);
    // points outside the unit cube contribute zero features (kernel_grid writes zeros):
    // flag them with a negative std
    s_pts[j][threadIdx.x] = make_float4(p.x, p.y, p.z, in_unit_cube(p.x, p.y, p.z) ? p.std : -1.0f);
  }
}

// erf-weighted sum over the 7 samples of the interpolated feature at one level
template <int C>
__device__ __forceinline__ void level_feature(const float* __restrict__ table, const Level3& lv, int gs,
                                              const float4 (*s_pts)[kEncThreads], float (&acc)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const float4 p = s_pts[j][threadIdx.x];
    if (p.w < 0.f) continue;
    const float wj = erf_weight(p.w, gs);
    float f[C];
    lookup<C>(table, lv, p.x, p.y, p.z, f);
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(f[c], wj));
  }
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = acc[c] / 7.0f;
}

// Scatter helper: accumulates corner coefficients of consecutive multisamples that
// fall in the same cell and issues one vector reduction per corner when the cell
// changes (coarse levels: 8 instead of 56 reductions per interval).
template <int C>
struct CellScatter {
  uint32_t cx, cy, cz;
  float w[8];
  bool live;
  __device__ __forceinline__ CellScatter() : cx(0), cy(0), cz(0), live(false) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0.f;
  }
  __device__ __forceinline__ void flush(float* __restrict__ grad_table, const Level3& lv, const float (&g)[C]) {
    if (!live) return;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t idx = vertex_index3(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = g[c] * w[i];
      red_add_row<C>(grad_table + ((size_t)lv.offset + idx) * C, v);
      w[i] = 0.f;
    }
    live = false;
  }
  __device__ __forceinline__ void add(float* __restrict__ grad_table, const Level3& lv, const float (&g)[C],
                                      float x, float y, float z, float coef) {
    uint32_t nx, ny, nz;
    float fx, fy, fz;
    cell_of(x, lv.scale, nx, fx);
    cell_of(y, lv.scale, ny, fy);
    cell_of(z, lv.scale, nz, fz);
    if (live && (nx != cx || ny != cy || nz != cz)) flush(grad_table, lv, g);
    cx = nx; cy = ny; cz = nz;
    live = true;
    float cw[8];
    corner_weights(fx, fy, fz, cw);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = fmaf(coef, cw[i], w[i]);
  }
};

// scatter of one level's feature gradient g (already divided by 7) over the 7 samples
template <int C>
__device__ __forceinline__ void level_scatter(float* __restrict__ grad_table, const Level3& lv, int gs,
                                              const float4 (*s_pts)[kEncThreads], const float (&g)[C]) {
  CellScatter<C> sc;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const float4 p = s_pts[j][threadIdx.x];
    if (p.w < 0.f) continue;
    sc.add(grad_table, lv, g, p.x, p.y, p.z, erf_weight(p.w, gs));
  }
  sc.flush(grad_table, lv, g);
}

// ----------------------------------------------------------------------------- NeRF level
template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_fwd(nlb_rays_t rays, nlb_table_t tab,
                                                            float* __restrict__ features) {
  __shared__ float4 s_pts[7][kEncThreads];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rays.N * rays.S) return;
  stage_points(rays, row, s_pts);
  float* out = features + (size_t)row * (tab.L * C);
#pragma unroll 1
  for (int level = 0; level < tab.L; ++level) {
    const Level3 lv = level3(tab.offsets, level, tab.S, tab.H);
    float acc[C];
    level_feature<C>(tab.embeddings, lv, __ldg(tab.grid_sizes + level), s_pts, acc);
    if constexpr (C == 4) {
      *reinterpret_cast<float4*>(out + level * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else if constexpr (C == 2) {
      *reinterpret_cast<float2*>(out + level * 2) = make_float2(acc[0], acc[1]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) out[level * C + c] = acc[c];
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_bwd(nlb_rays_t rays, nlb_table_t tab,
                                                            const float* __restrict__ grad_features,
                                                            float* __restrict__ grad_table) {
  __shared__ float4 s_pts[7][kEncThreads];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rays.N * rays.S) return;
  stage_points(rays, row, s_pts);
  const float* gin = grad_features + (size_t)row * (tab.L * C);
#pragma unroll 1
  for (int level = 0; level < tab.L; ++level) {
    float g[C];
    gather_row<C>(gin + level * C, g);
    bool any = false;
#pragma unroll
    for (int c = 0; c < C; ++c) { g[c] = g[c] / 7.0f; any |= (g[c] != 0.f); }
    if (!any) continue;
    const Level3 lv = level3(tab.offsets, level, tab.S, tab.H);
    level_scatter<C>(grad_table, lv, __ldg(tab.grid_sizes + level), s_pts, g);
  }
}

// Parity probe: the grid-space sample points (x,y,z in [0,1], contracted std/2) the
// fused kernels generate, [N,S,7,4].
__global__ void __launch_bounds__(128) k_sample_points(nlb_rays_t rays, float* __restrict__ points) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rays.N * rays.S) return;
  const int ray = row / rays.S, s = row - ray * rays.S;
  const RayGeom rg = load_ray(rays.origins, rays.directions, rays.base_x, rays.base_y, rays.radii, ray);
  const float t0 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s);
  const float t1 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s + 1);
  const bool has_noise = rays.deg_noise != nullptr;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const float noise = has_noise ? __ldg(rays.deg_noise + (size_t)row * 7 + j) : 0.f;
    const SamplePoint p = sample_point(rg, t0, t1, j, noise, has_noise, rays.std_scale);
    reinterpret_cast<float4*>(points)[(size_t)row * 7 + j] = make_float4(p.x, p.y, p.z, p.std);
  }
}

// ----------------------------------------------------------------------------- proposal levels
constexpr int kPropHidden = 64;
constexpr int kPropMaxL = 16;

struct PropSmem {
  float W0[kPropHidden * kPropMaxL];
  float b0[kPropHidden];
  float W1[kPropHidden];
  float b1;
};

__device__ __forceinline__ void load_prop_weights(PropSmem& sm, int L, const float* W0, const float* b0, const float* W1,
                                                  const float* b1) {
  for (int i = threadIdx.x; i < kPropHidden * L; i += blockDim.x) sm.W0[i] = __ldg(W0 + i);
  for (int i = threadIdx.x; i < kPropHidden; i += blockDim.x) { sm.b0[i] = __ldg(b0 + i); sm.W1[i] = __ldg(W1 + i); }
  if (threadIdx.x == 0) sm.b1 = __ldg(b1);
}

template <int L>
__global__ void __launch_bounds__(kEncThreads) k_prop_fwd(nlb_rays_t rays, nlb_table_t tab,
                                                          const float* __restrict__ W0, const float* __restrict__ b0,
                                                          const float* __restrict__ W1, const float* __restrict__ b1,
                                                          float* __restrict__ density, float* __restrict__ features) {
  __shared__ PropSmem sm;
  __shared__ float4 s_pts[7][kEncThreads];
  __shared__ float s_f[L][kEncThreads];
  load_prop_weights(sm, L, W0, b0, W1, b1);
  __syncthreads();
  const int tid = threadIdx.x;
  const int row = blockIdx.x * blockDim.x + tid;
  if (row >= rays.N * rays.S) return;
  stage_points(rays, row, s_pts);
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Level3 lv = level3(tab.offsets, l, tab.S, tab.H);
    float acc[1];
    level_feature<1>(tab.embeddings, lv, __ldg(tab.grid_sizes + l), s_pts, acc);
    s_f[l][tid] = acc[0];
    if (features) features[(size_t)row * L + l] = acc[0];
  }
  float f[L];
#pragma unroll
  for (int l = 0; l < L; ++l) f[l] = s_f[l][tid];
  float raw = sm.b1;
#pragma unroll 4
  for (int k = 0; k < kPropHidden; ++k) {
    float h = sm.b0[k];
#pragma unroll
    for (int l = 0; l < L; ++l) h = fmaf(sm.W0[k * L + l], f[l], h);
    raw = fmaf(sm.W1[k], fmaxf(h, 0.f), raw);
  }
  // softplus(raw + density_bias), density_bias = -1 (torch threshold 20)
  const float xin = raw - 1.0f;
  density[row] = xin > 20.f ? xin : log1pf(expf(xin));
}

// Backward of the proposal level, one 128-row tile per block.  Phase 1: per-interval
// MLP backward from the saved features; the tile's weight gradients are a
// [128 x 64]^T [128 x L] product reduced through shared memory and written to a
// per-block partial buffer (no atomics; k_prop_wgrad_reduce sums the partials).
// Phase 2: feature gradients are scattered into the table.
template <int L>
__global__ void __launch_bounds__(kEncThreads) k_prop_bwd(nlb_rays_t rays, nlb_table_t tab,
                                                          const float* __restrict__ W0, const float* __restrict__ b0,
                                                          const float* __restrict__ W1, const float* __restrict__ b1,
                                                          const float* __restrict__ features,
                                                          const float* __restrict__ grad_density,
                                                          float* __restrict__ grad_table,
                                                          float* __restrict__ partial /*[blocks][64*L+129]*/) {
  __shared__ PropSmem sm;
  // phase 1 view: relu output per row (gW1 and the relu mask); phase 2 view: the points
  // and the per-level feature gradients (phase-1 data is dead by then)
  __shared__ __align__(16) float s_raw[kEncThreads * (kPropHidden + 1)];
  float (*s_h)[kPropHidden + 1] = reinterpret_cast<float (*)[kPropHidden + 1]>(s_raw);
  __shared__ float s_f[kEncThreads][L + 1];
  __shared__ float s_graw[kEncThreads];
  static_assert(sizeof(float4) * 7 * kEncThreads + sizeof(float) * L * kEncThreads <= sizeof(s_raw), "smem views");
  load_prop_weights(sm, L, W0, b0, W1, b1);
  __syncthreads();
  const int tid = threadIdx.x;
  const int rows_total = rays.N * rays.S;
  const int row = blockIdx.x * kEncThreads + tid;
  const bool valid = row < rows_total;
  float f[L], gf[L];
#pragma unroll
  for (int l = 0; l < L; ++l) { f[l] = valid ? __ldg(features + (size_t)row * L + l) : 0.f; gf[l] = 0.f; }
  float raw = sm.b1;
#pragma unroll 4
  for (int k = 0; k < kPropHidden; ++k) {
    float h = sm.b0[k];
#pragma unroll
    for (int l = 0; l < L; ++l) h = fmaf(sm.W0[k * L + l], f[l], h);
    s_h[tid][k] = fmaxf(h, 0.f);
    raw = fmaf(sm.W1[k], fmaxf(h, 0.f), raw);
  }
  const float xin = raw - 1.0f;
  const float sig = xin > 20.f ? 1.0f : 1.0f / (1.0f + expf(-xin));  // d softplus
  const float graw = valid ? __ldg(grad_density + row) * sig : 0.f;
  s_graw[tid] = graw;
#pragma unroll 4
  for (int k = 0; k < kPropHidden; ++k) {
    const float gh = (s_h[tid][k] > 0.f) ? graw * sm.W1[k] : 0.f;
#pragma unroll
    for (int l = 0; l < L; ++l) gf[l] = fmaf(gh, sm.W0[k * L + l], gf[l]);
  }
#pragma unroll
  for (int l = 0; l < L; ++l) s_f[tid][l] = f[l];
  __syncthreads();
  constexpr int kEntries = kPropHidden * L + 2 * kPropHidden + 1;
  float* my = partial + (size_t)blockIdx.x * kEntries;
  for (int e = tid; e < kPropHidden * L; e += kEncThreads) {
    const int k = e / L, l = e - k * L;
    const float w1k = sm.W1[k];
    float a = 0.f;
#pragma unroll 4
    for (int r = 0; r < kEncThreads; ++r) a = fmaf((s_h[r][k] > 0.f) ? s_graw[r] * w1k : 0.f, s_f[r][l], a);
    my[e] = a;
  }
  if (tid < kPropHidden) {
    const float w1k = sm.W1[tid];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int r = 0; r < kEncThreads; ++r) {
      a0 += (s_h[r][tid] > 0.f) ? s_graw[r] * w1k : 0.f;
      a1 = fmaf(s_graw[r], s_h[r][tid], a1);
    }
    my[kPropHidden * L + tid] = a0;                // gb0
    my[kPropHidden * L + kPropHidden + tid] = a1;  // gW1
  } else if (tid == kPropHidden) {
    float a = 0.f;
    for (int r = 0; r < kEncThreads; ++r) a += s_graw[r];
    my[kPropHidden * L + 2 * kPropHidden] = a;     // gb1
  }
  __syncthreads();  // every thread is done reading s_h / s_f / s_graw
  if (!valid) return;
  bool any = false;
#pragma unroll
  for (int l = 0; l < L; ++l) { gf[l] = gf[l] / 7.0f; any |= (gf[l] != 0.f); }
  if (!any) return;
  // phase 2: scatter
  float4 (*s_pts)[kEncThreads] = reinterpret_cast<float4 (*)[kEncThreads]>(s_raw);
  float (*s_gf)[kEncThreads] = reinterpret_cast<float (*)[kEncThreads]>(s_raw + 4 * 7 * kEncThreads);
#pragma unroll
  for (int l = 0; l < L; ++l) s_gf[l][tid] = gf[l];
  stage_points(rays, row, s_pts);
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Level3 lv = level3(tab.offsets, l, tab.S, tab.H);
    float g1[1] = {s_gf[l][tid]};
    if (g1[0] == 0.f) continue;
    level_scatter<1>(grad_table, lv, __ldg(tab.grid_sizes + l), s_pts, g1);
  }
}

// sums the per-block partial weight gradients: one thread per entry, coalesced reads
__global__ void __launch_bounds__(256) k_prop_wgrad_reduce(const float* __restrict__ partial, int blocks, int entries,
                                                           int L, float* __restrict__ gW0, float* __restrict__ gb0,
                                                           float* __restrict__ gW1, float* __restrict__ gb1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  float a = 0.f;
  for (int b = 0; b < blocks; ++b) a += __ldg(partial + (size_t)b * entries + e);
  const int n0 = kPropHidden * L;
  if (e < n0) gW0[e] += a;
  else if (e < n0 + kPropHidden) gb0[e - n0] += a;
  else if (e < n0 + 2 * kPropHidden) gW1[e - n0 - kPropHidden] += a;
  else gb1[0] += a;
}

}  // namespace nlb

using namespace nlb;

static int check_rays_table(const nlb_rays_t* r, const nlb_table_t* t, const char* who) {
  if (!r || !t) { nlb_set_error("%s: null descriptor", who); return NLB_EINVAL; }
  if (r->N < 0 || r->S < 1) { nlb_set_error("%s: bad N/S", who); return NLB_EINVAL; }
  if (!r->tdist || !r->origins || !r->directions || !r->radii || !r->base_x || !r->base_y || !t->embeddings ||
      !t->offsets || !t->grid_sizes) {
    nlb_set_error("%s: null pointer", who);
    return NLB_EINVAL;
  }
  if ((int64_t)r->N * r->S > 0x7fffffffLL / 16) { nlb_set_error("%s: N*S too large for one launch; chunk the rays", who); return NLB_EINVAL; }
  return NLB_OK;
}

extern "C" int nlb_sample_points(const nlb_rays_t* rays, float* points, void* stream) {
  if (!rays || !points) { nlb_set_error("sample_points: null pointer"); return NLB_EINVAL; }
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  k_sample_points<<<div_up(rows, 128), 128, 0, (cudaStream_t)stream>>>(*rays, points);
  return nlb_check_launch("sample_points");
}

extern "C" int nlb_encode_forward(const nlb_rays_t* rays, const nlb_table_t* table, float* features, void* stream) {
  if (int e = check_rays_table(rays, table, "encode_forward")) return e;
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  dim3 grid(div_up(rows, kEncThreads));
  cudaStream_t st = (cudaStream_t)stream;
  switch (table->C) {
    case 1: k_encode_fwd<1><<<grid, 128, 0, st>>>(*rays, *table, features); break;
    case 2: k_encode_fwd<2><<<grid, 128, 0, st>>>(*rays, *table, features); break;
    case 4: k_encode_fwd<4><<<grid, 128, 0, st>>>(*rays, *table, features); break;
    case 8: k_encode_fwd<8><<<grid, 128, 0, st>>>(*rays, *table, features); break;
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
  return nlb_check_launch("encode_forward");
}

extern "C" int nlb_encode_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* grad_features,
                                   float* grad_embeddings, void* stream) {
  if (int e = check_rays_table(rays, table, "encode_backward")) return e;
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  dim3 grid(div_up(rows, kEncThreads));
  cudaStream_t st = (cudaStream_t)stream;
  switch (table->C) {
    case 1: k_encode_bwd<1><<<grid, 128, 0, st>>>(*rays, *table, grad_features, grad_embeddings); break;
    case 2: k_encode_bwd<2><<<grid, 128, 0, st>>>(*rays, *table, grad_features, grad_embeddings); break;
    case 4: k_encode_bwd<4><<<grid, 128, 0, st>>>(*rays, *table, grad_features, grad_embeddings); break;
    case 8: k_encode_bwd<8><<<grid, 128, 0, st>>>(*rays, *table, grad_features, grad_embeddings); break;
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
  return nlb_check_launch("encode_backward");
}

#define NLB_PROP_DISPATCH(L, ...)                         \
  switch (L) {                                            \
    case 4: { constexpr int L_ = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int L_ = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int L_ = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int L_ = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int L_ = 8; __VA_ARGS__; } break; \
    case 9: { constexpr int L_ = 9; __VA_ARGS__; } break; \
    case 10: { constexpr int L_ = 10; __VA_ARGS__; } break; \
    default: nlb_set_error("prop: num_levels %d not built (4..10)", L); return NLB_EUNSUPPORTED; \
  }

extern "C" int nlb_prop_forward(const nlb_rays_t* rays, const nlb_table_t* table, const float* W0, const float* b0,
                                const float* W1, const float* b1, float* density, float* features, void* stream) {
  if (int e = check_rays_table(rays, table, "prop_forward")) return e;
  if (table->C != 1) { nlb_set_error("prop_forward: PropMLP tables have level_dim 1"); return NLB_EINVAL; }
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NLB_PROP_DISPATCH(table->L, (k_prop_fwd<L_><<<div_up(rows, 128), 128, 0, st>>>(*rays, *table, W0, b0, W1, b1, density, features)));
  return nlb_check_launch("prop_forward");
}

extern "C" size_t nlb_prop_backward_workspace_bytes(int N, int S, int L) {
  const size_t blocks = ((size_t)N * S + kEncThreads - 1) / kEncThreads;
  return blocks * (size_t)(kPropHidden * L + 2 * kPropHidden + 1) * sizeof(float);
}

extern "C" int nlb_prop_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* W0, const float* b0,
                                 const float* W1, const float* b1, const float* features, const float* grad_density,
                                 float* grad_embeddings, float* gW0, float* gb0, float* gW1, float* gb1,
                                 float* workspace, void* stream) {
  if (int e = check_rays_table(rays, table, "prop_backward")) return e;
  if (table->C != 1) { nlb_set_error("prop_backward: PropMLP tables have level_dim 1"); return NLB_EINVAL; }
  if (!features || !grad_density) { nlb_set_error("prop_backward: features saved by the forward are required"); return NLB_EINVAL; }
  if (!workspace) { nlb_set_error("prop_backward: workspace of nlb_prop_backward_workspace_bytes() is required"); return NLB_EINVAL; }
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)div_up(rows, kEncThreads);
  NLB_PROP_DISPATCH(table->L, (k_prop_bwd<L_><<<blocks, kEncThreads, 0, st>>>(
      *rays, *table, W0, b0, W1, b1, features, grad_density, grad_embeddings, workspace)));
  if (int e = nlb_check_launch("prop_backward")) return e;
  const int entries = kPropHidden * table->L + 2 * kPropHidden + 1;
  k_prop_wgrad_reduce<<<div_up(entries, 256), 256, 0, st>>>(workspace, blocks, entries, table->L, gW0, gb0, gW1, gb1);
  return nlb_check_launch("prop_wgrad_reduce");
}
