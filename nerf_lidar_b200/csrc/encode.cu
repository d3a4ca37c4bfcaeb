// Fused sample-point generation + hash-grid encoding for the zipnerf path.
//
// Reference chain (per interval, 7 multisamples), materialised as separate tensors
// there: render.cast_rays (Z/internal/render.py:129-168) -> means[N,S,7,3], stds[N,S,7]
// -> coord.contract_mean_std and /2 (coord.py:51-63, models.py:968-973) ->
// GridEncoder (gridencoder.cu kernel_grid, outputs [L,B,C] + permute copy) ->
// erf re-weighting and mean over the 7 samples (models.py:974-977).
// Here one interval (7 sample points) belongs to a LANE PAIR in the forward kernels and in the scatter of
// hashed level groups (lane h takes the four corners at x + h of every cell: the pair's two addresses of one
// load / reduction instruction share a 128-byte line, see k_prop_fwd_pair), and to one thread in the scatter
// of dense levels (warp aggregation over runs of equal cells along the ray).  The sample points are
// generated once into shared memory (or, in the backward of a training step, read back from the forward's
// cache) and the levels are walked with rolled loops; the only HBM traffic is tdist + ray parameters in,
// table gathers / reductions, and features[N*S, L*C] (or the proposal density) out.  Lanes of a warp are
// consecutive intervals of the same ray, so coarse-level gathers coalesce in L1 and coarse-level reductions
// aggregate across the warp.  Forward: one launch per L2-sized level group.  Backward (scatter): persistent
// blocks, level groups sized to L2 -- see k_encode_bwd.  The one-lane-per-interval forward kernels
// (k_encode_fwd, k_prop_fwd) are kept for A/B timing (NLB_ENC_FWD_LEGACY / NLB_PROP_FWD_LEGACY).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

constexpr int kMaxLevelsEnc = 16;

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

// Per-level constants, computed once per block into shared memory.
struct LevelCache {
  Level3 lv[kMaxLevelsEnc];
  float inv_gs[kMaxLevelsEnc];  // 1 / grid_size
};

// erf re-weighting of one sample at one level (models.py:976): erf(1 / max(sqrt(8 std^2 gs^2), 1e-10)).  The
// staged points carry a = 1 / sqrt(8 std^2) (one MUFU.RSQ per point instead of an IEEE sqrt and divide per
// point-level: 30 of ~180 instructions of a lookup), so the argument is a * (1 / gs); std = 0 gives inf -> 1
// like the reference's clamp.  Agreement with the reference expression: ~2 ulp of the argument.
__device__ __forceinline__ float staged_a(float sd) { return rsqrtf(8.0f * sd * sd); }
// For x >= 4, erf(x) = 1 - 1.5e-8 or closer: 1.0f is the correctly rounded value (and what torch.erf returns), so the
// ~40-instruction erff is branched around.  On the driving-scene rays (pixel footprint 4.6e-4 x t) every sample of
// the levels up to resolution 512 and 87 % of those at 1024 saturate (whole warps take the short path): the erf
// loop was 160 of ~560 warp instructions per level of the forward kernels, 280 per level of the one-lane scatter.
__device__ __forceinline__ float erf_weight_a(float a, float inv_gs) {
  const float x = a * inv_gs;
  if (x >= 4.0f) return 1.0f;
  return erff(x);
}

__device__ __forceinline__ void fill_level_cache(LevelCache& lc, const nlb_table_t& tab) {
  if (threadIdx.x < tab.L) {
    lc.lv[threadIdx.x] = level3(tab.offsets, threadIdx.x, tab.S, tab.H);
    lc.inv_gs[threadIdx.x] = 1.0f / (float)__ldg(tab.grid_sizes + threadIdx.x);
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
  // (8-byte paired gathers for even cx on the hashed C=1 levels were measured: 5 % SLOWER in the
  // forward -- the extra divergent path costs more than the saved L1 wavefronts -- but 12 % faster as
  // paired vector reductions in the scatter, see level_scatter)
  {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t vx = cx + (i & 1), vy = cy + ((i >> 1) & 1), vz = cz + ((i >> 2) & 1);
      const uint32_t idx = (C == 1) ? vertex_index3(lv, vx, vy, vz) : vertex_index3_branchy(lv, vx, vy, vz);
      gather_row<C>(table + ((size_t)lv.offset + idx) * C, rows[i]);
    }
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
  const size_t rows = (size_t)rays.N * rays.S;
  float4* cache = reinterpret_cast<float4*>(rays.points_cache);
  if (rays.points_mode == 2) {  // backward: the forward's points, [7][rows] so a warp reads 512 contiguous bytes
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      float4 v = __ldg(cache + j * rows + row);
      if (v.w >= 0.f) v.w = staged_a(v.w);  // the cache keeps std (nlb200.h), the staged copy a
      s_pts[j][threadIdx.x] = v;
    }
    return;
  }
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
    const bool inside = in_unit_cube(p.x, p.y, p.z);
    s_pts[j][threadIdx.x] = make_float4(p.x, p.y, p.z, inside ? staged_a(p.std) : -1.0f);
    // evict-first: must not displace the table in L2
    if (rays.points_mode == 1) __stcs(cache + j * rows + row, make_float4(p.x, p.y, p.z, inside ? p.std : -1.0f));
  }
}

// erf-weighted sum over the 7 samples of the interpolated feature at one level
template <int C>
__device__ __forceinline__ void level_feature(const float* __restrict__ table, const Level3& lv, float inv_gs,
                                              const float4 (*s_pts)[kEncThreads], float (&acc)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const float4 p = s_pts[j][threadIdx.x];
    if (p.w < 0.f) continue;
    const float wj = erf_weight_a(p.w, inv_gs);
    float f[C];
    lookup<C>(table, lv, p.x, p.y, p.z, f);
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(f[c], wj));
  }
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = acc[c] / 7.0f;
}

// ----------------------------------------------------------------------------- scatter (backward)
// Adds v[C] into row `idx` of level `lv`: shared-memory accumulator (kStaged) or one
// vector reduction to the global gradient table.
template <int C, bool kStaged>
__device__ __forceinline__ void add_row(float* __restrict__ acc, const Level3& lv, uint32_t idx, const float (&v)[C]) {
  float* dst = acc + ((size_t)lv.offset + idx) * C;
  if constexpr (kStaged) {
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(dst + c, v[c]);
  } else {
    red_add_row<C>(dst, v);
  }
}

// Scatter of one level's feature gradient g (already divided by 7) over the 7 samples of
// one interval.  Consecutive samples that fall in the same cell are merged before the 8
// corner reductions are issued (coarse levels: 8 instead of 56 per interval); the loop
// runs one sentinel iteration so there is a single flush site.
//   kStaged: the level lives in the block's shared-memory accumulator `acc` (indexed like
//            the table, offset included); otherwise reductions go to the global table.
//   kWarpAgg (dense levels; the whole warp must call, `has_g` masks idle lanes): lanes are
//            consecutive intervals of one ray, so neighbouring lanes end in the same cell;
//            each lane's last cell group is summed over runs of equal cells with a
//            segmented shuffle reduction and only the run head issues the 8 reductions.
//            Every ray of a scene starts in the same few coarse cells -- without this,
//            L2 serialises millions of same-address atomics (measured: 0.4 ms per level).
template <int C, bool kStaged, bool kWarpAgg>
__device__ __forceinline__ void level_scatter(float* __restrict__ acc, const Level3& lv, float inv_gs,
                                              const float4 (*s_pts)[kEncThreads], const float (&g)[C], bool has_g) {
  uint32_t cx = 0, cy = 0, cz = 0;
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = 0.f;
  bool live = false;
  constexpr int kLast = kWarpAgg ? 6 : 7;  // the sentinel iteration flushes the last group
#pragma unroll 1
  for (int j = 0; j <= kLast; ++j) {
    float4 p = make_float4(0.f, 0.f, 0.f, -1.f);
    if (j < 7) p = s_pts[j][threadIdx.x];
    const bool valid = has_g && p.w >= 0.f;
    uint32_t nx = 0, ny = 0, nz = 0;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (valid) {
      cell_of(p.x, lv.scale, nx, fx);
      cell_of(p.y, lv.scale, ny, fy);
      cell_of(p.z, lv.scale, nz, fz);
    }
    if (live && (j == 7 || (valid && (nx != cx || ny != cy || nz != cz)))) {
      if (C == 1 && !kStaged && !kWarpAgg && (cx & 1u) == 0) {
        // hashed level (kWarpAgg is the dense path), even cx: corners x / x+1 are h / h ^ 1 -> one 8-byte
        // vector reduction per (y, z) instead of two scalar ones
        const uint32_t hy0 = cy * 2654435761u, hy1 = (cy + 1) * 2654435761u;
        const uint32_t hz0 = cz * 805459861u, hz1 = (cz + 1) * 805459861u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = ((k & 1) ? hy1 : hy0) ^ ((k >> 1) ? hz1 : hz0);
          const uint32_t i0 = (cx ^ a) & lv.mask;
          const float v0 = g[0] * w[2 * k], v1 = g[0] * w[2 * k + 1];
          atomicAdd(reinterpret_cast<float2*>(acc + lv.offset + (i0 & ~1u)),
                    (i0 & 1u) ? make_float2(v1, v0) : make_float2(v0, v1));
          w[2 * k] = 0.f;
          w[2 * k + 1] = 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t idx = vertex_index3(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
          float v[C];
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = g[c] * w[i];
          add_row<C, kStaged>(acc, lv, idx, v);
          w[i] = 0.f;
        }
      }
      live = false;
    }
    if (valid) {
      cx = nx; cy = ny; cz = nz;
      live = true;
      const float coef = erf_weight_a(p.w, inv_gs);
      float cw[8];
      corner_weights(fx, fy, fz, cw);
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = fmaf(coef, cw[i], w[i]);
    }
  }
  if constexpr (kWarpAgg) {
    const int lane = threadIdx.x & 31;
    const uint32_t key = live ? (cx + cy * lv.s1 + cz * lv.s2) : 0xffffffffu;  // dense level: unique per cell
    const uint32_t prev = __shfl_up_sync(NLB_FULL_MASK, key, 1);
    const bool head = lane == 0 || key != prev || !live;
    const uint32_t heads = __ballot_sync(NLB_FULL_MASK, head);
    const uint32_t later = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
    const int run_end = later ? __ffs(later) - 1 : 32;  // exclusive end of this lane's run
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = g[c] * w[i];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = __shfl_down_sync(NLB_FULL_MASK, v[c], o);
          if (lane + o < run_end) v[c] += t;
        }
      }
      if (head && live) {
        const uint32_t idx = vertex_index3(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
        add_row<C, kStaged>(acc, lv, idx, v);
      }
    }
  }
}

// ----------------------------------------------------------------------------- NeRF level
template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_fwd(nlb_rays_t rays, nlb_table_t tab,
                                                            float* __restrict__ features, int level_begin,
                                                            int level_end) {
  __shared__ float4 s_pts[7][kEncThreads];
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rays.N * rays.S) return;
  stage_points(rays, row, s_pts);
  float* out = features + (size_t)row * (tab.L * C);
#pragma unroll 1
  for (int level = level_begin; level < level_end; ++level) {
    const Level3 lv = lc.lv[level];
    float acc[C];
    level_feature<C>(tab.embeddings, lv, lc.inv_gs[level], s_pts, acc);
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

// Backward of the fused encode for the levels [level_begin, level_end): persistent blocks
// (a multiple of the SM count) walk the 128-interval tiles.  Dense (coarse) levels use the
// warp-aggregated scatter; the coarsest ones that fit the shared-memory budget
// (`staged_rows` table rows, chosen by the host) are additionally accumulated in a
// per-block shared-memory copy of those rows and flushed once per block; the next dense
// levels (up to `priv_rows` rows) go to one of `priv_copies` private copies in global
// memory that k_priv_reduce sums afterwards -- with a single copy the few cache lines
// around the scene centre serialise in L2 (measured: 0.13 ms for level 1 of a C=1 table
// against 0.02 ms for the 8x larger level 2); hashed levels go to L2 as one vector
// reduction per corner.  The host launches the fine levels in groups
// whose gradient rows fit L2 together (level-major order), so a 33.5 MB level stays
// resident while it is being updated.
template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_bwd(nlb_rays_t rays, nlb_table_t tab,
                                                            const float* __restrict__ grad_features,
                                                            float* __restrict__ grad_table, int staged_rows,
                                                            int num_tiles, int level_begin, int level_end,
                                                            float* __restrict__ priv, int priv_rows, int priv_copies) {
  extern __shared__ __align__(16) float s_acc[];  // [staged_rows * C]
  __shared__ float4 s_pts[7][kEncThreads];
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  for (int i = threadIdx.x; i < staged_rows * C; i += kEncThreads) s_acc[i] = 0.f;
  __syncthreads();
  const int rows_total = rays.N * rays.S;
  const int LC = tab.L * C;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row = tile * kEncThreads + threadIdx.x;
    const bool row_ok = row < rows_total;
    const float* gin = grad_features + (size_t)row * LC;
    // the next level's gradient row is fetched while this level is scattered (the first one under the staging):
    // the load was the one exposed global round trip per (tile, level) of a kernel that runs at 20 warps per SM
    float g_next[C];
#pragma unroll
    for (int c = 0; c < C; ++c) g_next[c] = 0.f;
    if (row_ok) gather_row<C>(gin + level_begin * C, g_next);
    if (row_ok) stage_points(rays, row, s_pts);  // column threadIdx.x is private to this thread: no barrier
#pragma unroll 1
    for (int level = level_begin; level < level_end; ++level) {
      float g[C];
      bool any = false;
#pragma unroll
      for (int c = 0; c < C; ++c) { g[c] = g_next[c] / 7.0f; any |= (g[c] != 0.f); }
      if (row_ok && level + 1 < level_end) gather_row<C>(gin + (level + 1) * C, g_next);
      const Level3 lv = lc.lv[level];
      if (lv.dense) {  // uniform per level: the whole warp takes the same branch
        if ((int)(lv.offset + lv.hashmap_size) <= staged_rows)
          level_scatter<C, true, true>(s_acc, lv, lc.inv_gs[level], s_pts, g, any);
        else if ((int)(lv.offset + lv.hashmap_size) <= priv_rows)  // privatised copy of the coarse rows
          level_scatter<C, false, true>(priv + (size_t)(blockIdx.x % priv_copies) * priv_rows * C, lv, lc.inv_gs[level],
                                        s_pts, g, any);
        else
          level_scatter<C, false, true>(grad_table, lv, lc.inv_gs[level], s_pts, g, any);
      } else if (any) {
        level_scatter<C, false, false>(grad_table, lv, lc.inv_gs[level], s_pts, g, true);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < staged_rows * C; i += kEncThreads) {
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(grad_table + i, v);
  }
}

// Sums the privatised copies of the coarse table rows into the gradient table.
__global__ void __launch_bounds__(256) k_priv_reduce(const float* __restrict__ priv, int copies, int n /*floats per copy*/,
                                                    float* __restrict__ grad_table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
#pragma unroll 8   // (the loads of the copies in flight together; same summation order)
  for (int c = 0; c < copies; ++c) a += __ldg(priv + (size_t)c * n + i);
  if (a != 0.f) grad_table[i] += a;
}

// ----------------------------------------------------------------------------- gradients w.r.t. the ray geometry
// Pose-refinement window (Z/train.py:200-221): origins / directions / base_x / base_y carry gradients, which the
// reference obtains through dy_dx + kernel_input_backward (gridencoder.cu:201-244,343-369), the erf weights, the
// contraction (coord.py:51-63 under autograd) and cast_rays (render.py:129-168).  One thread per interval recomputes
// its 7 samples; per sample and level it gathers the 8 corner rows once and forms both d feature / d x (the
// trilinear derivative, dy_dx of the reference) and the feature itself (the erf weight depends on the contracted
// std, which depends on |m|); the chain through the contraction Jacobian and the linear cast_rays map follows in
// registers, then one warp reduction and 12 atomic adds per warp.  tdist carries no gradient (models.py:368-369).
template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_input_bwd(nlb_rays_t rays, nlb_table_t tab,
                                                                  const float* __restrict__ grad_features,
                                                                  float* __restrict__ g_origins,
                                                                  float* __restrict__ g_directions,
                                                                  float* __restrict__ g_base_x,
                                                                  float* __restrict__ g_base_y, int level_begin,
                                                                  int level_end) {
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  __syncthreads();
  // one thread per (interval, sample): 7x the threads of the forward's layout, 10 x 8 dependent gathers each
  // (thread-per-interval with the sample loop inside was latency-bound: 0.72 ms against 0.39 ms for the forward)
  const long items = (long)rays.N * rays.S * 7;
  const long item_raw = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = item_raw < items;
  const long item = active ? item_raw : items - 1;
  const int row = (int)(item / 7), j = (int)(item - (long)row * 7);
  const int ray = row / rays.S, s = row - ray * rays.S;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  if (active) {
    const RayGeom rg = load_ray(rays.origins, rays.directions, rays.base_x, rays.base_y, rays.radii, ray);
    const float t0 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s);
    const float t1 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s + 1);
    const bool has_noise = rays.deg_noise != nullptr;
    const float* __restrict__ G = grad_features + (size_t)row * (tab.L * C);
    do {
      const float noise = has_noise ? __ldg(rays.deg_noise + (size_t)row * 7 + j) : 0.f;
      SampleGeom geo;
      const SamplePoint p = sample_point(rg, t0, t1, j, noise, has_noise, rays.std_scale, &geo);
      if (!in_unit_cube(p.x, p.y, p.z)) break;  // zero features, zero gradient (gridencoder.cu:110-135)
      const float a = staged_a(p.std);
      float gx = 0.f, gy = 0.f, gz = 0.f, gsd = 0.f;  // dL/d(x01), dL/d(std/2)
#pragma unroll 1
      for (int level = level_begin; level < level_end; ++level) {
        const Level3 lv = lc.lv[level];
        const float u = a * lc.inv_gs[level];
        const float w = u >= 4.0f ? 1.0f : erff(u);   // (see erf_weight_a: exact, and most levels saturate)
        // d erf(u)/d std with u = 1/(sqrt(8) std gs): -(2/sqrt(pi)) exp(-u^2) u / std; saturated weights are flat
        const float dw = (u < 12.f) ? -1.1283791671f * __expf(-u * u) * u / p.std : 0.f;
        uint32_t cx, cy, cz;
        float fx, fy, fz;
        cell_of(p.x, lv.scale, cx, fx);
        cell_of(p.y, lv.scale, cy, fy);
        cell_of(p.z, lv.scale, cz, fz);
        float rowsv[8][C];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t vx = cx + (i & 1), vy = cy + ((i >> 1) & 1), vz = cz + ((i >> 2) & 1);
          const uint32_t idx = (C == 1) ? vertex_index3(lv, vx, vy, vz) : vertex_index3_branchy(lv, vx, vy, vz);
          gather_row<C>(tab.embeddings + ((size_t)lv.offset + idx) * C, rowsv[i]);
        }
        float gf = 0.f, dfx = 0.f, dfy = 0.f, dfz = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float g = __ldg(G + level * C + c);
          // bilinear in (y,z) of the x-pairs, then the x interpolation and its derivative
          const float x00 = rowsv[1][c] - rowsv[0][c], x10 = rowsv[3][c] - rowsv[2][c];
          const float x01 = rowsv[5][c] - rowsv[4][c], x11 = rowsv[7][c] - rowsv[6][c];
          const float v00 = fmaf(fx, x00, rowsv[0][c]), v10 = fmaf(fx, x10, rowsv[2][c]);
          const float v01 = fmaf(fx, x01, rowsv[4][c]), v11 = fmaf(fx, x11, rowsv[6][c]);
          const float y0 = v10 - v00, y1 = v11 - v01;
          const float e0 = fmaf(fy, y0, v00), e1 = fmaf(fy, y1, v01);
          const float f = fmaf(fz, e1 - e0, e0);
          const float dx0 = fmaf(fy, x10 - x00, x00), dx1 = fmaf(fy, x11 - x01, x01);
          gf = fmaf(g, f, gf);
          dfx = fmaf(g, fmaf(fz, dx1 - dx0, dx0), dfx);
          dfy = fmaf(g, fmaf(fz, y1 - y0, y0), dfy);
          dfz = fmaf(g, e1 - e0, dfz);
        }
        const float ws = w * lv.scale;  // pos = x * scale + 0.5
        gx = fmaf(ws, dfx, gx);
        gy = fmaf(ws, dfy, gy);
        gz = fmaf(ws, dfz, gz);
        gsd = fmaf(gf, dw, gsd);
      }
      // mean over the 7 samples; x01 = z/4 + 1/2, std/2
      const float inv7 = 1.0f / 7.0f;
      float zx = gx * (0.25f * inv7), zy = gy * (0.25f * inv7), zz = gz * (0.25f * inv7);
      const float gsdc = gsd * (0.5f * inv7);
      if (!(geo.m2 <= 1.0f)) {
        // z = k(q) m with q = |m|^2, k = 2 q^-1/2 - q^-1; std' = std (k^2/q)^(1/3)
        const float q = geo.m2, k = geo.k;
        const float rq = 1.0f / q;
        const float dk = rq * rq - rq * rsqrtf(q);
        const float det = k * k * rq;
        const float ddet = (2.0f * k * dk - det) * rq;
        const float cb = cbrtf(det);
        const float dsd = geo.sd * ddet / (3.0f * cb * cb);
        const float dot = zx * geo.mx + zy * geo.my + zz * geo.mz;
        const float coef = 2.0f * fmaf(dot, dk, gsdc * dsd);
        zx = fmaf(coef, geo.mx, k * zx);
        zy = fmaf(coef, geo.my, k * zy);
        zz = fmaf(coef, geo.mz, k * zz);
      }
      acc[0] += zx; acc[1] += zy; acc[2] += zz;
      acc[3] = fmaf(geo.t, zx, acc[3]); acc[4] = fmaf(geo.t, zy, acc[4]); acc[5] = fmaf(geo.t, zz, acc[5]);
      acc[6] = fmaf(geo.lx, zx, acc[6]); acc[7] = fmaf(geo.lx, zy, acc[7]); acc[8] = fmaf(geo.lx, zz, acc[8]);
      acc[9] = fmaf(geo.ly, zx, acc[9]); acc[10] = fmaf(geo.ly, zy, acc[10]); acc[11] = fmaf(geo.ly, zz, acc[11]);
    } while (false);
  }
  // a warp's 32 samples belong to one ray when S % 32 == 0: one reduction, 12 atomics
  const int ray0 = __shfl_sync(NLB_FULL_MASK, ray, 0);
  const bool uniform = __all_sync(NLB_FULL_MASK, ray == ray0);
  if (uniform) {
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = warp_sum(acc[i]);
    if ((threadIdx.x & 31) != 0) return;
  } else if (!active) {
    return;
  }
  float* dst[4] = {g_origins, g_directions, g_base_x, g_base_y};
#pragma unroll
  for (int i = 0; i < 12; ++i)
    if (acc[i] != 0.f) atomicAdd(dst[i / 3] + 3 * (size_t)ray + (i % 3), acc[i]);
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
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  load_prop_weights(sm, L, W0, b0, W1, b1);
  __syncthreads();
  const int tid = threadIdx.x;
  const int row = blockIdx.x * blockDim.x + tid;
  if (row >= rays.N * rays.S) return;
  stage_points(rays, row, s_pts);
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Level3 lv = lc.lv[l];
    float acc[1];
    level_feature<1>(tab.embeddings, lv, lc.inv_gs[l], s_pts, acc);
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

// Proposal forward with TWO lanes per interval.  ncu + a cycle model of k_prop_fwd: on the hashed C=1 levels
// every lane of a gather instruction touches its own 128-byte line and L1 retires about one line per clock
// per SM (wavefronts = 32 per LDG; predicted 0.43 / 0.68 ms for the 6- / 8-level tables, measured 0.42 /
// 0.78).  The x and x+1 corners of a cell are h and h ^ (x ^ (x+1)) -- the same line for 31 of 32 cells --
// so lane pair (2i, 2i+1) of interval i takes the four (y, z) corners at x and at x+1 respectively: the pair's
// two addresses of one LDG coalesce into one wavefront (16 per instruction instead of 32), on dense levels
// they are adjacent words.  Each lane also does half of the staging, of the erf weights of a level and of the
// hidden units of the density MLP.  Summation order: per lane sum_j erf_j * (sum over its 4 corners), then the
// two lanes are added (the reference sums the 8 corners of a sample first; agreement is to fp32 rounding,
// the indices are the same).
constexpr int kPairIv = kEncThreads / 2;

// (floor / frac through the ALU -- pos + 2^23 trick, exact for pos < 2^22 -- instead of F2I + FRND per axis
// was measured: 0.282 / 0.366 ms against 0.262 / 0.350, the conversion pipe is not the limit)
__device__ __forceinline__ void stage_points_pair(const nlb_rays_t& rays, int row, int h, int iv,
                                                  float4 (*s_pts)[kPairIv]) {
  const size_t rows = (size_t)rays.N * rays.S;
  float4* cache = reinterpret_cast<float4*>(rays.points_cache);
  if (rays.points_mode == 2) {
    for (int j = h; j < 7; j += 2) {
      float4 v = __ldg(cache + j * rows + row);
      if (v.w >= 0.f) v.w = staged_a(v.w);
      s_pts[j][iv] = v;
    }
    return;
  }
  const int ray = row / rays.S, s = row - ray * rays.S;
  const RayGeom rg = load_ray(rays.origins, rays.directions, rays.base_x, rays.base_y, rays.radii, ray);
  const float t0 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s);
  const float t1 = __ldg(rays.tdist + (size_t)ray * (rays.S + 1) + s + 1);
  const bool has_noise = rays.deg_noise != nullptr;
#pragma unroll 1
  for (int j = h; j < 7; j += 2) {
    const float noise = has_noise ? __ldg(rays.deg_noise + (size_t)row * 7 + j) : 0.f;
    const SamplePoint p = sample_point(rg, t0, t1, j, noise, has_noise, rays.std_scale);
    const bool inside = in_unit_cube(p.x, p.y, p.z);
    s_pts[j][iv] = make_float4(p.x, p.y, p.z, inside ? staged_a(p.std) : -1.0f);
    if (rays.points_mode == 1) __stcs(cache + j * rows + row, make_float4(p.x, p.y, p.z, inside ? p.std : -1.0f));
  }
}

// Interpolation partial of one lane of a pair: its x corner (cx + h) of the four (y, z) corners of one cell.
// `rowbase(k)` is the table row of corner k = (y bit) + 2 (z bit).
template <int C, class RowOf>
__device__ __forceinline__ void pair_partial(const float* __restrict__ emb, RowOf row_of, int h, float fx, float fy,
                                             float fz, float wj, float (&acc)[C]) {
  const float wx = h ? fx : 1.f - fx;
  const float wy0 = wx * (1.f - fy), wy1 = wx * fy;
  float r[4][C];
#pragma unroll
  for (int k = 0; k < 4; ++k) gather_row<C>(emb + row_of(k) * (uint32_t)C, r[k]);  // 32-bit row * C: offsets are int32
  const float w0 = wy0 * (1.f - fz), w1 = wy1 * (1.f - fz), w2 = wy0 * fz, w3 = wy1 * fz;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float f = w0 * r[0][c];
    f = fmaf(w1, r[1][c], f);
    f = fmaf(w2, r[2][c], f);
    f = fmaf(w3, r[3][c], f);
    acc[c] = fmaf(f, wj, acc[c]);
  }
}

// One level of one pair lane with CONSECUTIVE SAME-CELL SAMPLES MERGED: the lane's four corner weights (erf weight x
// trilinear weight) are summed over the samples of a cell -- exactly what the scatter kernels do -- and the cell's
// four rows are gathered once.  On the bench's rays an interval's 7 multisamples fall into 1.1 / 1.3 / 1.5 / 2.0 /
// 2.9 / 4.4 / 5.3 / 6.3 cells at resolutions 16 .. 2048 (tests/analysis_cell_sharing.py), so the 8-level proposal
// table needs 25 instead of 56 gather groups per interval; lanes that merge drop out of the gather instructions,
// whose cost is one L1 wavefront per distinct line.  Same value up to fp32 rounding: sum_k (sum_j e_j w_kj) r_k
// instead of sum_j e_j (sum_k w_kj r_k).
template <int C, bool kDense>
__device__ __forceinline__ void pair_level_merged(const float* __restrict__ emb, const Level3& lv, int h, int iv,
                                                  const float4 (*s_pts)[kPairIv], const float (*s_w)[kPairIv],
                                                  float (&acc)[C]) {
  uint32_t cx = 0, cy = 0, cz = 0;
  float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
  bool live = false;
  const uint32_t off = lv.offset;
#pragma unroll 1
  for (int j = 0; j <= 7; ++j) {  // the sentinel iteration flushes the last cell
    float4 p = make_float4(0.f, 0.f, 0.f, -1.f);
    if (j < 7) p = s_pts[j][iv];
    const bool valid = p.w >= 0.f;
    uint32_t nx = 0, ny = 0, nz = 0;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (valid) {
      cell_of(p.x, lv.scale, nx, fx);
      cell_of(p.y, lv.scale, ny, fy);
      cell_of(p.z, lv.scale, nz, fz);
    }
    if (live && (j == 7 || (valid && (nx != cx || ny != cy || nz != cz)))) {
      float r[4][C];
      if constexpr (kDense) {
        const uint32_t i00 = off + cx + h + cy * lv.s1 + cz * lv.s2;
        gather_row<C>(emb + i00 * (uint32_t)C, r[0]);
        gather_row<C>(emb + (i00 + lv.s1) * (uint32_t)C, r[1]);
        gather_row<C>(emb + (i00 + lv.s2) * (uint32_t)C, r[2]);
        gather_row<C>(emb + (i00 + lv.s1 + lv.s2) * (uint32_t)C, r[3]);
      } else {
        const uint32_t vx = cx + h;
        const uint32_t hy0 = cy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = cz * 805459861u, hz1 = hz0 + 805459861u;
        gather_row<C>(emb + (off + ((vx ^ hy0 ^ hz0) & lv.mask)) * (uint32_t)C, r[0]);
        gather_row<C>(emb + (off + ((vx ^ hy1 ^ hz0) & lv.mask)) * (uint32_t)C, r[1]);
        gather_row<C>(emb + (off + ((vx ^ hy0 ^ hz1) & lv.mask)) * (uint32_t)C, r[2]);
        gather_row<C>(emb + (off + ((vx ^ hy1 ^ hz1) & lv.mask)) * (uint32_t)C, r[3]);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float f = w0 * r[0][c];
        f = fmaf(w1, r[1][c], f);
        f = fmaf(w2, r[2][c], f);
        f = fmaf(w3, r[3][c], f);
        acc[c] += f;
      }
      w0 = w1 = w2 = w3 = 0.f;
      live = false;
    }
    if (valid) {
      cx = nx; cy = ny; cz = nz;
      live = true;
      const float coef = s_w[j][iv];
      const float wx = h ? fx : 1.f - fx;
      const float wy0 = wx * (1.f - fy), wy1 = wx * fy;
      w0 = fmaf(coef, wy0 * (1.f - fz), w0);
      w1 = fmaf(coef, wy1 * (1.f - fz), w1);
      w2 = fmaf(coef, wy0 * fz, w2);
      w3 = fmaf(coef, wy1 * fz, w3);
    }
  }
}

// Fused encode forward (any level_dim) with two lanes per interval -- see k_prop_fwd_pair below for the why.
// At C = 4 the x / x+1 rows of an even-x cell are one 32-byte sector, so the pair's loads also halve the
// L2 -> L1 sector traffic of those cells.
template <int C, bool kMerge>
__global__ void __launch_bounds__(kEncThreads) k_encode_fwd_pair(nlb_rays_t rays, nlb_table_t tab,
                                                                 float* __restrict__ features, int level_begin,
                                                                 int level_end) {
  __shared__ float4 s_pts[7][kPairIv];
  __shared__ float s_w[7][kPairIv];
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  __syncthreads();
  const int tid = threadIdx.x, iv = tid >> 1, h = tid & 1;
  const int row = blockIdx.x * kPairIv + iv;
  const bool ok = row < rays.N * rays.S;
  if (ok) stage_points_pair(rays, row, h, iv, s_pts);
  __syncwarp();
  const float* __restrict__ emb = tab.embeddings;
#pragma unroll 1
  for (int level = level_begin; level < level_end; ++level) {
    const Level3 lv = lc.lv[level];
    const float inv_gs = lc.inv_gs[level];
    if (ok) {
      for (int j = h; j < 7; j += 2) {
        const float a = s_pts[j][iv].w;
        s_w[j][iv] = a < 0.f ? 0.f : erf_weight_a(a, inv_gs);
      }
    }
    __syncwarp();
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (ok && kMerge) {
      if (lv.dense) pair_level_merged<C, true>(emb, lv, h, iv, s_pts, s_w, acc);  // uniform per level
      else pair_level_merged<C, false>(emb, lv, h, iv, s_pts, s_w, acc);
    } else if (ok) {
      const uint32_t off = lv.offset;
      if (lv.dense) {  // uniform per level
#pragma unroll 1
        for (int j = 0; j < 7; ++j) {
          const float4 p = s_pts[j][iv];
          if (p.w < 0.f) continue;
          uint32_t cx, cy, cz;
          float fx, fy, fz;
          cell_of(p.x, lv.scale, cx, fx);
          cell_of(p.y, lv.scale, cy, fy);
          cell_of(p.z, lv.scale, cz, fz);
          const uint32_t i00 = off + cx + h + cy * lv.s1 + cz * lv.s2;
          pair_partial<C>(emb, [&](int k) { return i00 + ((k & 1) ? lv.s1 : 0u) + ((k & 2) ? lv.s2 : 0u); }, h, fx, fy, fz,
                          s_w[j][iv], acc);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < 7; ++j) {
          const float4 p = s_pts[j][iv];
          if (p.w < 0.f) continue;
          uint32_t cx, cy, cz;
          float fx, fy, fz;
          cell_of(p.x, lv.scale, cx, fx);
          cell_of(p.y, lv.scale, cy, fy);
          cell_of(p.z, lv.scale, cz, fz);
          const uint32_t vx = cx + h;
          const uint32_t hy0 = cy * 2654435761u, hy1 = hy0 + 2654435761u;
          const uint32_t hz0 = cz * 805459861u, hz1 = hz0 + 805459861u;
          pair_partial<C>(emb, [&](int k) { return off + ((vx ^ ((k & 1) ? hy1 : hy0) ^ ((k & 2) ? hz1 : hz0)) & lv.mask); },
                          h, fx, fy, fz, s_w[j][iv], acc);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] += __shfl_xor_sync(NLB_FULL_MASK, acc[c], 1);
    if (ok && h == 0) {
      float* out = features + (size_t)row * (tab.L * C) + level * C;
      if constexpr (C == 4) {
        *reinterpret_cast<float4*>(out) = make_float4(acc[0] / 7.0f, acc[1] / 7.0f, acc[2] / 7.0f, acc[3] / 7.0f);
      } else if constexpr (C == 2) {
        *reinterpret_cast<float2*>(out) = make_float2(acc[0] / 7.0f, acc[1] / 7.0f);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] = acc[c] / 7.0f;
      }
    }
  }
}

// Scatter of a level group that holds only HASHED levels, two lanes per interval (the NeRF table's groups
// after the dense one): lane h accumulates and reduces the four (y, z) corners at x + h, so the pair's two
// reductions of one instruction fall into the same 128-byte line (the same 32-byte sector for even x at C = 4:
// one L2 request instead of two).  Same per-corner values as level_scatter (weights accumulated over the
// samples of a cell in the same order), persistent blocks over 64-interval tiles.
template <int C>
__global__ void __launch_bounds__(kEncThreads) k_encode_bwd_hashed_pair(nlb_rays_t rays, nlb_table_t tab,
                                                                        const float* __restrict__ grad_features,
                                                                        float* __restrict__ grad_table, int num_tiles,
                                                                        int level_begin, int level_end) {
  __shared__ float4 s_pts[7][kPairIv];
  __shared__ float s_w[7][kPairIv];  // erf weights of the current level, computed once per pair
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  __syncthreads();
  const int tid = threadIdx.x, iv = tid >> 1, h = tid & 1;
  const unsigned pair_mask = 3u << (tid & 30);
  const int rows_total = rays.N * rays.S;
  const int LC = tab.L * C;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row = tile * kPairIv + iv;
    const bool row_ok = row < rows_total;
    const float* gin = grad_features + (size_t)row * LC;
    float g_next[C];   // prefetched one level ahead (see k_encode_bwd); the first one travels under the staging
#pragma unroll
    for (int c = 0; c < C; ++c) g_next[c] = 0.f;
    if (row_ok) gather_row<C>(gin + level_begin * C, g_next);
    __syncwarp(pair_mask);  // the pair is done with the previous tile's points
    if (row_ok) stage_points_pair(rays, row, h, iv, s_pts);
    __syncwarp(pair_mask);
    if (!row_ok) continue;
#pragma unroll 1
    for (int level = level_begin; level < level_end; ++level) {
      float g[C];
      bool any = false;
#pragma unroll
      for (int c = 0; c < C; ++c) { g[c] = g_next[c] / 7.0f; any |= (g[c] != 0.f); }
      if (level + 1 < level_end) gather_row<C>(gin + (level + 1) * C, g_next);
      if (!any) continue;
      const Level3 lv = lc.lv[level];
      const float inv_gs = lc.inv_gs[level];
      __syncwarp(pair_mask);  // the partner is done with the previous level's weights
      for (int j = h; j < 7; j += 2) {
        const float a = s_pts[j][iv].w;
        s_w[j][iv] = a < 0.f ? 0.f : erf_weight_a(a, inv_gs);
      }
      __syncwarp(pair_mask);
      uint32_t cx = 0, cy = 0, cz = 0;
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      bool live = false;
#pragma unroll 1
      for (int j = 0; j <= 7; ++j) {  // the sentinel iteration flushes the last cell
        float4 p = make_float4(0.f, 0.f, 0.f, -1.f);
        if (j < 7) p = s_pts[j][iv];
        const bool valid = p.w >= 0.f;
        uint32_t nx = 0, ny = 0, nz = 0;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (valid) {
          cell_of(p.x, lv.scale, nx, fx);
          cell_of(p.y, lv.scale, ny, fy);
          cell_of(p.z, lv.scale, nz, fz);
        }
        if (live && (j == 7 || (valid && (nx != cx || ny != cy || nz != cz)))) {
          const uint32_t vx = cx + h;
          const uint32_t hy0 = cy * 2654435761u, hy1 = hy0 + 2654435761u;
          const uint32_t hz0 = cz * 805459861u, hz1 = hz0 + 805459861u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t idx = (vx ^ ((k & 1) ? hy1 : hy0) ^ ((k & 2) ? hz1 : hz0)) & lv.mask;
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = g[c] * w[k];
            red_add_row<C>(grad_table + (size_t)(lv.offset + idx) * C, v);
            w[k] = 0.f;
          }
          live = false;
        }
        if (valid) {
          cx = nx; cy = ny; cz = nz;
          live = true;
          const float coef = s_w[j][iv];
          const float wx = h ? fx : 1.f - fx;
          const float wy0 = wx * (1.f - fy), wy1 = wx * fy;
          w[0] = fmaf(coef, wy0 * (1.f - fz), w[0]);
          w[1] = fmaf(coef, wy1 * (1.f - fz), w[1]);
          w[2] = fmaf(coef, wy0 * fz, w[2]);
          w[3] = fmaf(coef, wy1 * fz, w[3]);
        }
      }
    }
  }
}

template <int L, bool kMerge>
__global__ void __launch_bounds__(kEncThreads) k_prop_fwd_pair(nlb_rays_t rays, nlb_table_t tab,
                                                               const float* __restrict__ W0, const float* __restrict__ b0,
                                                               const float* __restrict__ W1, const float* __restrict__ b1,
                                                               float* __restrict__ density, float* __restrict__ features) {
  __shared__ PropSmem sm;
  __shared__ float4 s_pts[7][kPairIv];
  __shared__ float s_w[7][kPairIv];
  __shared__ float s_f[L][kPairIv];
  __shared__ LevelCache lc;
  fill_level_cache(lc, tab);
  load_prop_weights(sm, L, W0, b0, W1, b1);
  __syncthreads();
  const int tid = threadIdx.x, iv = tid >> 1, h = tid & 1;
  const int row = blockIdx.x * kPairIv + iv;
  const bool ok = row < rays.N * rays.S;  // the same for both lanes of a pair; no early exit (warp shuffles below)
  if (ok) stage_points_pair(rays, row, h, iv, s_pts);
  __syncwarp();
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Level3 lv = lc.lv[l];
    const float inv_gs = lc.inv_gs[l];
    if (ok) {
      for (int j = h; j < 7; j += 2) {
        const float a = s_pts[j][iv].w;
        s_w[j][iv] = a < 0.f ? 0.f : erf_weight_a(a, inv_gs);
      }
    }
    __syncwarp();
    const float* __restrict__ emb = tab.embeddings;  // row = offset + idx in 32 bits (offsets are int32)
    const uint32_t off = lv.offset;
    float acc = 0.f;
    if (ok && kMerge) {
      float a1[1] = {0.f};
      if (lv.dense) pair_level_merged<1, true>(emb, lv, h, iv, s_pts, s_w, a1);  // uniform per level
      else pair_level_merged<1, false>(emb, lv, h, iv, s_pts, s_w, a1);
      acc = a1[0];
    } else if (ok) {
      if (lv.dense) {  // uniform per level
#pragma unroll 1
        for (int j = 0; j < 7; ++j) {
          const float4 p = s_pts[j][iv];
          if (p.w < 0.f) continue;
          uint32_t cx, cy, cz;
          float fx, fy, fz;
          cell_of(p.x, lv.scale, cx, fx);
          cell_of(p.y, lv.scale, cy, fy);
          cell_of(p.z, lv.scale, cz, fz);
          const float wx = h ? fx : 1.f - fx;
          const float wy0 = wx * (1.f - fy), wy1 = wx * fy;
          const uint32_t i00 = off + cx + h + cy * lv.s1 + cz * lv.s2;
          const float r0 = __ldg(emb + i00), r1 = __ldg(emb + (i00 + lv.s1));
          const float r2 = __ldg(emb + (i00 + lv.s2)), r3 = __ldg(emb + (i00 + lv.s1 + lv.s2));
          float f = (wy0 * (1.f - fz)) * r0;
          f = fmaf(wy1 * (1.f - fz), r1, f);
          f = fmaf(wy0 * fz, r2, f);
          f = fmaf(wy1 * fz, r3, f);
          acc = fmaf(f, s_w[j][iv], acc);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < 7; ++j) {
          const float4 p = s_pts[j][iv];
          if (p.w < 0.f) continue;
          uint32_t cx, cy, cz;
          float fx, fy, fz;
          cell_of(p.x, lv.scale, cx, fx);
          cell_of(p.y, lv.scale, cy, fy);
          cell_of(p.z, lv.scale, cz, fz);
          const float wx = h ? fx : 1.f - fx;
          const float wy0 = wx * (1.f - fy), wy1 = wx * fy;
          const uint32_t vx = cx + h;
          const uint32_t hy0 = cy * 2654435761u, hy1 = hy0 + 2654435761u;
          const uint32_t hz0 = cz * 805459861u, hz1 = hz0 + 805459861u;
          const float r0 = __ldg(emb + (off + ((vx ^ hy0 ^ hz0) & lv.mask))), r1 = __ldg(emb + (off + ((vx ^ hy1 ^ hz0) & lv.mask)));
          const float r2 = __ldg(emb + (off + ((vx ^ hy0 ^ hz1) & lv.mask))), r3 = __ldg(emb + (off + ((vx ^ hy1 ^ hz1) & lv.mask)));
          float f = (wy0 * (1.f - fz)) * r0;
          f = fmaf(wy1 * (1.f - fz), r1, f);
          f = fmaf(wy0 * fz, r2, f);
          f = fmaf(wy1 * fz, r3, f);
          acc = fmaf(f, s_w[j][iv], acc);
        }
      }
    }
    acc += __shfl_xor_sync(NLB_FULL_MASK, acc, 1);
    if (h == 0) s_f[l][iv] = acc / 7.0f;
    // the next level may overwrite s_w: every lane has passed the warp-wide shuffle above, i.e. has finished
    // reading this level's weights
  }
  __syncwarp();
  float f[L];
#pragma unroll
  for (int l = 0; l < L; ++l) f[l] = s_f[l][iv];
  if (features && ok) {
    float* out = features + (size_t)row * L;
    if ((L & 1) == 0 && (reinterpret_cast<uintptr_t>(features) & 7) == 0) {
#pragma unroll
      for (int q = 0; q < L / 2; ++q)
        if ((q & 1) == h) *reinterpret_cast<float2*>(out + 2 * q) = make_float2(f[2 * q], f[2 * q + 1]);
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l)
        if ((l & 1) == h) out[l] = f[l];
    }
  }
  // density MLP: lane h takes the hidden units k = 2 i + h (interleaved: conflict-free weight reads)
  float raw = 0.f;
#pragma unroll 4
  for (int i = 0; i < kPropHidden / 2; ++i) {
    const int k = 2 * i + h;
    float a = sm.b0[k];
#pragma unroll
    for (int l = 0; l < L; ++l) a = fmaf(sm.W0[k * L + l], f[l], a);
    raw = fmaf(sm.W1[k], fmaxf(a, 0.f), raw);
  }
  raw += __shfl_xor_sync(NLB_FULL_MASK, raw, 1);
  if (ok && h == 0) {
    const float xin = (raw + sm.b1) - 1.0f;  // softplus(raw + density_bias), density_bias = -1 (torch threshold 20)
    density[row] = xin > 20.f ? xin : log1pf(expf(xin));
  }
}

// Backward of the proposal MLP (persistent blocks over 128-row tiles): per-interval
// data gradient from the saved features -> grad_features[rows, L] for the scatter
// kernel.  The tile's weight gradients are a [64 x 128] [128 x L] product: thread
// (k = tid & 63, half = tid >> 6) walks its half of the tile's rows and keeps the L
// products of hidden unit k (plus its gb0 / gW1 terms) in registers across the block's
// tiles; one partial row per (block, half) is written at the end and summed by
// k_prop_wgrad_reduce (no atomics).
template <int L>
__global__ void __launch_bounds__(kEncThreads) k_prop_mlp_bwd(int rows_total, int num_tiles,
                                                              const float* __restrict__ W0, const float* __restrict__ b0,
                                                              const float* __restrict__ W1, const float* __restrict__ b1,
                                                              const float* __restrict__ features,
                                                              const float* __restrict__ grad_density,
                                                              float* __restrict__ grad_features,
                                                              float* __restrict__ partial /*[2*blocks][64*L+129]*/) {
  static_assert(kEncThreads == 2 * kPropHidden, "one thread per (hidden unit, row half)");
  __shared__ PropSmem sm;
  __shared__ float s_h[kEncThreads][kPropHidden + 1];  // relu output per row
  __shared__ __align__(16) float s_f[kEncThreads][(L + 3) / 4 * 4];
  __shared__ float s_graw[kEncThreads];
  load_prop_weights(sm, L, W0, b0, W1, b1);
  __syncthreads();
  const int tid = threadIdx.x;
  const int k_own = tid & (kPropHidden - 1), r_begin = (tid >> 6) * (kEncThreads / 2);
  const float w1_own = sm.W1[k_own];
  float aW0[L];
#pragma unroll
  for (int l = 0; l < L; ++l) aW0[l] = 0.f;
  float a_b0 = 0.f, a_W1 = 0.f, a_b1 = 0.f;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row = tile * kEncThreads + tid;
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
    if (valid) {
#pragma unroll
      for (int l = 0; l < L; ++l) grad_features[(size_t)row * L + l] = gf[l];
    }
    __syncthreads();
    // weight gradients of hidden unit k_own over this thread's half of the rows
#pragma unroll 2
    for (int r = r_begin; r < r_begin + kEncThreads / 2; ++r) {
      const float h = s_h[r][k_own], gr = s_graw[r];
      const float gh = h > 0.f ? gr * w1_own : 0.f;
      a_b0 += gh;
      a_W1 = fmaf(gr, h, a_W1);
      if (k_own == 0) a_b1 += gr;
#pragma unroll
      for (int l = 0; l < L; ++l) aW0[l] = fmaf(gh, s_f[r][l], aW0[l]);
    }
    __syncthreads();  // every thread is done reading s_h / s_f / s_graw
  }
  constexpr int kW0 = kPropHidden * L;
  constexpr int kEntries = kW0 + 2 * kPropHidden + 1;
  float* my = partial + (size_t)(2 * blockIdx.x + (tid >> 6)) * kEntries;
#pragma unroll
  for (int l = 0; l < L; ++l) my[k_own * L + l] = aW0[l];
  my[kW0 + k_own] = a_b0;                // gb0
  my[kW0 + kPropHidden + k_own] = a_W1;  // gW1
  if (k_own == 0) my[kW0 + 2 * kPropHidden] = a_b1;  // gb1
}

// sums the per-block partial weight gradients: one thread per entry, coalesced reads
__global__ void __launch_bounds__(256) k_prop_wgrad_reduce(const float* __restrict__ partial, int blocks, int entries,
                                                           int L, float* __restrict__ gW0, float* __restrict__ gb0,
                                                           float* __restrict__ gW1, float* __restrict__ gb1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  float a = 0.f;
#pragma unroll 8   // (18 dependent-free loads per thread were issued one round trip at a time)
  for (int b = blockIdx.y; b < blocks; b += gridDim.y) a += __ldg(partial + (size_t)b * entries + e);
  const int n0 = kPropHidden * L;
  float* dst = e < n0 ? gW0 + e : e < n0 + kPropHidden ? gb0 + (e - n0)
             : e < n0 + 2 * kPropHidden ? gW1 + (e - n0 - kPropHidden) : gb1;
  atomicAdd(dst, a);
}

}  // namespace nlb

using namespace nlb;

// Host view of the table's levels (from nlb_table_t.offsets_host), same arithmetic as the
// kernels: resolution = ceil(2^(l*S)*H - 1) + 1, dense iff (res+1)^3 fits the level.
struct HostLevels {
  int rows[kMaxLevelsEnc];
  bool dense[kMaxLevelsEnc];
};

static int check_rays_table(const nlb_rays_t* r, const nlb_table_t* t, const char* who, HostLevels* hl = nullptr) {
  if (!r || !t) { nlb_set_error("%s: null descriptor", who); return NLB_EINVAL; }
  if (r->N < 0 || r->S < 1) { nlb_set_error("%s: bad N/S", who); return NLB_EINVAL; }
  if (r->N == 0) return NLB_OK;  // empty batch: the callers return before launching (pointers may be NULL)
  if (t->L < 1 || t->L > kMaxLevelsEnc) { nlb_set_error("%s: num_levels %d outside [1,%d]", who, t->L, kMaxLevelsEnc); return NLB_EUNSUPPORTED; }
  if (!r->tdist || !r->origins || !r->directions || !r->radii || !r->base_x || !r->base_y || !t->embeddings ||
      !t->offsets || !t->grid_sizes) {
    nlb_set_error("%s: null pointer", who);
    return NLB_EINVAL;
  }
  if (!t->offsets_host) { nlb_set_error("%s: nlb_table_t.offsets_host (host copy of the level offsets) is required", who); return NLB_EINVAL; }
  if ((int64_t)r->N * r->S > 0x7fffffffLL / 16) { nlb_set_error("%s: N*S too large for one launch; chunk the rays", who); return NLB_EINVAL; }
  if (r->points_mode < 0 || r->points_mode > 2 || (r->points_mode != 0 && !r->points_cache)) {
    nlb_set_error("%s: points_mode %d needs a points_cache buffer", who, r->points_mode);
    return NLB_EINVAL;
  }
  for (int l = 0; l < t->L; ++l) {
    const int64_t size = (int64_t)t->offsets_host[l + 1] - t->offsets_host[l];
    if (size < 8) { nlb_set_error("%s: level %d has %lld rows", who, l, (long long)size); return NLB_EINVAL; }
    const float scale = exp2f(l * t->S) * t->H - 1.0f;
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    uint32_t s1, s2;
    const bool dense = level_is_dense(resolution, (uint32_t)size, s1, s2);
    if (!dense && (size & (size - 1)) != 0) {
      nlb_set_error("%s: hashed level %d has %lld rows; the fused kernels need power-of-two hashed levels "
                    "(GridEncoder always builds them so)", who, l, (long long)size);
      return NLB_EUNSUPPORTED;
    }
    if (hl) { hl->rows[l] = (int)size; hl->dense[l] = dense; }
  }
  return NLB_OK;
}

static long env_long(const char* name, long dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atol(v) : dflt;
}

extern "C" int nlb_sample_points(const nlb_rays_t* rays, float* points, void* stream) {
  if (!rays || !points) { nlb_set_error("sample_points: null pointer"); return NLB_EINVAL; }
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  k_sample_points<<<div_up(rows, 128), 128, 0, (cudaStream_t)stream>>>(*rays, points);
  return nlb_check_launch("sample_points");
}

static long env_long(const char* name, long dflt);

// Forward over all levels.  A table larger than L2 (the NeRF table: 240 MB) is walked in level groups
// that fit L2 together, like the scatter: random 16-byte gathers from an L2-resident 33.5 MB level instead
// of 1.9 GB of DRAM sector reads per call.  A group writes whole 32-byte sectors of the feature rows when
// its column range is sector-aligned, which the host checks; later groups re-read the sample points the
// first group cached (training) or regenerate them.
template <int C>
static int encode_forward_launch(const nlb_rays_t& rays_in, const nlb_table_t& tab, const HostLevels& hl,
                                 float* features, cudaStream_t st) {
  static const double kL2Budget = (double)env_long("NLB_GATHER_L2_MB", 70) * 1048576.0;
  const int rows = rays_in.N * rays_in.S;
  static const bool kLegacy = getenv("NLB_ENC_FWD_LEGACY") != nullptr;  // one lane per interval (A/B timing)
  static const bool kMerge = env_long("NLB_FWD_MERGE", 0) != 0;          // same-cell samples share one gather (measured slower)
  dim3 grid(div_up(rows, kLegacy ? kEncThreads : kPairIv));
  auto launch = [&](const nlb_rays_t& r, int l0, int l1) {
    if (kLegacy) k_encode_fwd<C><<<grid, kEncThreads, 0, st>>>(r, tab, features, l0, l1);
    else if (kMerge) k_encode_fwd_pair<C, true><<<grid, kEncThreads, 0, st>>>(r, tab, features, l0, l1);
    else k_encode_fwd_pair<C, false><<<grid, kEncThreads, 0, st>>>(r, tab, features, l0, l1);
  };
  double total = 0.;
  for (int l = 0; l < tab.L; ++l) total += (double)hl.rows[l] * C * 4.0;
  if (total <= kL2Budget || kL2Budget <= 0.) {
    launch(rays_in, 0, tab.L);
    return nlb_check_launch("encode_forward");
  }
  nlb_rays_t rays = rays_in;
  int l0 = 0;
  while (l0 < tab.L) {
    int l1 = l0;
    double bytes = 0.;
    while (l1 < tab.L && (l1 == l0 || bytes + (double)hl.rows[l1] * C * 4.0 <= kL2Budget)) bytes += (double)hl.rows[l1++] * C * 4.0;
    // keep group boundaries on 32-byte boundaries of the feature row
    while (l1 < tab.L && (l1 * C * 4) % 32 != 0) ++l1;
    launch(rays, l0, l1);
    if (int e = nlb_check_launch("encode_forward")) return e;
    if (rays.points_mode == 1) rays.points_mode = 2;  // the first group wrote the cache
    l0 = l1;
  }
  return NLB_OK;
}

extern "C" int nlb_encode_forward(const nlb_rays_t* rays, const nlb_table_t* table, float* features, void* stream) {
  HostLevels hl;
  if (int e = check_rays_table(rays, table, "encode_forward", &hl)) return e;
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (table->C) {
    case 1: return encode_forward_launch<1>(*rays, *table, hl, features, st);
    case 2: return encode_forward_launch<2>(*rays, *table, hl, features, st);
    case 4: return encode_forward_launch<4>(*rays, *table, hl, features, st);
    case 8: return encode_forward_launch<8>(*rays, *table, hl, features, st);
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
}

static int check_ray_grads(const nlb_ray_grads_t* g, const char* who) {
  if (!g || !g->origins || !g->directions || !g->base_x || !g->base_y) {
    nlb_set_error("%s: the four [N,3] gradient buffers (pre-zeroed or holding earlier levels' sums) are required", who);
    return NLB_EINVAL;
  }
  return NLB_OK;
}

// Like the forward, a table larger than L2 (the NeRF table) is walked in level groups that fit L2 together: the
// gradients are sums over the levels, so every group simply adds its share (measured on the 10 240-ray batch:
// 0.72 ms in one launch, DRAM-sector-bound on the four 33.5 MB hashed levels).
template <int C>
static int input_bwd_launch(const nlb_rays_t& rays, const nlb_table_t& tab, const HostLevels& hl, const float* gfeat,
                            const nlb_ray_grads_t& g, cudaStream_t st, const char* who) {
  static const double kL2Budget = (double)env_long("NLB_GATHER_L2_MB", 70) * 1048576.0;
  const int rows = rays.N * rays.S;
  int l0 = 0;
  while (l0 < tab.L) {
    int l1 = l0;
    double bytes = 0.;
    while (l1 < tab.L && (l1 == l0 || kL2Budget <= 0. || bytes + (double)hl.rows[l1] * C * 4.0 <= kL2Budget))
      bytes += (double)hl.rows[l1++] * C * 4.0;
    k_encode_input_bwd<C><<<div_up(rows * 7, kEncThreads), kEncThreads, 0, st>>>(rays, tab, gfeat, g.origins,
                                                                                g.directions, g.base_x, g.base_y, l0, l1);
    if (int e = nlb_check_launch(who)) return e;
    l0 = l1;
  }
  return NLB_OK;
}

extern "C" int nlb_encode_input_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* grad_features,
                                         const nlb_ray_grads_t* grads, void* stream) {
  HostLevels hl;
  if (int e = check_rays_table(rays, table, "encode_input_backward", &hl)) return e;
  if (rays->N == 0) return NLB_OK;
  if (!grad_features) { nlb_set_error("encode_input_backward: null grad_features"); return NLB_EINVAL; }
  if (int e = check_ray_grads(grads, "encode_input_backward")) return e;
  cudaStream_t st = (cudaStream_t)stream;
  switch (table->C) {
    case 1: return input_bwd_launch<1>(*rays, *table, hl, grad_features, *grads, st, "encode_input_backward");
    case 2: return input_bwd_launch<2>(*rays, *table, hl, grad_features, *grads, st, "encode_input_backward");
    case 4: return input_bwd_launch<4>(*rays, *table, hl, grad_features, *grads, st, "encode_input_backward");
    case 8: return input_bwd_launch<8>(*rays, *table, hl, grad_features, *grads, st, "encode_input_backward");
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
}

// Persistent launch shape of k_encode_bwd: the coarse levels that fit the shared-memory
// budget are staged per block (see the kernel comment); the grid is the number of blocks
// the device can keep resident (a multiple of the SM count), capped by the tile count.
static int sm_count() { return nlb_sm_count(); }


// privatised copies of the coarse rows (a block adds into copy blockIdx % copies); 4 .. 32 measured alike
static int priv_copies() {
  static const int v = (int)env_long("NLB_SCATTER_PRIV_COPIES", 16);
  return v < 1 ? 1 : v;
}
// Budget per copy: the resolution-64 level of a C = 1 table (1.26 MB with the two coarser ones) is privatised too --
// measured 0.497 -> 0.478 / 0.721 -> 0.692 ms for the two proposal scatters; the same level of the C = 4 table
// (5 MB per copy) measured slower (0.712 -> 0.734 ms) and stays on warp aggregation.  NLB_SCATTER_PRIV_KB overrides.
static size_t priv_budget_bytes() {
  static const size_t v = (size_t)env_long("NLB_SCATTER_PRIV_KB", 1400) * 1024;
  return v;
}

// rows of the leading dense levels that get privatised global copies (beyond the staged ones)
static int priv_rows_for(const nlb_table_t& tab, const HostLevels& hl) {
  int rows = 0;
  for (int l = 0; l < tab.L && hl.dense[l]; ++l) {
    if ((size_t)(rows + hl.rows[l]) * tab.C * sizeof(float) > priv_budget_bytes()) break;
    rows += hl.rows[l];
  }
  return rows;
}

static size_t scatter_workspace_floats(const nlb_table_t& tab, const HostLevels& hl) {
  return (size_t)priv_copies() * priv_rows_for(tab, hl) * tab.C;
}

template <int C>
static int scatter_launch(const nlb_rays_t& rays, const nlb_table_t& tab, const HostLevels& hl,
                          const float* grad_features, float* grad_embeddings, float* workspace, cudaStream_t st) {
  static const size_t kStageBudget = (size_t)env_long("NLB_SCATTER_STAGE_KB", 24) * 1024;  // >= 6 blocks/SM resident
  static const double kL2Budget = (double)env_long("NLB_SCATTER_L2_MB", 70) * 1048576.0;
  // (measured and rejected: the hashed levels of a table that fits L2 in a launch of their own without the
  // staging buffer -- 9 instead of 6 resident blocks per SM -- 0.62 / 0.90 ms against 0.52 / 0.79 for the
  // proposal backward: the second pass over the points costs more than the occupancy gains)
  // coarsest dense levels accumulated per block in shared memory
  int staged_rows = 0;
  for (int l = 0; l < tab.L && hl.dense[l]; ++l) {
    if ((size_t)(staged_rows + hl.rows[l]) * C * sizeof(float) > kStageBudget) break;
    staged_rows += hl.rows[l];
  }
  const size_t smem = (size_t)staged_rows * C * sizeof(float);
  int priv_rows = workspace ? priv_rows_for(tab, hl) : 0;
  if (priv_rows <= staged_rows) priv_rows = 0;
  if (priv_rows > 0 &&
      cudaMemsetAsync(workspace, 0, (size_t)priv_copies() * priv_rows * C * sizeof(float), st) != cudaSuccess)
    return nlb_check_launch("encode_backward memset");
  const int tiles = (int)div_up(rays.N * rays.S, kEncThreads);
  // per device (function attributes are device state): a process may drive several GPUs
  static bool attr_set[64] = {false};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  if (dev_ < 0 || dev_ >= 64) dev_ = 0;
  if (!attr_set[dev_]) {
    cudaFuncSetAttribute(k_encode_bwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    attr_set[dev_] = true;
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_bwd<C>, kEncThreads, smem);
  if (per_sm < 1) per_sm = 1;
  // development switches (A/B timing): fewer resident blocks per SM than the occupancy limit
  static const int kBlocksPerSm = (int)env_long("NLB_SCATTER_BLOCKS_PER_SM", 0);
  static const int kPairBlocksPerSm = (int)env_long("NLB_SCATTER_PAIR_BLOCKS_PER_SM", 0);
  if (kBlocksPerSm > 0 && kBlocksPerSm < per_sm) per_sm = kBlocksPerSm;
  int blocks_staged = sm_count() * per_sm;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_bwd<C>, kEncThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (kBlocksPerSm > 0 && kBlocksPerSm < per_sm) per_sm = kBlocksPerSm;
  int blocks_plain = sm_count() * per_sm;
  if (blocks_staged > tiles) blocks_staged = tiles;
  if (blocks_plain > tiles) blocks_plain = tiles;
  static const bool kLegacyHashed = getenv("NLB_SCATTER_HASHED_LEGACY") != nullptr;  // A/B timing
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_bwd_hashed_pair<C>, kEncThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (kPairBlocksPerSm > 0 && kPairBlocksPerSm < per_sm) per_sm = kPairBlocksPerSm;
  const int blocks_pair = sm_count() * per_sm;
  // level groups: consecutive levels whose gradient rows fit the L2 budget together, so a
  // fine level stays L2-resident while every tile updates it (level-major order)
  int l0 = 0;
  while (l0 < tab.L) {
    int l1 = l0;
    double bytes = 0.;
    while (l1 < tab.L && (l1 == l0 || bytes + (double)hl.rows[l1] * C * 4.0 <= kL2Budget)) bytes += (double)hl.rows[l1++] * C * 4.0;
    const bool first = l0 == 0;
    if (first && !kLegacyHashed) {
      // a table that fits L2 as ONE group: give its hashed levels a pair launch of their own when there are
      // enough of them to pay for the second pass over the points (measured, proposal backward: 8 levels with
      // 5 hashed 0.804 -> 0.761 ms, 6 levels with 3 hashed 0.523 -> 0.551 ms, so the latter stays one launch)
      int nd = 0;
      while (nd < tab.L && hl.dense[nd]) ++nd;
      if (nd > 0 && nd < l1 && l1 - nd >= 5) l1 = nd;
    }
    bool hashed_only = !kLegacyHashed;
    for (int l = l0; l < l1; ++l) hashed_only = hashed_only && !hl.dense[l];
    if (hashed_only) {  // two lanes per interval
      const int tiles2 = (int)div_up(rays.N * rays.S, kPairIv);
      k_encode_bwd_hashed_pair<C><<<blocks_pair < tiles2 ? blocks_pair : tiles2, kEncThreads, 0, st>>>(
          rays, tab, grad_features, grad_embeddings, tiles2, l0, l1);
    } else {
      k_encode_bwd<C><<<first ? blocks_staged : blocks_plain, kEncThreads, first ? smem : 0, st>>>(
          rays, tab, grad_features, grad_embeddings, first ? staged_rows : 0, tiles, l0, l1, workspace, priv_rows,
          priv_copies());
    }
    if (int e = nlb_check_launch("encode_backward")) return e;
    l0 = l1;
  }
  if (priv_rows > 0) {
    const int n = priv_rows * C;
    k_priv_reduce<<<div_up(n, 256), 256, 0, st>>>(workspace, priv_copies(), n, grad_embeddings);
    if (int e = nlb_check_launch("encode_backward reduce")) return e;
  }
  return NLB_OK;
}

extern "C" size_t nlb_encode_backward_workspace_bytes(const nlb_table_t* table) {
  if (!table || !table->offsets_host || table->L < 1 || table->L > kMaxLevelsEnc) return 0;
  HostLevels hl;
  for (int l = 0; l < table->L; ++l) {
    const int64_t size = (int64_t)table->offsets_host[l + 1] - table->offsets_host[l];
    const uint32_t resolution = (uint32_t)ceilf(exp2f(l * table->S) * table->H - 1.0f) + 1;
    uint32_t s1, s2;
    hl.rows[l] = (int)size;
    hl.dense[l] = level_is_dense(resolution, (uint32_t)size, s1, s2);
  }
  return scatter_workspace_floats(*table, hl) * sizeof(float);
}

extern "C" int nlb_encode_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* grad_features,
                                   float* grad_embeddings, float* workspace, void* stream) {
  HostLevels hl;
  if (int e = check_rays_table(rays, table, "encode_backward", &hl)) return e;
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (table->C) {
    case 1: return scatter_launch<1>(*rays, *table, hl, grad_features, grad_embeddings, workspace, st);
    case 2: return scatter_launch<2>(*rays, *table, hl, grad_features, grad_embeddings, workspace, st);
    case 4: return scatter_launch<4>(*rays, *table, hl, grad_features, grad_embeddings, workspace, st);
    case 8: return scatter_launch<8>(*rays, *table, hl, grad_features, grad_embeddings, workspace, st);
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
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
  static const bool kLegacy = getenv("NLB_PROP_FWD_LEGACY") != nullptr;  // one lane per interval (A/B timing)
  if (kLegacy) {
    NLB_PROP_DISPATCH(table->L, (k_prop_fwd<L_><<<div_up(rows, kEncThreads), kEncThreads, 0, st>>>(*rays, *table, W0, b0, W1, b1, density, features)));
  } else {
    static const bool kMerge = env_long("NLB_FWD_MERGE", 0) != 0;  // same-cell samples share one gather (measured slower)
    if (kMerge) {
      NLB_PROP_DISPATCH(table->L, (k_prop_fwd_pair<L_, true><<<div_up(rows, kPairIv), kEncThreads, 0, st>>>(*rays, *table, W0, b0, W1, b1, density, features)));
    } else {
      NLB_PROP_DISPATCH(table->L, (k_prop_fwd_pair<L_, false><<<div_up(rows, kPairIv), kEncThreads, 0, st>>>(*rays, *table, W0, b0, W1, b1, density, features)));
    }
  }
  return nlb_check_launch("prop_forward");
}

static int prop_bwd_blocks(int rows) {
  const int tiles = (int)div_up(rows, kEncThreads);
  const int want = sm_count() * 4;
  return tiles < want ? tiles : want;
}

// workspace layout: per-block weight-gradient partials | grad_features[rows, L] | privatised coarse rows
static size_t prop_ws_partial_floats(int rows, int L) {
  return (((size_t)2 * prop_bwd_blocks(rows) * (kPropHidden * L + 2 * kPropHidden + 1) + 63) / 64) * 64;
}
static size_t prop_ws_gfeat_floats(int rows, int L) { return (((size_t)rows * L + 63) / 64) * 64; }

extern "C" size_t nlb_prop_backward_workspace_bytes(int N, int S, const nlb_table_t* table) {
  if (!table) return 0;
  const int rows = N * S;
  return (prop_ws_partial_floats(rows, table->L) + prop_ws_gfeat_floats(rows, table->L)) * sizeof(float) +
         nlb_encode_backward_workspace_bytes(table);
}

extern "C" int nlb_prop_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* W0, const float* b0,
                                 const float* W1, const float* b1, const float* features, const float* grad_density,
                                 float* grad_embeddings, float* gW0, float* gb0, float* gW1, float* gb1,
                                 float* workspace, void* stream) {
  HostLevels hl;
  if (int e = check_rays_table(rays, table, "prop_backward", &hl)) return e;
  if (rays->N == 0) return NLB_OK;
  if (table->C != 1) { nlb_set_error("prop_backward: PropMLP tables have level_dim 1"); return NLB_EINVAL; }
  if (!features || !grad_density) { nlb_set_error("prop_backward: features saved by the forward are required"); return NLB_EINVAL; }
  if (!workspace) { nlb_set_error("prop_backward: workspace of nlb_prop_backward_workspace_bytes() is required"); return NLB_EINVAL; }
  const int rows = rays->N * rays->S;
  if (rows == 0) return NLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = prop_bwd_blocks(rows);
  const int tiles = (int)div_up(rows, kEncThreads);
  const int entries = kPropHidden * table->L + 2 * kPropHidden + 1;
  float* partial = workspace;
  float* gfeat = workspace + prop_ws_partial_floats(rows, table->L);  // 256-byte aligned
  float* priv = gfeat + prop_ws_gfeat_floats(rows, table->L);
  if (nlb_encode_backward_workspace_bytes(table) == 0) priv = nullptr;
  NLB_PROP_DISPATCH(table->L, (k_prop_mlp_bwd<L_><<<blocks, kEncThreads, 0, st>>>(
      rows, tiles, W0, b0, W1, b1, features, grad_density, gfeat, partial)));
  if (int e = nlb_check_launch("prop_mlp_backward")) return e;
  // 2 * blocks partial rows (1184 on a B200) over 64 row slices: 16 slices left 48 blocks on 148 SMs and the
  // launch latency-bound (37 us under ncu for 3 MB)
  k_prop_wgrad_reduce<<<dim3(div_up(entries, 256), 64), 256, 0, st>>>(partial, 2 * blocks, entries, table->L, gW0, gb0, gW1, gb1);
  if (int e = nlb_check_launch("prop_wgrad_reduce")) return e;
  return scatter_launch<1>(*rays, *table, hl, gfeat, grad_embeddings, priv, st);
}

// The feature gradients nlb_prop_backward left in its workspace (same N, S, table) -> ray-geometry gradients.
extern "C" int nlb_prop_input_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* workspace,
                                       const nlb_ray_grads_t* grads, void* stream) {
  HostLevels hl;
  if (int e = check_rays_table(rays, table, "prop_input_backward", &hl)) return e;
  if (rays->N == 0) return NLB_OK;
  if (table->C != 1) { nlb_set_error("prop_input_backward: PropMLP tables have level_dim 1"); return NLB_EINVAL; }
  if (!workspace) { nlb_set_error("prop_input_backward: the workspace nlb_prop_backward filled is required"); return NLB_EINVAL; }
  if (int e = check_ray_grads(grads, "prop_input_backward")) return e;
  const int rows = rays->N * rays->S;
  const float* gfeat = workspace + prop_ws_partial_floats(rows, table->L);
  return input_bwd_launch<1>(*rays, *table, hl, gfeat, *grads, (cudaStream_t)stream, "prop_input_backward");
}
