// Shared device helpers for libnlb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define NLB_FULL_MASK 0xffffffffu

namespace nlb {

constexpr float kEps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
constexpr float kPi = 3.14159265358979323846f;

__host__ __device__ inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// Level geometry, identical arithmetic to the reference kernel
// (gridencoder.cu:137-139): scale = exp2f(level*S)*H - 1, resolution = ceil(scale)+1.
// ---------------------------------------------------------------------------
struct LevelGeom {
  float scale;
  uint32_t resolution;
  uint32_t hashmap_size;
  uint32_t offset;
};

__device__ __forceinline__ LevelGeom level_geom(const int32_t* __restrict__ offsets, uint32_t level, float S,
                                                uint32_t H) {
  LevelGeom g;
  g.offset = (uint32_t)__ldg(offsets + level);
  g.hashmap_size = (uint32_t)__ldg(offsets + level + 1) - g.offset;
  g.scale = exp2f(level * S) * H - 1.0f;
  g.resolution = (uint32_t)ceilf(g.scale) + 1;
  return g;
}

// Row index of one grid vertex (gridencoder.cu:50-84): dense stride walk while the
// stride fits, otherwise xor-of-primes hash, always modulo the level size.
template <uint32_t D>
__device__ __forceinline__ uint32_t vertex_index(const uint32_t pg[D], uint32_t hashmap_size, uint32_t resolution,
                                                 uint32_t gridtype, bool align_corners) {
  constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                  2097192037u, 1434869437u, 2165219737u};
  uint32_t stride = 1, index = 0;
  const uint32_t step = align_corners ? resolution : resolution + 1;
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    if (stride <= hashmap_size) {
      index += pg[d] * stride;
      stride *= step;
    }
  }
  if (gridtype == 0 && stride > hashmap_size) {
    uint32_t h = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) h ^= pg[d] * primes[d];
    index = h;
  }
  return index % hashmap_size;
}

// D = 3, hash grid, align_corners = false: the configuration of every encoder
// on the zipnerf path.  `dense` is a per-level constant.
struct Level3 {
  float scale;
  uint32_t hashmap_size;
  uint32_t offset;
  uint32_t s1, s2;  // dense strides (res+1), (res+1)^2
  uint32_t mask;    // hashmap_size - 1 (hashed levels are powers of two, validated on the host)
  bool dense;       // (res+1)^3 <= hashmap_size
};

// Host/device replay of the reference's stride walk (gridencoder.cu:66-84): a level is
// dense when (res+1)^3 fits the level, hashed otherwise.
__host__ __device__ inline bool level_is_dense(uint32_t resolution, uint32_t hashmap_size, uint32_t& s1, uint32_t& s2) {
  uint32_t stride = 1;
  const uint32_t step = resolution + 1;
  s1 = s2 = 0;
  bool stopped = false;
  for (int d = 0; d < 3; ++d) {
    if (!stopped && stride <= hashmap_size) {
      if (d == 1) s1 = stride;
      if (d == 2) s2 = stride;
      stride *= step;
    } else {
      stopped = true;
    }
  }
  return !(stride > hashmap_size);
}

__device__ __forceinline__ Level3 level3(const int32_t* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
  LevelGeom g = level_geom(offsets, level, S, H);
  Level3 v;
  v.scale = g.scale;
  v.hashmap_size = g.hashmap_size;
  v.offset = g.offset;
  v.mask = g.hashmap_size - 1;
  v.dense = level_is_dense(g.resolution, g.hashmap_size, v.s1, v.s2);
  return v;
}

// Row index of a vertex of a cell inside the unit cube, branch-free.  Dense levels need no
// modulo there (corner <= res on every axis, so idx <= (res+1)^3 - 1 < hashmap_size);
// hashed levels are powers of two (GridEncoder sizes them 2^log2_hashmap_size; the C API
// rejects anything else), so the reference's `% hashmap_size` is a mask.
__device__ __forceinline__ uint32_t vertex_index3(const Level3& lv, uint32_t x, uint32_t y, uint32_t z) {
  const uint32_t lin = x + y * lv.s1 + z * lv.s2;
  const uint32_t h = (x ^ (y * 2654435761u) ^ (z * 805459861u)) & lv.mask;
  return lv.dense ? lin : h;
}

// Same index with a (warp-uniform) branch on the level type and the reference's general
// modulo kept for non-power-of-two levels.  Which form ptxas schedules better was measured
// on B200 (tools/kernel_times.py): with this form it keeps all eight 16-byte corner loads
// of a C=4 lookup in flight (64 registers, 0.555 ms for the NeRF level forward) where the
// select form splits them in two groups (47 registers, 0.72 ms); for the 4-byte C=1
// lookups the select form wins (prop levels 0.38 / 0.67 ms against 0.55 / 1.01 ms, L1
// sector hit rate 51 % against 16 %).
__device__ __forceinline__ uint32_t vertex_index3_branchy(const Level3& lv, uint32_t x, uint32_t y, uint32_t z) {
  if (lv.dense) return x + y * lv.s1 + z * lv.s2;
  const uint32_t idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
  return ((lv.hashmap_size & (lv.hashmap_size - 1)) == 0) ? (idx & (lv.hashmap_size - 1)) : (idx % lv.hashmap_size);
}

// pos = x*scale + 0.5 is ONE fma in the reference binary (nvcc contracts
// gridencoder.cu:148); written explicitly so no compiler flag can change it.
__device__ __forceinline__ void cell_of(float x01, float scale, uint32_t& cell, float& frac) {
  float pos = fmaf(x01, scale, 0.5f);
  float fl = floorf(pos);
  cell = (uint32_t)fl;
  frac = pos - fl;
}

// ---------------------------------------------------------------------------
// Ray warps: power transformation with lambda (coord.py:103-162).
// Operation order mirrors the torch expression so results agree to ~1 ulp of powf.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float power_fwd(float x, float lam) {
  const float a = fabsf(lam - 1.0f);
  const float c = a / lam;
  return __fmul_rn(c, __fsub_rn(powf(__fadd_rn(__fdiv_rn(x, a), 1.0f), lam), 1.0f));
}

__device__ __forceinline__ float power_inv(float y, float lam) {
  const float a = fabsf(lam - 1.0f);
  const float inv = 1.0f / lam;
  float base = __fadd_rn(__fadd_rn(__fdiv_rn(__fmul_rn(y, lam), a), 1.0f), kEps);
  return __fmul_rn(__fsub_rn(powf(base, inv), 1.0f), a);
}

struct RayWarp {
  float s_near, s_far, lam;
};

__device__ __forceinline__ RayWarp make_warp(float near, float far, float lam) {
  RayWarp w;
  w.lam = lam;
  w.s_near = power_fwd(__fmul_rn(near, 2.0f), lam);
  w.s_far = power_fwd(__fmul_rn(far, 2.0f), lam);
  return w;
}

__device__ __forceinline__ float s_to_t(const RayWarp& w, float s) {
  float y = __fadd_rn(__fmul_rn(s, w.s_far), __fmul_rn(__fsub_rn(1.0f, s), w.s_near));
  return __fdiv_rn(power_inv(y, w.lam), 2.0f);
}

// ---------------------------------------------------------------------------
// Sample-point generation: render.cast_rays (render.py:129-168) for multisample j
// of the interval [t0,t1], then contract_mean_std (coord.py:51-63), the /2 of
// models.py:970-973 and GridEncoder's (x+1)/2 (grid.py:162).
// ---------------------------------------------------------------------------
struct RayGeom {
  float ox, oy, oz;
  float dx, dy, dz;
  float bxx, bxy, bxz;
  float byx, byy, byz;
  float radius;
};

__device__ __forceinline__ RayGeom load_ray(const float* __restrict__ origins, const float* __restrict__ directions,
                                            const float* __restrict__ base_x, const float* __restrict__ base_y,
                                            const float* __restrict__ radii, int ray) {
  RayGeom r;
  r.ox = __ldg(origins + 3 * ray);
  r.oy = __ldg(origins + 3 * ray + 1);
  r.oz = __ldg(origins + 3 * ray + 2);
  r.dx = __ldg(directions + 3 * ray);
  r.dy = __ldg(directions + 3 * ray + 1);
  r.dz = __ldg(directions + 3 * ray + 2);
  r.bxx = __ldg(base_x + 3 * ray);
  r.bxy = __ldg(base_x + 3 * ray + 1);
  r.bxz = __ldg(base_x + 3 * ray + 2);
  r.byx = __ldg(base_y + 3 * ray);
  r.byy = __ldg(base_y + 3 * ray + 1);
  r.byz = __ldg(base_y + 3 * ray + 2);
  r.radius = __ldg(radii + ray);
  return r;
}

struct SamplePoint {
  float x, y, z;  // in [0,1]^3 (grid coordinates)
  float std;      // contracted std / 2
};

// Intermediates of one sample for the gradient w.r.t. the ray geometry (k_encode_input_bwd): the
// pre-contraction mean m = lx*base_x + ly*base_y + t*d + o, |m|^2, the contraction factor k (1 inside the
// unit ball) and the uncontracted std.
struct SampleGeom {
  float t, lx, ly;
  float mx, my, mz, m2, k, sd;
};

// j in [0,7): t = t0 + (t1-t0)*(j+.5)/7 ; deg = 2*pi*3*j/7 (+ 2*pi*noise).
__device__ __forceinline__ SamplePoint sample_point(const RayGeom& r, float t0, float t1, int j, float noise,
                                                    bool has_noise, float std_scale, SampleGeom* geo = nullptr) {
  const float n = 7.0f;
  float t = __fadd_rn(t0, __fdiv_rn(__fmul_rn(__fsub_rn(t1, t0), (float)j + 0.5f), n));
  // 2*pi*m is folded in double then rounded (python float -> float32 scalar)
  float deg = __fdiv_rn(__fmul_rn(18.849555921538759f, (float)j), n);
  if (has_noise) deg = __fadd_rn(deg, __fmul_rn(__fmul_rn(noise, kPi), 2.0f));
  float sn, cs;
  sincosf(deg, &sn, &cs);
  float rt = __fmul_rn(r.radius, t);
  // x / 2 is written x * 0.5f throughout: the same correctly rounded value without the IEEE-division sequence
  float lx = __fmul_rn(__fmul_rn(rt, cs), 0.5f);
  float ly = __fmul_rn(__fmul_rn(rt, sn), 0.5f);
  float mx = fmaf(lx, r.bxx, fmaf(ly, r.byx, t * r.dx)) + r.ox;
  float my = fmaf(lx, r.bxy, fmaf(ly, r.byy, t * r.dy)) + r.oy;
  float mz = fmaf(lx, r.bxz, fmaf(ly, r.byz, t * r.dz)) + r.oz;
  float sd = __fmul_rn(__fmul_rn(std_scale, r.radius), t);
  // contract
  float m2 = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)), kEps);
  if (geo) {
    geo->t = t; geo->lx = lx; geo->ly = ly;
    geo->mx = mx; geo->my = my; geo->mz = mz;
    geo->m2 = m2; geo->k = 1.0f; geo->sd = sd;
  }
  if (!(m2 <= 1.0f)) {
    float m = sqrtf(m2);
    float k = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, m), 1.0f), m2);
    if (geo) geo->k = k;
    mx *= k;
    my *= k;
    mz *= k;
    float q = __fsub_rn(__fdiv_rn(2.0f, m), __fdiv_rn(1.0f, m2));
    float det = __fmul_rn(__fdiv_rn(1.0f, m2), __fmul_rn(q, q));
    sd = __fmul_rn(powf(det, 0.3333333333333333f), sd);
  }
  SamplePoint p;
  // /2 (contract [-2,2] -> [-1,1]) then (x + 1) / 2
  p.x = __fmul_rn(__fadd_rn(__fmul_rn(mx, 0.5f), 1.0f), 0.5f);
  p.y = __fmul_rn(__fadd_rn(__fmul_rn(my, 0.5f), 1.0f), 0.5f);
  p.z = __fmul_rn(__fadd_rn(__fmul_rn(mz, 0.5f), 1.0f), 0.5f);
  p.std = __fmul_rn(sd, 0.5f);
  return p;
}

// erf(1 / max(sqrt(8 * std^2 * grid_size^2), 1e-10))  (models.py:976)
__device__ __forceinline__ float erf_weight(float sd, int grid_size) {
  float gs = (float)(grid_size * grid_size);
  float v = __fmul_rn(__fmul_rn(8.0f, __fmul_rn(sd, sd)), gs);
  return erff(__fdiv_rn(1.0f, fmaxf(sqrtf(v), 1e-10f)));
}

__device__ __forceinline__ bool in_unit_cube(float x, float y, float z) {
  return !(x < 0.f || x > 1.f || y < 0.f || y > 1.f || z < 0.f || z > 1.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NLB_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(NLB_FULL_MASK, v, o));
  return v;
}

// inclusive warp scan (sum)
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(NLB_FULL_MASK, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

}  // namespace nlb

// error plumbing shared by all translation units
void nlb_set_error(const char* fmt, ...);
int nlb_check_launch(const char* what);
// device buffer of step-varying scalars {anneal, lr_c, rsqrt_bc2} or nullptr (nlb_set_dynamic_scalars)
const float* nlb_dynamic_scalars();
// SM count of the current device (per-device cache, capi.cu)
int nlb_sm_count();
enum { NLB_DYN_ANNEAL = 0, NLB_DYN_LR_C = 1, NLB_DYN_RSQRT_BC2 = 2 };
