// Multi-resolution hash-grid encoder behind the reference's `_gridencoder` ABI
// (Z/gridencoder/src/gridencoder.h:12-15).  Generic over D in {2,3}, C in {1,2,4,8},
// hash / tiled grids, align_corners, linear / smoothstep interpolation.
//
// B200 notes: one thread per (point, level), level-major launch order (blockIdx.y
// = level) so a level's table (<= 33.5 MB) stays resident in the 126 MB L2 while
// all points visit it; C-wide vector gathers (LDG.128 for C=4); outputs written
// level-major [L,B,C] so every store is a full coalesced vector; backward uses
// one vector reduction (RED.ADD.F32x4 / x2) per corner instead of C scalar
// atomics.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {

template <uint32_t C>
struct Vec;
template <>
struct Vec<1> { using T = float; };
template <>
struct Vec<2> { using T = float2; };
template <>
struct Vec<4> { using T = float4; };

template <uint32_t C>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[C]) {
  if constexpr (C == 1) {
    v[0] = __ldg(p);
  } else if constexpr (C == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (uint32_t c = 0; c < C; c += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
      v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    }
  }
}

template <uint32_t C>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C == 1) {
    p[0] = v[0];
  } else if constexpr (C == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (uint32_t c = 0; c < C; c += 4)
      *reinterpret_cast<float4*>(p + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
}

// vector reduction into global memory (no return value -> RED)
template <uint32_t C>
__device__ __forceinline__ void red_row(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C == 1) {
    atomicAdd(p, v[0]);
  } else if constexpr (C == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
#pragma unroll
    for (uint32_t c = 0; c < C; c += 4)
      atomicAdd(reinterpret_cast<float4*>(p + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
  }
}

template <uint32_t D>
__device__ __forceinline__ bool load_point(const float* __restrict__ inputs, uint32_t b, float (&x)[D]) {
  bool oob = false;
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    x[d] = __ldg(inputs + (size_t)b * D + d);
    if (x[d] < 0.f || x[d] > 1.f) oob = true;
  }
  return oob;
}

template <uint32_t D>
__device__ __forceinline__ void locate(const float (&x)[D], float scale, bool align_corners, uint32_t interp,
                                       uint32_t (&pg)[D], float (&f)[D], float (&df)[D]) {
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    float pos = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
    float fl = floorf(pos);
    pg[d] = (uint32_t)fl;
    pos -= fl;
    if (interp == 1) {
      df[d] = 6.0f * pos * (1.0f - pos);
      pos = pos * pos * (3.0f - 2.0f * pos);
    } else {
      df[d] = 1.0f;
    }
    f[d] = pos;
  }
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_forward(const float* __restrict__ inputs,
                                                      const float* __restrict__ grid,
                                                      const int32_t* __restrict__ offsets,
                                                      float* __restrict__ outputs, uint32_t B, uint32_t L, float S,
                                                      uint32_t H, float* __restrict__ dy_dx, uint32_t gridtype,
                                                      bool align_corners, uint32_t interp) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  float x[D];
  const bool oob = load_point<D>(inputs, b, x);
  float* out = outputs + ((size_t)level * B + b) * C;
  float acc[C];
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) acc[c] = 0.f;
  if (oob) {
    store_row<C>(out, acc);
    if (dy_dx) {
      float* dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
      for (uint32_t i = 0; i < D * C; ++i) dd[i] = 0.f;
    }
    return;
  }
  const LevelGeom g = level_geom(offsets, level, S, H);
  const float* table = grid + (size_t)g.offset * C;
  uint32_t pg[D];
  float f[D], df[D];
  locate<D>(x, g.scale, align_corners, interp, pg, f, df);

  // issue all 2^D gathers, then combine in the reference's corner order
  float rows[1 << D][C];
  float w[1 << D];
#pragma unroll
  for (uint32_t i = 0; i < (1u << D); ++i) {
    float wi = 1.f;
    uint32_t q[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
      if (i & (1u << d)) { wi *= f[d]; q[d] = pg[d] + 1; }
      else { wi *= 1.f - f[d]; q[d] = pg[d]; }
    }
    w[i] = wi;
    const uint32_t idx = vertex_index<D>(q, g.hashmap_size, g.resolution, gridtype, align_corners);
    load_row<C>(table + (size_t)idx * C, rows[i]);
  }
#pragma unroll
  for (uint32_t i = 0; i < (1u << D); ++i)
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) acc[c] = fmaf(w[i], rows[i][c], acc[c]);
  store_row<C>(out, acc);

  if (dy_dx) {
    float* dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
#pragma unroll
    for (uint32_t gd = 0; gd < D; ++gd) {
      float gacc[C];
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) gacc[c] = 0.f;
#pragma unroll
      for (uint32_t i = 0; i < (1u << (D - 1)); ++i) {
        float wi = g.scale;
        uint32_t q[D];
#pragma unroll
        for (uint32_t nd = 0; nd < D - 1; ++nd) {
          const uint32_t d = (nd >= gd) ? nd + 1 : nd;
          if (i & (1u << nd)) { wi *= f[d]; q[d] = pg[d] + 1; }
          else { wi *= 1.f - f[d]; q[d] = pg[d]; }
        }
        q[gd] = pg[gd];
        const uint32_t il = vertex_index<D>(q, g.hashmap_size, g.resolution, gridtype, align_corners);
        q[gd] = pg[gd] + 1;
        const uint32_t ir = vertex_index<D>(q, g.hashmap_size, g.resolution, gridtype, align_corners);
        float lo[C], hi[C];
        load_row<C>(table + (size_t)il * C, lo);
        load_row<C>(table + (size_t)ir * C, hi);
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) gacc[c] += wi * (hi[c] - lo[c]) * df[gd];
      }
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) dd[gd * C + c] = gacc[c];
    }
  }
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_backward(const float* __restrict__ grad,
                                                       const float* __restrict__ inputs,
                                                       const int32_t* __restrict__ offsets,
                                                       float* __restrict__ grad_grid, uint32_t B, uint32_t L,
                                                       float S, uint32_t H, uint32_t gridtype, bool align_corners,
                                                       uint32_t interp) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  float x[D];
  if (load_point<D>(inputs, b, x)) return;  // grad buffer is pre-zeroed
  const LevelGeom g = level_geom(offsets, level, S, H);
  float* table = grad_grid + (size_t)g.offset * C;
  uint32_t pg[D];
  float f[D], df[D];
  locate<D>(x, g.scale, align_corners, interp, pg, f, df);
  float gr[C];
  load_row<C>(grad + ((size_t)level * B + b) * C, gr);
#pragma unroll
  for (uint32_t i = 0; i < (1u << D); ++i) {
    float wi = 1.f;
    uint32_t q[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
      if (i & (1u << d)) { wi *= f[d]; q[d] = pg[d] + 1; }
      else { wi *= 1.f - f[d]; q[d] = pg[d]; }
    }
    const uint32_t idx = vertex_index<D>(q, g.hashmap_size, g.resolution, gridtype, align_corners);
    float v[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) v[c] = wi * gr[c];
    red_row<C>(table + (size_t)idx * C, v);
  }
}

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_input_backward(const float* __restrict__ grad,
                                                        const float* __restrict__ dy_dx,
                                                        float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * D) return;
  const uint32_t b = t / D, d = t - b * D;
  const float* dd = dy_dx + (size_t)b * L * D * C;
  float r = 0.f;
  for (uint32_t l = 0; l < L; ++l)
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) r += __ldg(grad + ((size_t)l * B + b) * C + c) * __ldg(dd + l * D * C + d * C + c);
  grad_inputs[t] = r;
}

// gridencoder.cu:506-610
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grad_tv(const float* __restrict__ inputs, const float* __restrict__ grid,
                                                 float* __restrict__ grad, const int32_t* __restrict__ offsets,
                                                 float weight, uint32_t B, uint32_t L, float S, uint32_t H,
                                                 uint32_t gridtype, bool align_corners) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  float x[D];
  if (load_point<D>(inputs, b, x)) return;
  const LevelGeom g = level_geom(offsets, level, S, H);
  const float* table = grid + (size_t)g.offset * C;
  float* gtable = grad + (size_t)g.offset * C;
  uint32_t pg[D];
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) pg[d] = (uint32_t)floorf(fmaf(x[d], g.scale, align_corners ? 0.0f : 0.5f));
  float res[C], idelta[C], centre[C];
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) res[c] = idelta[c] = 0.f;
  const uint32_t index = vertex_index<D>(pg, g.hashmap_size, g.resolution, gridtype, align_corners);
  load_row<C>(table + (size_t)index * C, centre);
  const float w = weight / (2 * D);
#pragma unroll
  for (uint32_t d = 0; d < D; ++d) {
    const uint32_t cur = pg[d];
    if (cur < g.resolution) {
      pg[d] = cur + 1;
      float nb[C];
      load_row<C>(table + (size_t)vertex_index<D>(pg, g.hashmap_size, g.resolution, gridtype, align_corners) * C, nb);
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) { float gv = centre[c] - nb[c]; res[c] += gv; idelta[c] += gv * gv; }
    }
    if (cur > 0) {
      pg[d] = cur - 1;
      float nb[C];
      load_row<C>(table + (size_t)vertex_index<D>(pg, g.hashmap_size, g.resolution, gridtype, align_corners) * C, nb);
#pragma unroll
      for (uint32_t c = 0; c < C; ++c) { float gv = centre[c] - nb[c]; res[c] += gv; idelta[c] += gv * gv; }
    }
    pg[d] = cur;
  }
  float v[C];
#pragma unroll
  for (uint32_t c = 0; c < C; ++c) v[c] = w * res[c] * rsqrtf(idelta[c] + 1e-9f);
  red_row<C>(gtable + (size_t)index * C, v);
}

template <uint32_t D>
__global__ void __launch_bounds__(256) k_corner_indices(const float* __restrict__ inputs,
                                                        const int32_t* __restrict__ offsets,
                                                        uint32_t* __restrict__ indices, uint32_t B, uint32_t L, float S,
                                                        uint32_t H, uint32_t gridtype, bool align_corners) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t level = blockIdx.y;
  uint32_t* out = indices + ((size_t)level * B + b) * (1u << D);
  float x[D];
  if (load_point<D>(inputs, b, x)) {
    for (uint32_t i = 0; i < (1u << D); ++i) out[i] = 0xFFFFFFFFu;
    return;
  }
  const LevelGeom g = level_geom(offsets, level, S, H);
  uint32_t pg[D];
  float f[D], df[D];
  locate<D>(x, g.scale, align_corners, 0, pg, f, df);
#pragma unroll
  for (uint32_t i = 0; i < (1u << D); ++i) {
    uint32_t q[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) q[d] = pg[d] + ((i >> d) & 1u);
    out[i] = vertex_index<D>(q, g.hashmap_size, g.resolution, gridtype, align_corners);
  }
}

}  // namespace nlb

using namespace nlb;

#define NLB_DISPATCH_DC(D, C, ...)                                        \
  do {                                                                    \
    if (D == 3 && C == 1) { constexpr uint32_t D_ = 3, C_ = 1; __VA_ARGS__; } \
    else if (D == 3 && C == 2) { constexpr uint32_t D_ = 3, C_ = 2; __VA_ARGS__; } \
    else if (D == 3 && C == 4) { constexpr uint32_t D_ = 3, C_ = 4; __VA_ARGS__; } \
    else if (D == 3 && C == 8) { constexpr uint32_t D_ = 3, C_ = 8; __VA_ARGS__; } \
    else if (D == 2 && C == 1) { constexpr uint32_t D_ = 2, C_ = 1; __VA_ARGS__; } \
    else if (D == 2 && C == 2) { constexpr uint32_t D_ = 2, C_ = 2; __VA_ARGS__; } \
    else if (D == 2 && C == 4) { constexpr uint32_t D_ = 2, C_ = 4; __VA_ARGS__; } \
    else if (D == 2 && C == 8) { constexpr uint32_t D_ = 2, C_ = 8; __VA_ARGS__; } \
  } while (0)

static int check_dc(uint32_t D, uint32_t C) {
  if (!(C == 1 || C == 2 || C == 4 || C == 8)) {
    nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8.");
    return NLB_EINVAL;
  }
  if (D < 2 || D > 5) {
    nlb_set_error("GridEncoding: D must be 2, 3, 4 or 5.");
    return NLB_EINVAL;
  }
  if (D > 3) {
    nlb_set_error("GridEncoding: D=%u is not built in libnlb200 (zipnerf uses D=3)", D);
    return NLB_EUNSUPPORTED;
  }
  return NLB_OK;
}

extern "C" int nlb_grid_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets,
                                       float* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                       uint32_t H, float* dy_dx, uint32_t gridtype, int align_corners,
                                       uint32_t interp, void* stream) {
  if (int e = check_dc(D, C)) return e;
  if (B == 0 || L == 0) return NLB_OK;
  if (!inputs || !embeddings || !offsets || !outputs) { nlb_set_error("grid_encode_forward: null pointer"); return NLB_EINVAL; }
  dim3 grid(div_up(B, 256), L);
  NLB_DISPATCH_DC(D, C, (k_grid_forward<D_, C_><<<grid, 256, 0, (cudaStream_t)stream>>>(
      inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, align_corners != 0, interp)));
  return nlb_check_launch("grid_encode_forward");
}

extern "C" int nlb_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings,
                                        const int32_t* offsets, float* grad_embeddings, uint32_t B, uint32_t D,
                                        uint32_t C, uint32_t L, float S, uint32_t H, const float* dy_dx,
                                        float* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                        void* stream) {
  (void)embeddings;
  if (int e = check_dc(D, C)) return e;
  if (B == 0 || L == 0) return NLB_OK;
  if (!grad || !inputs || !offsets || !grad_embeddings) { nlb_set_error("grid_encode_backward: null pointer"); return NLB_EINVAL; }
  dim3 grid(div_up(B, 256), L);
  NLB_DISPATCH_DC(D, C, (k_grid_backward<D_, C_><<<grid, 256, 0, (cudaStream_t)stream>>>(
      grad, inputs, offsets, grad_embeddings, B, L, S, H, gridtype, align_corners != 0, interp)));
  if (int e = nlb_check_launch("grid_encode_backward")) return e;
  if (dy_dx && grad_inputs) {
    NLB_DISPATCH_DC(D, C, (k_input_backward<D_, C_><<<div_up(B * D, 256), 256, 0, (cudaStream_t)stream>>>(
        grad, dy_dx, grad_inputs, B, L)));
    return nlb_check_launch("grid_input_backward");
  }
  return NLB_OK;
}

extern "C" int nlb_grad_total_variation(const float* inputs, const float* embeddings, float* grad,
                                        const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                                        uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                        void* stream) {
  if (int e = check_dc(D, C)) return e;
  if (B == 0 || L == 0) return NLB_OK;
  if (!inputs || !embeddings || !grad || !offsets) { nlb_set_error("grad_total_variation: null pointer"); return NLB_EINVAL; }
  dim3 grid(div_up(B, 256), L);
  NLB_DISPATCH_DC(D, C, (k_grad_tv<D_, C_><<<grid, 256, 0, (cudaStream_t)stream>>>(
      inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, align_corners != 0)));
  return nlb_check_launch("grad_total_variation");
}

extern "C" int nlb_grid_corner_indices(const float* inputs, const int32_t* offsets, uint32_t* indices, uint32_t B,
                                       uint32_t D, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                       int align_corners, void* stream) {
  if (int e = check_dc(D, 1)) return e;
  if (B == 0 || L == 0) return NLB_OK;
  if (!inputs || !offsets || !indices) { nlb_set_error("grid_corner_indices: null pointer"); return NLB_EINVAL; }
  dim3 grid(div_up(B, 256), L);
  if (D == 3)
    k_corner_indices<3><<<grid, 256, 0, (cudaStream_t)stream>>>(inputs, offsets, indices, B, L, S, H, gridtype, align_corners != 0);
  else
    k_corner_indices<2><<<grid, 256, 0, (cudaStream_t)stream>>>(inputs, offsets, indices, B, L, S, H, gridtype, align_corners != 0);
  return nlb_check_launch("grid_corner_indices");
}
