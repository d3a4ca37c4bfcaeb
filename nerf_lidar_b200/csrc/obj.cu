// Dynamic-object branch of Model.forward (Z/internal/models.py:306-315,401-477,
// Z/internal/obj_utils.py:76-111 rotate_yaw_z, :5-29 scale_frames, :116-194 world2object, :196-234 box_pts,
// :431-475 get_pose).
//
// The reference, per sampling level and per track: transforms every sample midpoint of every ray into the
// object's box frame, builds a boolean intersection map, synchronises the host (`intersect_idx.sum() == 0`),
// compacts the hit points with boolean indexing, evaluates the class's ObjMLP on them (hash grid L7 x C2 +
// split shape / texture latent, Z/internal/models.py:1000-1034,1036-1263 with warp_fn=None, re_weights=False,
// fixed_semantic=True) and merges the results back with zeros_like + masked assignment + where for every key.
//
// Here ONE kernel per (level, track) does all of it with no host round trip and no intermediate tensor:
// a thread per (ray, sample) runs the box test; the hit lanes of a warp are then served one after the other
// by the WHOLE warp -- lane l gathers grid level l, lanes split the output units of each dense layer, whose
// transposed weights sit in shared memory -- and the results overwrite density / rgb / semantic in place
// ("compaction" is the warp ballot; the masked merge is the store).  Later tracks overwrite earlier ones,
// as the reference's loop order does.
#include "common.cuh"
#include "../../include/nlb200.h"
#include <algorithm>

namespace nlb {
namespace obj {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kInfo = 9;  // centre(3), yaw, wlh(3), timestamp, track id (Z/internal/datasets.py:1442-1452)

// obj_utils.get_pose: the two track entries closest in time, blended by |t - t2| / (|t1 - t2| + 1e-9)
__global__ void k_obj_pose(const float* __restrict__ time, const float* __restrict__ tracks, int N, int n_obj, int T,
                           float* __restrict__ pose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * n_obj) return;
  const int ray = i / n_obj, o = i % n_obj;
  const float t = __ldg(time + ray);
  const float* tr = tracks + (size_t)o * T * kInfo;
  float d1 = INFINITY, d2 = INFINITY;
  int i1 = 0, i2 = 0;
  for (int k = 0; k < T; ++k) {  // ascending sort of |dt|, first two (stable for ties)
    const float d = fabsf(__fsub_rn(t, __ldg(tr + k * kInfo + kInfo - 2)));
    if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = k; }
    else if (d < d2) { d2 = d; i2 = k; }
  }
  const float t1 = __ldg(tr + i1 * kInfo + kInfo - 2), t2 = __ldg(tr + i2 * kInfo + kInfo - 2);
  const float total = __fadd_rn(fabsf(__fsub_rn(t1, t2)), 1e-9f);
  float w1 = __fdiv_rn(fabsf(__fsub_rn(t, t2)), total);
  w1 = fminf(fmaxf(w1, 0.f), 1.f);
  const float w2 = __fsub_rn(1.f, w1);
#pragma unroll
  for (int f = 0; f < kInfo; ++f)
    pose[(size_t)i * kInfo + f] = __fadd_rn(__fmul_rn(w1, __ldg(tr + i1 * kInfo + f)), __fmul_rn(w2, __ldg(tr + i2 * kInfo + f)));
}

// obj_utils.rotate_yaw_z INCLUDING its in-place quirk: p_y uses the already rotated p_x
// (`p_x = c*p_x - s*p_y; p_y = s*p_x + c*p_y`, obj_utils.py:103-106).
__device__ __forceinline__ void rotate_yaw_z(float& x, float& y, float c, float s) {
  x = __fsub_rn(__fmul_rn(c, x), __fmul_rn(s, y));
  y = __fadd_rn(__fmul_rn(s, x), __fmul_rn(c, y));
}


// World -> box frame of one sample midpoint (models.py:403-404, obj_utils.py:158-190,196-206); returns the hit flag.
__device__ __forceinline__ bool box_point(const float* __restrict__ tdist, const float* __restrict__ origins,
                                          const float* __restrict__ directions, const float* __restrict__ viewdirs,
                                          const float* __restrict__ pose, int n_obj, int track, int S, int pt, float& px,
                                          float& py, float& pz, float& vx, float& vy, float& vz) {
  const int ray = pt / S, s = pt - ray * S;
  const float* td = tdist + (size_t)ray * (S + 1) + s;
  const float tm = __fmul_rn(0.5f, __fadd_rn(__ldg(td), __ldg(td + 1)));
  px = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray)), __ldg(origins + 3 * ray));
  py = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray + 1)), __ldg(origins + 3 * ray + 1));
  pz = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray + 2)), __ldg(origins + 3 * ray + 2));
  const float* ps = pose + ((size_t)ray * n_obj + track) * kInfo;
  const float theta = __ldg(ps + 3);
  const float c = cosf(theta), sn = sinf(theta);
  float tx = -__ldg(ps), ty = -__ldg(ps + 1);
  const float tz = -__ldg(ps + 2);
  rotate_yaw_z(tx, ty, c, sn);
  rotate_yaw_z(px, py, c, sn);
  px = __fadd_rn(px, tx); py = __fadd_rn(py, ty); pz = __fadd_rn(pz, tz);
  vx = __ldg(viewdirs + 3 * ray); vy = __ldg(viewdirs + 3 * ray + 1); vz = __ldg(viewdirs + 3 * ray + 2);
  rotate_yaw_z(vx, vy, c, sn);
  const float sx = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 4), 0.5f), 1e-9f));
  const float sy = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 5), 0.5f), 1e-9f));
  const float sz = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 6), 0.5f), 1e-9f));
  px = __fmul_rn(sx, px); py = __fmul_rn(sy, py); pz = __fmul_rn(sz, pz);
  vx = __fmul_rn(sx, vx); vy = __fmul_rn(sy, vy); vz = __fmul_rn(sz, vz);
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  vx = __fdiv_rn(vx, nrm); vy = __fdiv_rn(vy, nrm); vz = __fdiv_rn(vz, nrm);
  return fabsf(px) < 1.f && fabsf(py) < 1.f && fabsf(pz) < 1.f;
}

// Backward of box_point w.r.t. the interpolated pose (track refinement, Z/train.py:244-257: the track's centre and
// yaw are functions of Track_opt's corrections): gq = dL/d(box coordinates), gu = dL/d(unit view direction in the
// box frame) -> gp[0..6] = dL/d(centre xyz, yaw, wlh).  rotate_yaw_z with its quirk is x' = c x - s y,
// y' = s x' + c y, linear in (x, y): rot(p) + rot(-t) = rot(p - t).
__device__ __forceinline__ void box_point_grad(const float* __restrict__ tdist, const float* __restrict__ origins,
                                               const float* __restrict__ directions, const float* __restrict__ viewdirs,
                                               const float* __restrict__ pose, int n_obj, int track, int S, int pt,
                                               const float (&gq)[3], const float (&gu)[3], float (&gp)[7]) {
  const int ray = pt / S, s = pt - ray * S;
  const float* td = tdist + (size_t)ray * (S + 1) + s;
  const float tm = 0.5f * (__ldg(td) + __ldg(td + 1));
  const float* ps = pose + ((size_t)ray * n_obj + track) * kInfo;
  const float ax = tm * __ldg(directions + 3 * ray) + __ldg(origins + 3 * ray) - __ldg(ps);
  const float ay = tm * __ldg(directions + 3 * ray + 1) + __ldg(origins + 3 * ray + 1) - __ldg(ps + 1);
  const float az = tm * __ldg(directions + 3 * ray + 2) + __ldg(origins + 3 * ray + 2) - __ldg(ps + 2);
  const float theta = __ldg(ps + 3);
  const float c = cosf(theta), sn = sinf(theta);
  float sc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) sc[k] = 1.f / (__ldg(ps + 4 + k) * 0.5f + 1e-9f);
  // position
  const float x1 = c * ax - sn * ay, y1 = sn * x1 + c * ay;
  const float gx1 = gq[0] * sc[0], gy1 = gq[1] * sc[1], gz1 = gq[2] * sc[2];
  const float dx1 = -sn * ax - c * ay, dy1 = c * x1 + sn * dx1 - sn * ay;
  gp[0] = -(gx1 * c + gy1 * sn * c);
  gp[1] = -(-gx1 * sn + gy1 * (c - sn * sn));
  gp[2] = -gz1;
  gp[3] = gx1 * dx1 + gy1 * dy1;
  gp[4] = -0.5f * gq[0] * x1 * sc[0] * sc[0];
  gp[5] = -0.5f * gq[1] * y1 * sc[1] * sc[1];
  gp[6] = -0.5f * gq[2] * az * sc[2] * sc[2];
  // view direction: u = w / |w|, w = rot(v) * scale
  const float v0 = __ldg(viewdirs + 3 * ray), v1 = __ldg(viewdirs + 3 * ray + 1), v2 = __ldg(viewdirs + 3 * ray + 2);
  const float vx1 = c * v0 - sn * v1, vy1 = sn * vx1 + c * v1;
  const float w[3] = {vx1 * sc[0], vy1 * sc[1], v2 * sc[2]};
  const float nrm = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const float u[3] = {w[0] / nrm, w[1] / nrm, w[2] / nrm};
  const float ug = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
  const float gw[3] = {(gu[0] - u[0] * ug) / nrm, (gu[1] - u[1] * ug) / nrm, (gu[2] - u[2] * ug) / nrm};
  const float dvx1 = -sn * v0 - c * v1, dvy1 = c * vx1 + sn * dvx1 - sn * v1;
  gp[3] += gw[0] * sc[0] * dvx1 + gw[1] * sc[1] * dvy1;
  gp[4] += -0.5f * gw[0] * vx1 * sc[0] * sc[0];
  gp[5] += -0.5f * gw[1] * vy1 * sc[1] * sc[1];
  gp[6] += -0.5f * gw[2] * v2 * sc[2] * sc[2];
}

struct Dims {
  int F, K0, hidden, bott, dir_dim, Kv, vw;  // K0 = F + latent_shape, Kv = bott + dir_dim + latent_tex
  int o_wd0, o_bd0, o_wd2, o_bd2, o_wv0, o_bv0, o_wv1, o_bv1, o_wrgb, o_brgb, total;  // float offsets in shared memory
  int scratch;  // floats per warp
};
__host__ __device__ inline Dims make_dims(const nlb_obj_mlp_t& m, int F) {
  Dims d;
  d.F = F; d.K0 = F + m.latent_shape; d.hidden = m.hidden; d.bott = m.bottleneck;
  d.dir_dim = 3 + 6 * m.deg_view; d.Kv = m.bottleneck + d.dir_dim + m.latent_tex; d.vw = m.view_width;
  int o = 0;
  d.o_wd0 = o; o += d.K0 * d.hidden;
  d.o_bd0 = o; o += d.hidden;
  d.o_wd2 = o; o += d.hidden * d.bott;
  d.o_bd2 = o; o += d.bott;
  d.o_wv0 = o; o += d.Kv * d.vw;
  d.o_bv0 = o; o += d.vw;
  d.o_wv1 = o; o += (d.vw + d.Kv) * d.vw;
  d.o_bv1 = o; o += d.vw;
  d.o_wrgb = o; o += d.vw * 3;
  d.o_brgb = o; o += 4;
  d.total = (o + 3) & ~3;
  d.scratch = (d.K0 + d.hidden + d.vw + d.Kv + d.vw + 3) & ~3;
  return d;
}

// W [J, K] row-major in global memory -> Wt[k * J + j] in shared memory (lanes read consecutive j)
__device__ __forceinline__ void stage_transposed(float* __restrict__ dst, const float* __restrict__ W, int J, int K) {
  for (int e = threadIdx.x; e < J * K; e += kThreads) {
    const int j = e / K, k = e % K;
    dst[k * J + j] = __ldg(W + e);
  }
}

// out[j] = act(b[j] + sum_k Wt[k * J + j] * in[k]) for the warp's lanes j, j + 32, ...
template <bool kRelu>
__device__ __forceinline__ void dense(const float* __restrict__ Wt, const float* __restrict__ b, const float* __restrict__ in,
                                      int K, int J, float* __restrict__ out, int lane) {
  for (int j = lane; j < J; j += 32) {
    float acc = b[j];
    for (int k = 0; k < K; ++k) acc = fmaf(Wt[k * J + j], in[k], acc);
    out[j] = kRelu ? fmaxf(acc, 0.f) : acc;
  }
}

template <int C>
__device__ __forceinline__ void level_features(const nlb_table_t& tab, int level, float x, float y, float z, float* out) {
  const Level3 lv = level3(tab.offsets, level, tab.S, tab.H);
  uint32_t cx, cy, cz;
  float fx, fy, fz;
  cell_of(x, lv.scale, cx, fx);
  cell_of(y, lv.scale, cy, fy);
  cell_of(z, lv.scale, cz, fz);
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // weight product in the reference kernel's order (gridencoder.cu:166-191)
    float w = 1.f;
    w *= (i & 1) ? fx : 1.f - fx;
    w *= (i & 2) ? fy : 1.f - fy;
    w *= (i & 4) ? fz : 1.f - fz;
    const uint32_t idx = vertex_index3_branchy(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
    const float* row = tab.embeddings + ((size_t)lv.offset + idx) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fmaf(w, __ldg(row + c), acc[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) out[level * C + c] = acc[c];
}

template <int C>
__global__ void __launch_bounds__(kThreads) k_obj_forward(const float* __restrict__ tdist, const float* __restrict__ origins,
                                                         const float* __restrict__ directions,
                                                         const float* __restrict__ viewdirs, const float* __restrict__ pose,
                                                         int n_obj, int track, int N, int S, nlb_table_t tab,
                                                         nlb_obj_mlp_t m, float* __restrict__ density,
                                                         float* __restrict__ rgb, float* __restrict__ semantic,
                                                         uint8_t* __restrict__ obj_mask, int32_t* __restrict__ owner,
                                                         int num_tiles) {
  extern __shared__ __align__(16) float smem[];
  const Dims d = make_dims(m, tab.L * C);
  float* W = smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = smem + d.total + warp * d.scratch;
  float* in0 = sc;                        // [K0]  grid features | shape latent
  float* h = in0 + d.K0;                  // [hidden]
  float* h1 = h + d.hidden;               // [vw]   } contiguous: the input of lin_second_stage_1 is cat[h1, inv]
  float* inv = h1 + d.vw;                 // [Kv]   } bottleneck | dir enc | texture latent
  float* h2 = inv + d.Kv;                 // [vw]
  const bool want_rgb = rgb != nullptr;
  stage_transposed(W + d.o_wd0, m.W_d0, d.hidden, d.K0);
  stage_transposed(W + d.o_wd2, m.W_d2, d.bott, d.hidden);
  for (int e = threadIdx.x; e < d.hidden; e += kThreads) W[d.o_bd0 + e] = __ldg(m.b_d0 + e);
  for (int e = threadIdx.x; e < d.bott; e += kThreads) W[d.o_bd2 + e] = __ldg(m.b_d2 + e);
  if (want_rgb) {
    stage_transposed(W + d.o_wv0, m.W_v0, d.vw, d.Kv);
    stage_transposed(W + d.o_wv1, m.W_v1, d.vw, d.vw + d.Kv);
    for (int e = threadIdx.x; e < d.vw; e += kThreads) { W[d.o_bv0 + e] = __ldg(m.b_v0 + e); W[d.o_bv1 + e] = __ldg(m.b_v1 + e); }
    for (int e = threadIdx.x; e < 3 * d.vw; e += kThreads) W[d.o_wrgb + e] = __ldg(m.W_rgb + e);   // [3, vw] as is
    if (threadIdx.x < 3) W[d.o_brgb + threadIdx.x] = __ldg(m.b_rgb + threadIdx.x);
  }
  // the latent halves never change between points: written once per warp
  for (int k = lane; k < m.latent_shape; k += 32) in0[d.F + k] = __ldg(m.latent + k);
  for (int k = lane; k < m.latent_tex; k += 32) inv[d.bott + d.dir_dim + k] = __ldg(m.latent + m.latent_shape + k);
  __syncthreads();

  const int total = N * S;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int pt = tile * kThreads + threadIdx.x;
    bool hit = false;
    float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    if (pt < total) hit = box_point(tdist, origins, directions, viewdirs, pose, n_obj, track, S, pt, px, py, pz, vx, vy, vz);
    unsigned hits = __ballot_sync(NLB_FULL_MASK, hit);
    while (hits) {
      const int src = __ffs(hits) - 1;
      hits &= hits - 1;
      const float qx = __shfl_sync(NLB_FULL_MASK, px, src), qy = __shfl_sync(NLB_FULL_MASK, py, src),
                  qz = __shfl_sync(NLB_FULL_MASK, pz, src);
      const float ux = __shfl_sync(NLB_FULL_MASK, vx, src), uy = __shfl_sync(NLB_FULL_MASK, vy, src),
                  uz = __shfl_sync(NLB_FULL_MASK, vz, src);
      const int q = __shfl_sync(NLB_FULL_MASK, pt, src);
      // GridEncoder.forward(bound=1): (x + 1) / 2, one level per lane
      if (lane < tab.L)
        level_features<C>(tab, lane, __fmul_rn(__fadd_rn(qx, 1.f), 0.5f), __fmul_rn(__fadd_rn(qy, 1.f), 0.5f),
                          __fmul_rn(__fadd_rn(qz, 1.f), 0.5f), in0);
      __syncwarp();
      dense<true>(W + d.o_wd0, W + d.o_bd0, in0, d.K0, d.hidden, h, lane);
      __syncwarp();
      dense<false>(W + d.o_wd2, W + d.o_bd2, h, d.hidden, d.bott, inv, lane);   // x = bottleneck, in place in inv
      if (want_rgb && lane < 3) {
        // coord.pos_enc(viewdirs_o, 0, deg_view, append_identity=True): [v, sin(2^s v) (scale-major), sin(2^s v + pi/2)]
        const float v = lane == 0 ? ux : (lane == 1 ? uy : uz);
        float* de = inv + d.bott;
        de[lane] = v;
        for (int s = 0; s < m.deg_view; ++s) {
          const float a = __fmul_rn(v, (float)(1 << s));
          de[3 + 3 * s + lane] = sinf(a);
          de[3 + 3 * m.deg_view + 3 * s + lane] = sinf(__fadd_rn(a, 1.5707963267948966f));
        }
      }
      __syncwarp();
      if (lane == 0) {
        const float xin = __fadd_rn(inv[0], m.density_bias);
        density[q] = xin > 20.f ? xin : log1pf(expf(xin));   // F.softplus (threshold 20)
        obj_mask[q] = 1;
        if (owner) owner[q] = track;   // the LAST track that hits a sample owns it (and its gradient)
      }
      if (semantic) {
        // fixed_semantic: one-hot of the class, all zeros for class 255 (models.py:1128-1133)
        for (int k = lane; k < m.class_num; k += 32) semantic[(size_t)q * m.class_num + k] = (k == m.class_type) ? 1.f : 0.f;
      }
      if (want_rgb) {
        dense<true>(W + d.o_wv0, W + d.o_bv0, inv, d.Kv, d.vw, h1, lane);
        __syncwarp();
        dense<true>(W + d.o_wv1, W + d.o_bv1, h1, d.vw + d.Kv, d.vw, h2, lane);
        __syncwarp();
        if (lane < 3) {
          float acc = W[d.o_brgb + lane];
          for (int k = 0; k < d.vw; ++k) acc = fmaf(W[d.o_wrgb + lane * d.vw + k], h2[k], acc);
          const float sg = 1.0f / (1.0f + expf(-(__fadd_rn(__fmul_rn(m.rgb_premultiplier, acc), m.rgb_bias))));
          rgb[(size_t)q * 3 + lane] = __fsub_rn(__fmul_rn(sg, 1.0f + 2.0f * m.rgb_padding), m.rgb_padding);
        }
      }
      __syncwarp();
    }
  }
}


// ----------------------------------------------------------------------------- backward (final level, training)
// Gradients of the ObjMLP weights, the track's latent code and the object table for the samples this track OWNS
// (owner[pt] == track: later tracks overwrite earlier ones in the forward).  One warp per block: the forward is
// recomputed per hit (nothing was saved), the chain rule runs with the lanes splitting the units of every layer, and
// the weight gradients accumulate in a shared-memory image of the weights that is flushed once per block.  Poses are
// constants (track refinement is outside this library), so nothing flows to the points themselves.
//   out = act(b + W in):  gW[j][k] += dz[j] in[k],  gb[j] += dz[j],  din[k] = sum_j W[j][k] dz[j]
__device__ __forceinline__ void dense_bwd(const float* __restrict__ Wt, float* __restrict__ gWt, float* __restrict__ gb,
                                          const float* __restrict__ in, const float* __restrict__ dz, int K, int J, int ld,
                                          float* __restrict__ din, int lane) {
  for (int j = lane; j < J; j += 32) {
    const float d = dz[j];
    gb[j] += d;
    for (int k = 0; k < K; ++k) gWt[k * ld + j] = fmaf(d, in[k], gWt[k * ld + j]);
  }
  if (din) {
    for (int k = lane; k < K; k += 32) {   // ld = J + 1: lane k reads bank (k + j) mod 32
      float acc = 0.f;
      for (int j = 0; j < J; ++j) acc = fmaf(Wt[k * ld + j], dz[j], acc);
      din[k] = acc;
    }
  }
}

template <bool kRelu>
__device__ __forceinline__ void dense_ld(const float* __restrict__ Wt, const float* __restrict__ b, const float* __restrict__ in,
                                         int K, int J, int ld, float* __restrict__ out, int lane) {
  for (int j = lane; j < J; j += 32) {
    float acc = b[j];
    for (int k = 0; k < K; ++k) acc = fmaf(Wt[k * ld + j], in[k], acc);
    out[j] = kRelu ? fmaxf(acc, 0.f) : acc;
  }
}

struct BDims {
  int F, K0, hidden, bott, dir_dim, Kv, vw;
  int o_wd0, o_bd0, o_wd2, o_bd2, o_wv0, o_bv0, o_wv1, o_bv1, o_wrgb, o_brgb, o_lat, total;   // padded strides J + 1
};
__host__ __device__ inline BDims make_bdims(const nlb_obj_mlp_t& m, int F) {
  BDims d;
  d.F = F; d.K0 = F + m.latent_shape; d.hidden = m.hidden; d.bott = m.bottleneck;
  d.dir_dim = 3 + 6 * m.deg_view; d.Kv = m.bottleneck + d.dir_dim + m.latent_tex; d.vw = m.view_width;
  int o = 0;
  d.o_wd0 = o; o += d.K0 * (d.hidden + 1);
  d.o_bd0 = o; o += d.hidden;
  d.o_wd2 = o; o += d.hidden * (d.bott + 1);
  d.o_bd2 = o; o += d.bott;
  d.o_wv0 = o; o += d.Kv * (d.vw + 1);
  d.o_bv0 = o; o += d.vw;
  d.o_wv1 = o; o += (d.vw + d.Kv) * (d.vw + 1);
  d.o_bv1 = o; o += d.vw;
  d.o_wrgb = o; o += d.vw * 3;
  d.o_brgb = o; o += 4;
  d.o_lat = o; o += m.latent_shape + m.latent_tex;
  d.total = (o + 3) & ~3;
  return d;
}

__device__ __forceinline__ void stage_transposed_ld(float* __restrict__ dst, const float* __restrict__ W, int J, int K, int ld) {
  for (int e = threadIdx.x; e < J * K; e += 32) {
    const int j = e / K, k = e % K;
    dst[k * ld + j] = __ldg(W + e);
  }
}
__device__ __forceinline__ void flush_transposed_ld(const float* __restrict__ src, float* __restrict__ gW, int J, int K, int ld) {
  for (int e = threadIdx.x; e < J * K; e += 32) {
    const int j = e / K, k = e % K;
    const float v = src[k * ld + j];
    if (v != 0.f) atomicAdd(gW + e, v);
  }
}

template <int C>
__global__ void __launch_bounds__(32) k_obj_backward(const float* __restrict__ tdist, const float* __restrict__ origins,
                                                     const float* __restrict__ directions, const float* __restrict__ viewdirs,
                                                     const float* __restrict__ pose, int n_obj, int track, int N, int S,
                                                     nlb_table_t tab, nlb_obj_mlp_t m, const int32_t* __restrict__ owner,
                                                     const float* __restrict__ g_density, const float* __restrict__ g_rgb,
                                                     nlb_obj_grads_t g, int num_tiles) {
  extern __shared__ __align__(16) float smem[];
  const BDims d = make_bdims(m, tab.L * C);
  float* W = smem;                   // weights, transposed, stride J + 1
  float* G = smem + d.total;         // their gradients, same layout (+ the latent's)
  float* sc = G + d.total;
  const int lane = threadIdx.x;
  float* in0 = sc;                   // [K0]
  float* h = in0 + d.K0;             // [hidden]
  float* h1 = h + d.hidden;          // [vw] | inv [Kv]  (contiguous: in2 = cat[h1, inv])
  float* inv = h1 + d.vw;
  float* h2 = inv + d.Kv;            // [vw]
  const int mx = max(d.hidden, max(d.bott, d.vw)), nd = max(d.vw + d.Kv, max(d.K0, d.hidden));
  float* dz = h2 + d.vw;             // [mx]  pre-activation gradient of the current layer
  float* din = dz + mx;              // [nd]  input gradient of the current layer
  float* dx = din + nd;              // [bott]
  for (int e = lane; e < d.total; e += 32) G[e] = 0.f;
  stage_transposed_ld(W + d.o_wd0, m.W_d0, d.hidden, d.K0, d.hidden + 1);
  stage_transposed_ld(W + d.o_wd2, m.W_d2, d.bott, d.hidden, d.bott + 1);
  stage_transposed_ld(W + d.o_wv0, m.W_v0, d.vw, d.Kv, d.vw + 1);
  stage_transposed_ld(W + d.o_wv1, m.W_v1, d.vw, d.vw + d.Kv, d.vw + 1);
  for (int e = lane; e < d.hidden; e += 32) W[d.o_bd0 + e] = __ldg(m.b_d0 + e);
  for (int e = lane; e < d.bott; e += 32) W[d.o_bd2 + e] = __ldg(m.b_d2 + e);
  for (int e = lane; e < d.vw; e += 32) { W[d.o_bv0 + e] = __ldg(m.b_v0 + e); W[d.o_bv1 + e] = __ldg(m.b_v1 + e); }
  for (int e = lane; e < 3 * d.vw; e += 32) W[d.o_wrgb + e] = __ldg(m.W_rgb + e);
  if (lane < 3) W[d.o_brgb + lane] = __ldg(m.b_rgb + lane);
  for (int k = lane; k < m.latent_shape; k += 32) in0[d.F + k] = __ldg(m.latent + k);
  for (int k = lane; k < m.latent_tex; k += 32) inv[d.bott + d.dir_dim + k] = __ldg(m.latent + m.latent_shape + k);
  __syncwarp();

  const int total = N * S;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int pt = tile * 32 + lane;
    float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    bool hit = false;
    if (pt < total && owner[pt] == track)
      hit = box_point(tdist, origins, directions, viewdirs, pose, n_obj, track, S, pt, px, py, pz, vx, vy, vz);
    unsigned hits = __ballot_sync(NLB_FULL_MASK, hit);
    while (hits) {
      const int src = __ffs(hits) - 1;
      hits &= hits - 1;
      const float qx = __shfl_sync(NLB_FULL_MASK, px, src), qy = __shfl_sync(NLB_FULL_MASK, py, src),
                  qz = __shfl_sync(NLB_FULL_MASK, pz, src);
      const float ux = __shfl_sync(NLB_FULL_MASK, vx, src), uy = __shfl_sync(NLB_FULL_MASK, vy, src),
                  uz = __shfl_sync(NLB_FULL_MASK, vz, src);
      const int q = __shfl_sync(NLB_FULL_MASK, pt, src);
      const float gx = __fmul_rn(__fadd_rn(qx, 1.f), 0.5f), gy = __fmul_rn(__fadd_rn(qy, 1.f), 0.5f),
                  gz = __fmul_rn(__fadd_rn(qz, 1.f), 0.5f);
      // ---- forward, recomputed
      if (lane < tab.L) level_features<C>(tab, lane, gx, gy, gz, in0);
      __syncwarp();
      dense_ld<true>(W + d.o_wd0, W + d.o_bd0, in0, d.K0, d.hidden, d.hidden + 1, h, lane);
      __syncwarp();
      dense_ld<false>(W + d.o_wd2, W + d.o_bd2, h, d.hidden, d.bott, d.bott + 1, inv, lane);
      if (lane < 3) {
        const float v = lane == 0 ? ux : (lane == 1 ? uy : uz);
        float* de = inv + d.bott;
        de[lane] = v;
        for (int s = 0; s < m.deg_view; ++s) {
          const float a = __fmul_rn(v, (float)(1 << s));
          de[3 + 3 * s + lane] = sinf(a);
          de[3 + 3 * m.deg_view + 3 * s + lane] = sinf(__fadd_rn(a, 1.5707963267948966f));
        }
      }
      __syncwarp();
      dense_ld<true>(W + d.o_wv0, W + d.o_bv0, inv, d.Kv, d.vw, d.vw + 1, h1, lane);
      __syncwarp();
      dense_ld<true>(W + d.o_wv1, W + d.o_bv1, h1, d.vw + d.Kv, d.vw, d.vw + 1, h2, lane);
      __syncwarp();
      // ---- rgb head: d(rgb)/d(pre) = premult * (1 + 2 pad) * s (1 - s)
      float d3 = 0.f;
      if (lane < 3 && g_rgb) {
        float acc = W[d.o_brgb + lane];
        for (int k = 0; k < d.vw; ++k) acc = fmaf(W[d.o_wrgb + lane * d.vw + k], h2[k], acc);
        const float sg = 1.0f / (1.0f + expf(-(__fadd_rn(__fmul_rn(m.rgb_premultiplier, acc), m.rgb_bias))));
        d3 = __ldg(g_rgb + (size_t)q * 3 + lane) * (1.0f + 2.0f * m.rgb_padding) * sg * (1.0f - sg) * m.rgb_premultiplier;
        G[d.o_brgb + lane] += d3;
      }
      const float d30 = __shfl_sync(NLB_FULL_MASK, d3, 0), d31 = __shfl_sync(NLB_FULL_MASK, d3, 1),
                  d32 = __shfl_sync(NLB_FULL_MASK, d3, 2);
      for (int k = lane; k < d.vw; k += 32) {
        const float hk = h2[k];
        G[d.o_wrgb + k] = fmaf(d30, hk, G[d.o_wrgb + k]);
        G[d.o_wrgb + d.vw + k] = fmaf(d31, hk, G[d.o_wrgb + d.vw + k]);
        G[d.o_wrgb + 2 * d.vw + k] = fmaf(d32, hk, G[d.o_wrgb + 2 * d.vw + k]);
        const float dh = W[d.o_wrgb + k] * d30 + W[d.o_wrgb + d.vw + k] * d31 + W[d.o_wrgb + 2 * d.vw + k] * d32;
        dz[k] = hk > 0.f ? dh : 0.f;
      }
      __syncwarp();
      // ---- lin_second_stage_1: in2 = [h1 | inv]
      dense_bwd(W + d.o_wv1, G + d.o_wv1, G + d.o_bv1, h1, dz, d.vw + d.Kv, d.vw, d.vw + 1, din, lane);
      __syncwarp();
      for (int k = lane; k < d.vw; k += 32) dz[k] = h1[k] > 0.f ? din[k] : 0.f;       // through the ReLU of h1
      __syncwarp();
      // ---- lin_second_stage_0 (its input gradient lands after the skip part of din: d_inv = din[vw:] + ...)
      float* dinv = din + d.vw;
      for (int j = lane; j < d.vw; j += 32) {
        const float dj = dz[j];
        G[d.o_bv0 + j] += dj;
        for (int k = 0; k < d.Kv; ++k) G[d.o_wv0 + k * (d.vw + 1) + j] = fmaf(dj, inv[k], G[d.o_wv0 + k * (d.vw + 1) + j]);
      }
      for (int k = lane; k < d.Kv; k += 32) {
        float acc = dinv[k];
        for (int j = 0; j < d.vw; ++j) acc = fmaf(W[d.o_wv0 + k * (d.vw + 1) + j], dz[j], acc);
        dinv[k] = acc;
      }
      __syncwarp();
      // ---- unit view direction (track refinement): through coord.pos_enc, d sin(a + pi/2) = -sin(a)
      float gu_l = 0.f;
      if (g.g_pose && lane < 3) {
        const float v = lane == 0 ? ux : (lane == 1 ? uy : uz);
        const float* ge = dinv + d.bott;
        gu_l = ge[lane];
        for (int s = 0; s < m.deg_view; ++s) {
          const float sc2 = (float)(1 << s), a = v * sc2;
          gu_l += sc2 * (cosf(a) * ge[3 + 3 * s + lane] - sinf(a) * ge[3 + 3 * m.deg_view + 3 * s + lane]);
        }
      }
      // ---- bottleneck: dx = d_inv[:bott] (+ the density path on unit 0), texture latent
      for (int k = lane; k < d.bott; k += 32) dx[k] = dinv[k];
      for (int k = lane; k < m.latent_tex; k += 32) G[d.o_lat + m.latent_shape + k] += dinv[d.bott + d.dir_dim + k];
      __syncwarp();
      if (lane == 0 && g_density) {
        const float xin = __fadd_rn(inv[0], m.density_bias);
        dx[0] += __ldg(g_density + q) * (1.0f / (1.0f + expf(-xin)));     // softplus'
      }
      __syncwarp();
      // ---- density_layer.2 (no activation) and density_layer.0 (ReLU)
      dense_bwd(W + d.o_wd2, G + d.o_wd2, G + d.o_bd2, h, dx, d.hidden, d.bott, d.bott + 1, din, lane);
      __syncwarp();
      for (int k = lane; k < d.hidden; k += 32) dz[k] = h[k] > 0.f ? din[k] : 0.f;
      __syncwarp();
      dense_bwd(W + d.o_wd0, G + d.o_wd0, G + d.o_bd0, in0, dz, d.K0, d.hidden, d.hidden + 1, din, lane);
      __syncwarp();
      for (int k = lane; k < m.latent_shape; k += 32) G[d.o_lat + k] += din[d.F + k];
      // ---- grid: scatter the feature gradient to the 8 corners of every level (one level per lane); with track
      // refinement also the trilinear slope of the level (the reference's dy_dx) for dL/d(grid coordinates)
      float sl[3] = {0.f, 0.f, 0.f};
      if (lane < tab.L && (g.g_table || g.g_pose)) {
        const Level3 lv = level3(tab.offsets, lane, tab.S, tab.H);
        uint32_t cx, cy, cz;
        float fx, fy, fz;
        cell_of(gx, lv.scale, cx, fx);
        cell_of(gy, lv.scale, cy, fy);
        cell_of(gz, lv.scale, cz, fz);
        uint32_t idx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) idx[i] = vertex_index3_branchy(lv, cx + (i & 1), cy + ((i >> 1) & 1), cz + ((i >> 2) & 1));
        if (g.g_table) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float w = 1.f;
            w *= (i & 1) ? fx : 1.f - fx;
            w *= (i & 2) ? fy : 1.f - fy;
            w *= (i & 4) ? fz : 1.f - fz;
#pragma unroll
            for (int c = 0; c < C; ++c) atomicAdd(g.g_table + ((size_t)lv.offset + idx[i]) * C + c, w * din[lane * C + c]);
          }
        }
        if (g.g_pose) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __ldg(tab.embeddings + ((size_t)lv.offset + idx[i]) * C + c);
            const float x00 = r[1] - r[0], x10 = r[3] - r[2], x01 = r[5] - r[4], x11 = r[7] - r[6];
            const float v00 = fmaf(fx, x00, r[0]), v10 = fmaf(fx, x10, r[2]), v01 = fmaf(fx, x01, r[4]), v11 = fmaf(fx, x11, r[6]);
            const float y0 = v10 - v00, y1 = v11 - v01;
            const float e0 = fmaf(fy, y0, v00), e1 = fmaf(fy, y1, v01);
            const float dx0 = fmaf(fy, x10 - x00, x00), dx1 = fmaf(fy, x11 - x01, x01);
            const float gf = din[lane * C + c] * lv.scale;
            sl[0] = fmaf(gf, fmaf(fz, dx1 - dx0, dx0), sl[0]);
            sl[1] = fmaf(gf, fmaf(fz, y1 - y0, y0), sl[1]);
            sl[2] = fmaf(gf, e1 - e0, sl[2]);
          }
        }
      }
      if (g.g_pose) {
        float gq[3], gu[3], gp[7];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          gq[k] = 0.5f * warp_sum(sl[k]);                    // grid coordinate = (box coordinate + 1) / 2
          gu[k] = __shfl_sync(NLB_FULL_MASK, gu_l, k);
        }
        box_point_grad(tdist, origins, directions, viewdirs, pose, n_obj, track, S, q, gq, gu, gp);
        if (lane < 7) {
          float v = gp[0];
#pragma unroll
          for (int k = 1; k < 7; ++k) v = lane == k ? gp[k] : v;
          atomicAdd(g.g_pose + ((size_t)(q / S) * n_obj + track) * kInfo + lane, v);
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // ---- flush the block's gradients
  flush_transposed_ld(G + d.o_wd0, g.g_W_d0, d.hidden, d.K0, d.hidden + 1);
  flush_transposed_ld(G + d.o_wd2, g.g_W_d2, d.bott, d.hidden, d.bott + 1);
  flush_transposed_ld(G + d.o_wv0, g.g_W_v0, d.vw, d.Kv, d.vw + 1);
  flush_transposed_ld(G + d.o_wv1, g.g_W_v1, d.vw, d.vw + d.Kv, d.vw + 1);
  for (int e = lane; e < d.hidden; e += 32) if (G[d.o_bd0 + e] != 0.f) atomicAdd(g.g_b_d0 + e, G[d.o_bd0 + e]);
  for (int e = lane; e < d.bott; e += 32) if (G[d.o_bd2 + e] != 0.f) atomicAdd(g.g_b_d2 + e, G[d.o_bd2 + e]);
  for (int e = lane; e < d.vw; e += 32) {
    if (G[d.o_bv0 + e] != 0.f) atomicAdd(g.g_b_v0 + e, G[d.o_bv0 + e]);
    if (G[d.o_bv1 + e] != 0.f) atomicAdd(g.g_b_v1 + e, G[d.o_bv1 + e]);
  }
  for (int e = lane; e < 3 * d.vw; e += 32) if (G[d.o_wrgb + e] != 0.f) atomicAdd(g.g_W_rgb + e, G[d.o_wrgb + e]);
  if (lane < 3 && G[d.o_brgb + lane] != 0.f) atomicAdd(g.g_b_rgb + lane, G[d.o_brgb + lane]);
  if (g.g_latent)
    for (int e = lane; e < m.latent_shape + m.latent_tex; e += 32) if (G[d.o_lat + e] != 0.f) atomicAdd(g.g_latent + e, G[d.o_lat + e]);
}

}  // namespace obj
}  // namespace nlb

using namespace nlb;

extern "C" int nlb_obj_pose(const float* time, const float* tracks, int N, int n_obj, int T, float* pose, void* stream) {
  if (N == 0 || n_obj == 0) return NLB_OK;
  if (N < 0 || n_obj < 0 || !time || !tracks || !pose) { nlb_set_error("obj_pose: null pointer or negative size"); return NLB_EINVAL; }
  if (T < 2) { nlb_set_error("obj_pose: a track needs at least two timestamps (get_pose takes the two closest)"); return NLB_EINVAL; }
  const int total = N * n_obj;
  obj::k_obj_pose<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(time, tracks, N, n_obj, T, pose);
  return nlb_check_launch("obj_pose");
}

extern "C" int nlb_obj_forward(const float* tdist, const float* origins, const float* directions, const float* viewdirs,
                               const float* pose, int n_obj, int track, int N, int S, const nlb_table_t* table,
                               const nlb_obj_mlp_t* mlp, float* density, float* rgb, float* semantic, uint8_t* obj_mask,
                               int32_t* owner, void* stream) {
  if (N == 0 || S == 0) return NLB_OK;
  if (N < 0 || S < 0 || !tdist || !origins || !directions || !viewdirs || !pose || !table || !mlp || !density || !obj_mask) {
    nlb_set_error("obj_forward: null pointer or negative size");
    return NLB_EINVAL;
  }
  if (track < 0 || track >= n_obj) { nlb_set_error("obj_forward: track %d outside [0, %d)", track, n_obj); return NLB_EINVAL; }
  if (!table->embeddings || !table->offsets || table->L < 1 || table->L > 32) {
    nlb_set_error("obj_forward: bad table (1 <= L <= 32: one grid level per lane)");
    return NLB_EINVAL;
  }
  const nlb_obj_mlp_t& m = *mlp;
  if (!m.W_d0 || !m.b_d0 || !m.W_d2 || !m.b_d2 || (rgb && (!m.W_v0 || !m.b_v0 || !m.W_v1 || !m.b_v1 || !m.W_rgb || !m.b_rgb))) {
    nlb_set_error("obj_forward: null weight pointer");
    return NLB_EINVAL;
  }
  if ((m.latent_shape > 0 || m.latent_tex > 0) && !m.latent) { nlb_set_error("obj_forward: latent sizes without a latent vector"); return NLB_EINVAL; }
  if (m.hidden < 1 || m.bottleneck < 1 || m.view_width < 1 || m.deg_view < 0 || m.latent_shape < 0 || m.latent_tex < 0 ||
      (semantic && (m.class_num < 1))) {
    nlb_set_error("obj_forward: bad layer sizes");
    return NLB_EINVAL;
  }
  const obj::Dims d = obj::make_dims(m, table->L * table->C);
  const size_t smem = (size_t)(d.total + obj::kWarps * d.scratch) * sizeof(float);
  if (smem > 200 * 1024) { nlb_set_error("obj_forward: the ObjMLP weights (%zu bytes) do not fit shared memory", smem); return NLB_EUNSUPPORTED; }
  const int tiles = (int)(((int64_t)N * S + obj::kThreads - 1) / obj::kThreads);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
  cudaStream_t st = (cudaStream_t)stream;
#define NLB_OBJ_LAUNCH(C_)                                                                                              \
  {                                                                                                                     \
    cudaFuncSetAttribute(obj::k_obj_forward<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    obj::k_obj_forward<C_><<<grid, obj::kThreads, smem, st>>>(tdist, origins, directions, viewdirs, pose, n_obj, track, \
                                                              N, S, *table, m, density, rgb, semantic, obj_mask, owner, tiles); \
  }
  switch (table->C) {
    case 1: NLB_OBJ_LAUNCH(1) break;
    case 2: NLB_OBJ_LAUNCH(2) break;
    case 4: NLB_OBJ_LAUNCH(4) break;
    case 8: NLB_OBJ_LAUNCH(8) break;
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
#undef NLB_OBJ_LAUNCH
  return nlb_check_launch("obj_forward");
}

extern "C" int nlb_obj_backward(const float* tdist, const float* origins, const float* directions, const float* viewdirs,
                                const float* pose, int n_obj, int track, int N, int S, const nlb_table_t* table,
                                const nlb_obj_mlp_t* mlp, const int32_t* owner, const float* g_density, const float* g_rgb,
                                const nlb_obj_grads_t* grads, void* stream) {
  if (N == 0 || S == 0) return NLB_OK;
  if (N < 0 || S < 0 || !tdist || !origins || !directions || !viewdirs || !pose || !table || !mlp || !owner || !grads) {
    nlb_set_error("obj_backward: null pointer or negative size");
    return NLB_EINVAL;
  }
  if (track < 0 || track >= n_obj) { nlb_set_error("obj_backward: track %d outside [0, %d)", track, n_obj); return NLB_EINVAL; }
  const nlb_obj_mlp_t& m = *mlp;
  const nlb_obj_grads_t& g = *grads;
  if (!m.W_d0 || !m.b_d0 || !m.W_d2 || !m.b_d2 || !m.W_v0 || !m.b_v0 || !m.W_v1 || !m.b_v1 || !m.W_rgb || !m.b_rgb ||
      !g.g_W_d0 || !g.g_b_d0 || !g.g_W_d2 || !g.g_b_d2 || !g.g_W_v0 || !g.g_b_v0 || !g.g_W_v1 || !g.g_b_v1 || !g.g_W_rgb || !g.g_b_rgb) {
    nlb_set_error("obj_backward: null weight / gradient pointer");
    return NLB_EINVAL;
  }
  if ((m.latent_shape > 0 || m.latent_tex > 0) && !m.latent) { nlb_set_error("obj_backward: latent sizes without a latent vector"); return NLB_EINVAL; }
  if (!table->embeddings || !table->offsets || table->L < 1 || table->L > 32) { nlb_set_error("obj_backward: bad table"); return NLB_EINVAL; }
  const obj::BDims d = obj::make_bdims(m, table->L * table->C);
  const int mx = std::max(d.hidden, std::max(d.bott, d.vw)), nd = std::max(d.vw + d.Kv, std::max(d.K0, d.hidden));
  const size_t scratch = (size_t)d.K0 + d.hidden + d.vw + d.Kv + d.vw + mx + nd + d.bott + 8;
  const size_t smem = ((size_t)2 * d.total + scratch) * sizeof(float);
  if (smem > 220 * 1024) { nlb_set_error("obj_backward: the ObjMLP (%zu bytes of weights + gradients) does not fit shared memory", smem); return NLB_EUNSUPPORTED; }
  const int tiles = (int)(((int64_t)N * S + 31) / 32);
  const int sms = nlb_sm_count();
  const int grid = tiles < sms ? tiles : sms;
  cudaStream_t st = (cudaStream_t)stream;
#define NLB_OBJ_BWD(C_)                                                                                                \
  {                                                                                                                    \
    cudaFuncSetAttribute(obj::k_obj_backward<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    obj::k_obj_backward<C_><<<grid, 32, smem, st>>>(tdist, origins, directions, viewdirs, pose, n_obj, track, N, S,     \
                                                    *table, m, owner, g_density, g_rgb, g, tiles);                     \
  }
  switch (table->C) {
    case 1: NLB_OBJ_BWD(1) break;
    case 2: NLB_OBJ_BWD(2) break;
    case 4: NLB_OBJ_BWD(4) break;
    case 8: NLB_OBJ_BWD(8) break;
    default: nlb_set_error("GridEncoding: C must be 1, 2, 4, or 8."); return NLB_EINVAL;
  }
#undef NLB_OBJ_BWD
  return nlb_check_launch("obj_backward");
}
