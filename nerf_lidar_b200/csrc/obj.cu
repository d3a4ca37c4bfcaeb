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
                                                         uint8_t* __restrict__ obj_mask, int num_tiles) {
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
    if (pt < total) {
      const int ray = pt / S, s = pt - ray * S;
      const float* td = tdist + (size_t)ray * (S + 1) + s;
      const float tm = __fmul_rn(0.5f, __fadd_rn(__ldg(td), __ldg(td + 1)));
      // pts_w = t_mids * directions + origins (models.py:404)
      px = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray)), __ldg(origins + 3 * ray));
      py = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray + 1)), __ldg(origins + 3 * ray + 1));
      pz = __fadd_rn(__fmul_rn(tm, __ldg(directions + 3 * ray + 2)), __ldg(origins + 3 * ray + 2));
      const float* ps = pose + ((size_t)ray * n_obj + track) * kInfo;
      const float theta = __ldg(ps + 3);
      const float c = cosf(theta), sn = sinf(theta);
      // t_w_o = rotate_yaw_z(-centre, yaw); pts_o = rotate_yaw_z(pts_w, yaw) + t_w_o (obj_utils.py:158-170)
      float tx = -__ldg(ps), ty = -__ldg(ps + 1);
      const float tz = -__ldg(ps + 2);
      rotate_yaw_z(tx, ty, c, sn);
      rotate_yaw_z(px, py, c, sn);
      px = __fadd_rn(px, tx); py = __fadd_rn(py, ty); pz = __fadd_rn(pz, tz);
      vx = __ldg(viewdirs + 3 * ray); vy = __ldg(viewdirs + 3 * ray + 1); vz = __ldg(viewdirs + 3 * ray + 2);
      rotate_yaw_z(vx, vy, c, sn);
      // scale_frames: 1 / (wlh / 2 + 1e-9) per axis (obj_utils.py:17-25), then the directions are re-normalised
      const float sx = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 4), 0.5f), 1e-9f));
      const float sy = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 5), 0.5f), 1e-9f));
      const float sz = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__ldg(ps + 6), 0.5f), 1e-9f));
      px = __fmul_rn(sx, px); py = __fmul_rn(sy, py); pz = __fmul_rn(sz, pz);
      vx = __fmul_rn(sx, vx); vy = __fmul_rn(sy, vy); vz = __fmul_rn(sz, vz);
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
      vx = __fdiv_rn(vx, nrm); vy = __fdiv_rn(vy, nrm); vz = __fdiv_rn(vz, nrm);
      hit = fabsf(px) < 1.f && fabsf(py) < 1.f && fabsf(pz) < 1.f;   // box_pts (obj_utils.py:205)
    }
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
                               void* stream) {
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
                                                              N, S, *table, m, density, rgb, semantic, obj_mask, tiles); \
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
