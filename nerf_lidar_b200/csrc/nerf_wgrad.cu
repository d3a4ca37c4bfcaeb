// Weight gradients of the NerfMLP dense layers on tcgen05 (autograd of Z/internal/models.py:1192-1251):
//   dW = dZ^T A  for every layer, dZ = the bf16 pre-activation gradients written by k_nerf_mlp_bwd, A = the bf16
//   activations saved by k_nerf_mlp_fwd.  The contraction index is the SAMPLE index m, so both operands are read
//   from their row-major [M, features] matrices exactly as they lie: a [32 rows][64 features] slab with 128-byte
//   rows and the 128-byte swizzle IS the canonical MN-major UMMA operand layout (rows = K, LBO = next 64 features,
//   SBO = next 8 rows; validated by tools/umma_mn_probe.cu) -- no transposes, no packing.
//
// The accumulators (253 K fp32 values) do not fit one SM's tensor memory (64 K), so the products are split into
// six ROLES; a role's CTAs share the M rows between them (split-K), stream their 32-row slabs through a 6-stage
// cp.async ring, keep the role's accumulators in TMEM for their whole row range and add them into the fp32
// gradient tensors at the end (coalesced RED.ADD.F32 -- straight into Trainer.flat_grad, no partial buffers, no
// AccumulateGrad adds).  The kernel is HBM-bound: every operand byte is read once per role that needs it, 5.6 KB
// per row in total (x three times, d_v1 twice; 1.83 GB for the bench's 327 680 rows, 1.49 GB of DRAM reads after L2
// hits) against ~270 tensor cycles per 32 KB stage.  Measured 0.39 ms = 4.7 TB/s algorithmic (cuBLAS: seven GEMMs +
// split-K reductions, 0.40-0.45 ms, plus the bf16 copy of the features and ~40 AccumulateGrad adds).
//
//   role 0: d_v0^T x   -> W_v0[:, 0:256]        role 3: d_g^T  x   -> W_s0 | W_i0
//   role 1: d_v1^T x   -> W_v1[:, 256:512]      role 4: d_x^T  h0  -> W_d2 ;  d_h0^T f0 -> W_d0
//   role 2: d_v1^T h1  -> W_v1[:, 0:256]        role 5: g^T d_hs1  -> W_s2^T | W_i2^T ;  h2^T d_rgb -> W_rgb^T
//
// Bias gradients and the view-direction columns (the direction encoding is a per-ray constant) come from the
// column / per-ray sums of csrc/reduce.cu, folded in by k_wgrad_finish below.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/nlb200.h"
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>

namespace nlb {
namespace wgrad {

using namespace nlb::umma;

constexpr int kRows = 32;                 // sample rows per pipeline stage (2 K-steps of 16)
constexpr int kBlk = kRows * 128;         // one [32 rows][64 features] bf16 slab
constexpr int kStageBytes = 8 * kBlk;     // every role fills (at most) eight slabs per stage
// Deep ring of small stages: a slot is refilled only after the MMAs that read it have retired, so with S slots about
// S - 2 stages of loads are in flight; HBM latency x the per-SM share of the bandwidth needs > 100 KB in flight.
constexpr int kStages = 7;
// A stage is announced to the MMA warp once the copies of the kInFlight - 1 stages issued after it are in flight
// (cp.async groups complete in order); the remaining kStages - kInFlight + 1 slots are the slack that lets the
// producers run ahead of the tensor pipe instead of shaking hands with it every stage.
constexpr int kInFlight = 6;
// 8 producer warps: with 256 threads every stream's (row, 16-byte chunk) pattern repeats with a row period that is
// a multiple of 8, so a thread's swizzle term and source column are loop constants and a copy costs ~5 instructions
// (the first version spent 39 instructions per copy on index arithmetic with 4 warps and was issue-bound at 3 TB/s)
constexpr int kProdWarps = 8, kMmaWarp = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kProdWarps + 1) * 32;
constexpr int kNumRoles = 6;
constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + 256;

struct Stream {                 // one operand matrix of a role: rows [r0, r0+64) x `cols` columns -> slabs at smem_off
  const __nv_bfloat16* ptr;
  int ld;                       // elements
  int log2_cpr;                 // log2(16-byte chunks per row)
  int smem_off;
};
struct Op {                     // D[128, n] (+)= A^T B on the stage's slabs, accumulated in TMEM columns [tmem_col, +n)
  int a_off, b_off, n, tmem_col;
  // where the accumulator goes: row r of D, column c ->
  //   mode 0: base[r / 64] + (r % 64) * stride + c          (D row = weight row)
  //   mode 1: base[r / 64] + c * stride + (r % 64)          (D row = weight COLUMN: transposed products)
  // for c in [col_lo[h], col_hi[h]), r < rows_valid
  float* base[2];
  int stride;
  int col_lo[2], col_hi[2];
  int rows_valid;
  int mode;
};
struct Role {
  Stream s[4];
  Op op[4];
  int ns, nop;
  int cta0, nctas;
};
struct Params {
  Role role[kNumRoles];
  int M;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint64_t desc_mn(const void* p) {     // MN-major, SWIZZLE_128B, LBO = slab, SBO = 8 rows
  const uint32_t a = smem_u32(p);
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)(kBlk >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// rows [row0, row0 + kRows) x (8 << LOG2CPR) columns of a row-major bf16 matrix -> 64-column slabs (128-byte rows,
// 16-byte chunks XOR-swizzled by row & 7).  256 threads: consecutive threads take consecutive chunks of a row.
template <int LOG2CPR>
__device__ __forceinline__ void copy_stream(uint8_t* slab0, const __nv_bfloat16* ptr, int ld, int row0, int M, int tid) {
  constexpr int CPR = 1 << LOG2CPR;
  constexpr int TOTAL = kRows * CPR;
  constexpr int ITERS = (TOTAL + kProdThreads - 1) / kProdThreads;
  constexpr int ROWS_PER_ITER = kProdThreads / CPR;          // 8, 16, 32, (64, 128: one partial iteration)
  if (TOTAL < kProdThreads && tid >= TOTAL) return;
  const int r = tid >> LOG2CPR, ch = tid & (CPR - 1);
  uint8_t* dst = slab0 + (ch >> 3) * kBlk + (r >> 3) * 1024 + (r & 7) * 128 + (((ch & 7) ^ (r & 7)) * 16);
  const __nv_bfloat16* src = ptr + (size_t)(row0 + r) * ld + ch * 8;
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const bool ok = row0 + r + k * ROWS_PER_ITER < M;
    cp_async16(dst + k * (ROWS_PER_ITER / 8) * 1024, ok ? src + (size_t)k * ROWS_PER_ITER * ld : ptr, ok);
  }
}

struct Bars {
  uint64_t full[kStages], empty[kStages], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1) k_nerf_mlp_wgrad(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Bars& bars = *reinterpret_cast<Bars*>(base + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int r = 0;
#pragma unroll
  for (int i = 1; i < kNumRoles; ++i)
    if ((int)blockIdx.x >= P.role[i].cta0) r = i;
  const Role& role = P.role[r];
  const int local = (int)blockIdx.x - role.cta0;
  const int total = (P.M + kRows - 1) / kRows;
  const int per = (total + role.nctas - 1) / role.nctas;
  const int t0 = local * per;
  const int t1 = t0 + per < total ? t0 + per : total;
  const int n_it = t1 > t0 ? t1 - t0 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars.full[i], kProdThreads); mbar_init(&bars.empty[i], 1); }
    mbar_init(&bars.done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&bars.tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp < kProdWarps) {
    // ===== producers: 16-byte async copies, rows beyond M are zero-filled; a stage is handed over when every
    // producer thread's copies of it have landed (wait_group) and are visible to the tensor pipe (proxy fence)
    const int tid = threadIdx.x;
    for (int it = 0; it < n_it + kInFlight - 1; ++it) {
      if (it < n_it) {
        const int slot = it % kStages;
        if (it >= kStages) mbar_wait_relaxed(&bars.empty[slot], ((it / kStages) & 1) ^ 1);
        uint8_t* st = base + slot * kStageBytes;
        const int row0 = (t0 + it) * kRows;
        for (int si = 0; si < role.ns; ++si) {
          const Stream& s = role.s[si];
          switch (s.log2_cpr) {
            case 5: copy_stream<5>(st + s.smem_off, s.ptr, s.ld, row0, P.M, tid); break;
            case 4: copy_stream<4>(st + s.smem_off, s.ptr, s.ld, row0, P.M, tid); break;
            case 3: copy_stream<3>(st + s.smem_off, s.ptr, s.ld, row0, P.M, tid); break;
            case 2: copy_stream<2>(st + s.smem_off, s.ptr, s.ld, row0, P.M, tid); break;
            default: copy_stream<1>(st + s.smem_off, s.ptr, s.ld, row0, P.M, tid); break;
          }
        }
      }
      cp_async_commit();
      if (it >= kInFlight - 1) {
        cp_async_wait_group<kInFlight - 1>();
        fence_proxy_async();
        mbar_arrive(&bars.full[(it - (kInFlight - 1)) % kStages]);
      }
    }
  } else if (elect_one_sync()) {
    // ===== MMA issuer
    for (int it = 0; it < n_it; ++it) {
      const int slot = it % kStages;
      mbar_wait(&bars.full[slot], (it / kStages) & 1);
      tcgen05_fence_after();
      const uint8_t* st = base + slot * kStageBytes;
      for (int o = 0; o < role.nop; ++o) {
        const Op& op = role.op[o];
        const uint32_t idesc = make_idesc_bf16(128, op.n) | (1u << 15) | (1u << 16);   // A and B MN-major
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk)
          mma_bf16_ss(tmem + op.tmem_col, desc_mn(st + op.a_off + kk * 2048), desc_mn(st + op.b_off + kk * 2048), idesc,
                      (it | kk) != 0);
      }
      mma_commit(&bars.empty[slot]);
    }
    mma_commit(&bars.done);
  }
  __syncwarp();

  // ===== epilogue (the producer warps: warps w and w + 4 own TMEM lanes 32 (w & 3) .. + 31 and alternate over the
  // 32-column groups)
  if (warp < kProdWarps && n_it > 0) {
    mbar_wait_warp(&bars.done, 0);
    tcgen05_fence_after();
    float* tile = reinterpret_cast<float*>(base) + warp * (32 * 33);     // the ring is free now
    const int q = warp & 3, par = warp >> 2;
    const int rrow = q * 32 + lane, half = q >> 1, rloc = rrow & 63;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    int grp = 0;
    for (int o = 0; o < role.nop; ++o) {
      const Op& op = role.op[o];
      float* bp = op.base[half];
      const int lo = op.col_lo[half], hi = op.col_hi[half];
      for (int c0 = 0; c0 < op.n; c0 += 32, ++grp) {
        if ((grp & 1) != par) continue;
        float v[32];
        if (op.n - c0 >= 32) {
          tmem_ld32(tl + op.tmem_col + c0, v);
        } else {
          float t[16];
          tmem_ld16(tl + op.tmem_col + c0, t);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = t[j]; v[16 + j] = 0.f; }
        }
        if (op.mode == 1) {
          // transposed product: for a fixed column the lanes' rows are consecutive addresses
          if (rrow < op.rows_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c0 + j;
              if (c >= lo && c < hi) atomicAdd(bp + (size_t)c * op.stride + rloc, v[j]);
            }
          }
        } else {
          // rows of D are weight rows: turn the 32 x 32 tile through shared memory so that a warp adds 32
          // consecutive floats of ONE weight row per instruction
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j];
          __syncwarp();
          const int c = c0 + lane;
          const bool cok = c >= lo && c < hi;
          const int rbase = (q * 32) & 63;
          for (int rr = 0; rr < 32; ++rr) {
            if (cok && q * 32 + rr < op.rows_valid) atomicAdd(bp + (size_t)(rbase + rr) * op.stride + c, tile[rr * 33 + lane]);
          }
          __syncwarp();
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// Bias gradients and the view-direction columns.  cs_* = column sums of the pre-activation gradients, rs_v0 /
// rs_v1 [N,256] = their per-ray sums (csrc/reduce.cu).  grid = (8 column chunks of 32, ray chunks, 2 matrices):
//   W_v[n, off + j] += sum_ray rs[ray, n] * dir_enc(viewdirs[ray])[j],  b_v[n] += sum_ray rs[ray, n]
// block (0, 0, 0) also adds the column sums into the remaining bias gradients.
constexpr int kDirCols = 27, kRayChunk = 64;

__device__ __forceinline__ void dir_enc27(float vx, float vy, float vz, float* d /*[28]*/) {
  // coord.pos_enc(viewdirs, 0, 4, append_identity=True), Z/internal/coord.py:199-210 -- same expressions as
  // the forward kernel's stage_dirs
  d[0] = vx; d[1] = vy; d[2] = vz;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float sc = (float)(1 << s);
    const float ax = vx * sc, ay = vy * sc, az = vz * sc;
    d[3 + s * 3] = sinf(ax); d[4 + s * 3] = sinf(ay); d[5 + s * 3] = sinf(az);
    d[15 + s * 3] = sinf(ax + 1.5707963267948966f);
    d[16 + s * 3] = sinf(ay + 1.5707963267948966f);
    d[17 + s * 3] = sinf(az + 1.5707963267948966f);
  }
  d[27] = 1.0f;   // the bias gradient rides along as a 28th "direction" column
}

struct FinishArgs {
  const float* rs[2];        // [N,256]
  float* gW[2];              // W_v0 / W_v1 gradient
  float* gb[2];
  int stride[2], off[2];     // 283 / 539, 256 / 512
  const float* viewdirs;     // [N,3]
  int N;
  const float *cs_x, *cs_g, *cs_h0, *cs_hs1, *cs_rgb;
  float *gb_d2, *gb_s0, *gb_i0, *gb_d0, *gb_s2, *gb_i2, *gb_rgb;
};

__global__ void __launch_bounds__(256) k_wgrad_finish(FinishArgs a) {
  __shared__ __align__(16) float s_de[kRayChunk][28];     // rows of 7 float4: the product loop reads them as vectors
  __shared__ float s_red[8][32][29];
  const int which = blockIdx.z;
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);      // output row (unit of the 256-wide layer)
  const int rl = threadIdx.x >> 5;                          // ray lane 0..7
  const int rays_per_block = (a.N + gridDim.y - 1) / gridDim.y;
  const int ray_lo = blockIdx.y * rays_per_block;
  const int ray_hi = ray_lo + rays_per_block < a.N ? ray_lo + rays_per_block : a.N;
  float acc[28];
#pragma unroll
  for (int j = 0; j < 28; ++j) acc[j] = 0.f;
  const float* rs = a.rs[which];
  for (int c0 = ray_lo; c0 < ray_hi; c0 += kRayChunk) {
    const int cn = ray_hi - c0 < kRayChunk ? ray_hi - c0 : kRayChunk;
    __syncthreads();
    if (threadIdx.x < cn) {
      const int ray = c0 + threadIdx.x;
      float d[28];
      dir_enc27(__ldg(a.viewdirs + 3 * ray), __ldg(a.viewdirs + 3 * ray + 1), __ldg(a.viewdirs + 3 * ray + 2), d);
#pragma unroll
      for (int j = 0; j < 28; ++j) s_de[threadIdx.x][j] = d[j];
    }
    __syncthreads();
#pragma unroll 4   // (one load in flight per thread left the kernel latency-bound: 48 us for 21 MB)
    for (int i = rl; i < cn; i += 8) {
      const float v = __ldg(rs + (size_t)(c0 + i) * 256 + n);
      const float4* de = reinterpret_cast<const float4*>(s_de[i]);   // (28 scalar broadcast reads per ray made the loop LDS-bound)
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float4 d = de[q];
        acc[4 * q] = fmaf(v, d.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(v, d.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, d.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, d.w, acc[4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 28; ++j) s_red[rl][threadIdx.x & 31][j] = acc[j];
  __syncthreads();
  // 32 rows x 28 columns of this block: thread -> (row, column) pairs
  for (int e = threadIdx.x; e < 32 * 28; e += 256) {
    const int rr = e / 28, j = e - rr * 28;
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += s_red[l][rr][j];
    const int nn = blockIdx.x * 32 + rr;
    if (j < kDirCols) atomicAdd(a.gW[which] + (size_t)nn * a.stride[which] + a.off[which] + j, t);
    else atomicAdd(a.gb[which] + nn, t);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    const int t = threadIdx.x;
    a.gb_d2[t] += a.cs_x[t];
    if (t < 64) { a.gb_s0[t] += a.cs_g[t]; a.gb_i0[t] += a.cs_g[64 + t]; a.gb_d0[t] += a.cs_h0[t]; }
    if (t < 19) a.gb_s2[t] += a.cs_hs1[t];
    if (t == 19) a.gb_i2[0] += a.cs_hs1[19];
    if (t < 3) a.gb_rgb[t] += a.cs_rgb[t];
  }
}

}  // namespace wgrad
}  // namespace nlb

using namespace nlb;
using namespace nlb::wgrad;

static int log2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

extern "C" int nlb_nerf_mlp_wgrad(const nlb_nerf_mlp_saved_t* sv, const nlb_nerf_mlp_grad_out_t* go, int M,
                                  const nlb_nerf_mlp_wgrads_t* g, void* stream) {
  if (M == 0) return NLB_OK;
  if (!sv || !go || !g) { nlb_set_error("nerf_mlp_wgrad: null pointer"); return NLB_EINVAL; }
  if (!sv->f0 || !sv->h0 || !sv->x || !sv->g || !sv->h1 || !sv->h2) {
    nlb_set_error("nerf_mlp_wgrad: the activations saved by the forward (f0, h0, x, g, h1, h2) are required");
    return NLB_EINVAL;
  }
  if (!go->d_rgb || !go->d_v1 || !go->d_v0 || !go->d_hs1 || !go->d_g || !go->d_x || !go->d_h0) {
    nlb_set_error("nerf_mlp_wgrad: the pre-activation gradients written by nlb_nerf_mlp_backward are required");
    return NLB_EINVAL;
  }
  if (!g->W_d0 || !g->W_d2 || !g->W_s0 || !g->W_s2 || !g->W_i0 || !g->W_i2 || !g->W_v0 || !g->W_v1 || !g->W_rgb) {
    nlb_set_error("nerf_mlp_wgrad: null weight-gradient pointer");
    return NLB_EINVAL;
  }
  const int ld_v1 = go->ld_v1 ? go->ld_v1 : 256, ld_v0 = go->ld_v0 ? go->ld_v0 : 256, ld_g = go->ld_g ? go->ld_g : 128;
  const void* ptrs[] = {sv->f0, sv->h0, sv->x, sv->g, sv->h1, sv->h2, go->d_rgb, go->d_v1, go->d_v0, go->d_hs1, go->d_g, go->d_x, go->d_h0};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) { nlb_set_error("nerf_mlp_wgrad: operand matrices must be 16-byte aligned"); return NLB_EINVAL; }
  if ((ld_v1 | ld_v0 | ld_g) & 7) { nlb_set_error("nerf_mlp_wgrad: leading dimensions must be multiples of 8 elements"); return NLB_EINVAL; }

  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(k_nerf_mlp_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);

  auto bf = [](const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); };
  auto stream_of = [&](const void* p, int ld, int cols, int off) { return Stream{bf(p), ld, log2i(cols / 8), off}; };
  auto plain = [](int a_off, int b_off, int n, int col, float* base, int stride, int cols_valid, int rows_valid) {
    Op o{};
    o.a_off = a_off; o.b_off = b_off; o.n = n; o.tmem_col = col;
    o.base[0] = base; o.base[1] = base + (size_t)64 * stride;
    o.stride = stride;
    o.col_lo[0] = o.col_lo[1] = 0; o.col_hi[0] = o.col_hi[1] = cols_valid;
    o.rows_valid = rows_valid; o.mode = 0;
    return o;
  };
  Params P{};
  P.M = M;
  // CTA shares of the roles.  Bytes per row are 1024 / 1024 / 1024 / 768 / 896 / 864, but the two roles with four
  // operand streams and small-N products cost more per stage than their bytes: measured on B200
  // (tools/wgrad_bench.py, 327 680 rows) 0.42 ms with byte-proportional shares, 0.40 equal, 0.387 with these.
  int bytes_per_row[kNumRoles] = {4, 4, 4, 3, 4, 4};
  if (const char* e = getenv("NLB_WGRAD_SPLIT")) {   // dev switch: relative CTA shares of the six roles
    int v[kNumRoles];
    if (sscanf(e, "%d,%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4, v + 5) == kNumRoles)
      for (int i = 0; i < kNumRoles; ++i) bytes_per_row[i] = v[i] > 0 ? v[i] : 1;
  }
  {
    Role& R = P.role[0];   // d_v0^T x
    R.ns = 2; R.s[0] = stream_of(sv->x, 256, 256, 0); R.s[1] = stream_of(go->d_v0, ld_v0, 256, 4 * kBlk);
    R.nop = 2;
    for (int mt = 0; mt < 2; ++mt) R.op[mt] = plain(4 * kBlk + mt * 2 * kBlk, 0, 256, mt * 256, g->W_v0 + (size_t)mt * 128 * 283, 283, 256, 128);
  }
  {
    Role& R = P.role[1];   // d_v1^T x
    R.ns = 2; R.s[0] = stream_of(sv->x, 256, 256, 0); R.s[1] = stream_of(go->d_v1, ld_v1, 256, 4 * kBlk);
    R.nop = 2;
    for (int mt = 0; mt < 2; ++mt) R.op[mt] = plain(4 * kBlk + mt * 2 * kBlk, 0, 256, mt * 256, g->W_v1 + (size_t)mt * 128 * 539 + 256, 539, 256, 128);
  }
  {
    Role& R = P.role[2];   // d_v1^T h1
    R.ns = 2; R.s[0] = stream_of(sv->h1, 256, 256, 0); R.s[1] = stream_of(go->d_v1, ld_v1, 256, 4 * kBlk);
    R.nop = 2;
    for (int mt = 0; mt < 2; ++mt) R.op[mt] = plain(4 * kBlk + mt * 2 * kBlk, 0, 256, mt * 256, g->W_v1 + (size_t)mt * 128 * 539, 539, 256, 128);
  }
  {
    Role& R = P.role[3];   // d_g^T x -> sem_layer.0 (rows 0..63) | intensity_layer.0 (rows 64..127)
    R.ns = 2; R.s[0] = stream_of(sv->x, 256, 256, 0); R.s[1] = stream_of(go->d_g, ld_g, 128, 4 * kBlk);
    R.nop = 1;
    R.op[0] = plain(4 * kBlk, 0, 256, 0, g->W_s0, 256, 256, 128);
    R.op[0].base[1] = g->W_i0;
  }
  {
    Role& R = P.role[4];   // d_x^T h0 -> density_layer.2 ; d_h0^T f0 -> density_layer.0
    R.ns = 4;
    R.s[0] = stream_of(sv->h0, 64, 64, 0);
    R.s[1] = stream_of(go->d_x, 256, 256, kBlk);
    R.s[2] = stream_of(sv->f0, 64, 64, 5 * kBlk);
    R.s[3] = stream_of(go->d_h0, 64, 64, 6 * kBlk);   // (the M = 128 operand also covers the unused slab 7: rows 64..127 of D are ignored)
    R.nop = 3;
    for (int mt = 0; mt < 2; ++mt) R.op[mt] = plain(kBlk + mt * 2 * kBlk, 0, 64, mt * 64, g->W_d2 + (size_t)mt * 128 * 64, 64, 64, 128);
    R.op[2] = plain(6 * kBlk, 5 * kBlk, 64, 128, g->W_d0, 40, 40, 64);
  }
  {
    Role& R = P.role[5];   // g^T d_hs1 -> sem_layer.2^T | intensity_layer.2^T ; h2^T d_rgb -> rgb_layer^T
    R.ns = 4;
    R.s[0] = stream_of(sv->g, 128, 128, 0);
    R.s[1] = stream_of(go->d_hs1, 32, 32, 2 * kBlk);
    R.s[2] = stream_of(sv->h2, 256, 256, 3 * kBlk);
    R.s[3] = stream_of(go->d_rgb, 16, 16, 7 * kBlk);
    R.nop = 3;
    Op o{};
    o.a_off = 0; o.b_off = 2 * kBlk; o.n = 32; o.tmem_col = 0; o.mode = 1; o.stride = 64; o.rows_valid = 128;
    o.base[0] = g->W_s2; o.col_lo[0] = 0; o.col_hi[0] = 19;                       // D[j, k] -> W_s2[k, j], j < 64
    o.base[1] = g->W_i2 - 19 * 64; o.col_lo[1] = 19; o.col_hi[1] = 20;            // D[64 + j, 19] -> W_i2[0, j]
    R.op[0] = o;
    for (int mt = 0; mt < 2; ++mt) {
      Op q{};
      q.a_off = 3 * kBlk + mt * 2 * kBlk; q.b_off = 7 * kBlk; q.n = 16; q.tmem_col = 32 + mt * 16; q.mode = 1; q.stride = 256;
      q.rows_valid = 128;
      q.base[0] = g->W_rgb + mt * 128; q.base[1] = g->W_rgb + mt * 128 + 64;      // D[u, c] -> W_rgb[c, u]
      q.col_lo[0] = q.col_lo[1] = 0; q.col_hi[0] = q.col_hi[1] = 3;
      R.op[1 + mt] = q;
    }
  }
  // CTAs per role in proportion to the bytes it streams, at most one per stage
  const int total_stages = (M + kRows - 1) / kRows;
  int sum_b = 0;
  for (int b : bytes_per_row) sum_b += b;
  int cta = 0, left = sms;
  for (int i = 0; i < kNumRoles; ++i) {
    int n = i == kNumRoles - 1 ? left : (int)((long long)sms * bytes_per_row[i] / sum_b);
    if (n < 1) n = 1;
    if (n > total_stages) n = total_stages;
    P.role[i].cta0 = cta;
    P.role[i].nctas = n;
    cta += n;
    left -= n;
    if (left < kNumRoles - 1 - i) left = kNumRoles - 1 - i;
  }
  if (const char* e = getenv("NLB_WGRAD_ONLY")) {    // dev switch: time ONE role on all SMs (results incomplete)
    const int only = atoi(e);
    if (only >= 0 && only < kNumRoles) {
      P.role[0] = P.role[only];
      P.role[0].cta0 = 0;
      P.role[0].nctas = sms < total_stages ? sms : total_stages;
      for (int i = 1; i < kNumRoles; ++i) P.role[i].cta0 = 1 << 30;
      cta = P.role[0].nctas;
    }
  }
  k_nerf_mlp_wgrad<<<cta, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
  return nlb_check_launch("nerf_mlp_wgrad");
}

extern "C" int nlb_nerf_mlp_wgrad_finish(const float* rs_v0, const float* rs_v1, const float* viewdirs, int N,
                                         const float* cs_x, const float* cs_g, const float* cs_h0, const float* cs_hs1,
                                         const float* cs_rgb, const nlb_nerf_mlp_wgrads_t* g, void* stream) {
  if (N == 0) return NLB_OK;
  if (!rs_v0 || !rs_v1 || !viewdirs || !cs_x || !cs_g || !cs_h0 || !cs_hs1 || !cs_rgb || !g) { nlb_set_error("nerf_mlp_wgrad_finish: null pointer"); return NLB_EINVAL; }
  if (!g->W_v0 || !g->W_v1 || !g->b_v0 || !g->b_v1 || !g->b_d2 || !g->b_s0 || !g->b_i0 || !g->b_d0 || !g->b_s2 || !g->b_i2 || !g->b_rgb) {
    nlb_set_error("nerf_mlp_wgrad_finish: null gradient pointer");
    return NLB_EINVAL;
  }
  FinishArgs a{};
  a.rs[0] = rs_v0; a.rs[1] = rs_v1;
  a.gW[0] = g->W_v0; a.gW[1] = g->W_v1;
  a.gb[0] = g->b_v0; a.gb[1] = g->b_v1;
  a.stride[0] = 283; a.stride[1] = 539; a.off[0] = 256; a.off[1] = 512;
  a.viewdirs = viewdirs; a.N = N;
  a.cs_x = cs_x; a.cs_g = cs_g; a.cs_h0 = cs_h0; a.cs_hs1 = cs_hs1; a.cs_rgb = cs_rgb;
  a.gb_d2 = g->b_d2; a.gb_s0 = g->b_s0; a.gb_i0 = g->b_i0; a.gb_d0 = g->b_d0; a.gb_s2 = g->b_s2; a.gb_i2 = g->b_i2; a.gb_rgb = g->b_rgb;
  int chunks = (N + 2 * kRayChunk - 1) / (2 * kRayChunk);
  if (chunks > 80) chunks = 80;
  if (chunks < 1) chunks = 1;
  dim3 grid(8, chunks, 2);
  k_wgrad_finish<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return nlb_check_launch("nerf_mlp_wgrad_finish");
}
