// NerfMLP forward as ONE persistent tcgen05 kernel (Z/internal/models.py:996-997,
// 1116-1251): 40->64->256 trunk, density softplus, semantic 256->64->19 softmax,
// intensity 256->64->1, view branch cat[x, pos_enc(viewdirs)] (283) -> 256 ->
// cat (539) -> 256 -> 3 sigmoid.  bf16 operands, fp32 accumulation in TMEM.
//
// Per CTA (one per SM, 416 threads = 8 epilogue warps + 1 MMA warp + 4 weight-producer
// warps): a 128-row tile of samples stays on chip for the whole chain.  Activations live in shared memory as UMMA operand blocks
// ([128 rows][64 k] bf16, K-major, SWIZZLE_128B, 16 KB each): X0-3 (bottleneck),
// H0-3 (layer outputs, reused), D (view-direction encoding).  Weights are pre-packed
// (nlb_nerf_mlp_pack) into the same block format and streamed from L2 through a
// 4 x 16 KB ring with 1-D bulk async copies (UBLKCP), one producer warp per ring stage;
// one elected thread issues tcgen05.mma (M=128, N<=128 per instruction) into two
// 256-column TMEM accumulators; eight epilogue warps (two per TMEM lane quarter, splitting
// the columns) read TMEM (LDTM) in 64-column groups, apply bias/activation and write the
// next layer's A operand back to shared memory; the next tile's inputs are staged while
// the last layer's MMAs run.  The concatenations
// of the reference are just extra K blocks (X, D) of the next GEMM.
// Layer order on the tensor pipe: L0, L1, HS0 (sem|int hidden), V0, HS1 (sem|int
// out), V1, RGB; the HS0/HS1 epilogues overlap the V0 MMAs.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/nlb200.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace nlb {
namespace mlp {

using namespace nlb::umma;

constexpr int kFeat = 40, kDir = 27, kSem = 19;
constexpr int kBlockBytes = 16384;            // [128][64] bf16
// shared-memory A blocks
constexpr int BX = 0, BH = 4, BD = 8, kNumABlocks = 9;
constexpr int kStages = 4;
// warps 0-7 epilogue (warp w and w+4 share TMEM lane quarter w%4 and split the columns),
// warp 8 MMA issuer, warps 9-12 weight producers (one per ring stage)
constexpr int kEpiWarps = 8, kMmaWarp = 8, kProdWarp0 = 9;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = (kProdWarp0 + kStages) * 32;

struct LayerDef {
  int N, nkb;
  int a_blk[9];
  int ksteps[9];
  int tmem_col;
};
enum { L0 = 0, L1, HS0, V0, HS1, V1, RGB, kNumLayers };
__host__ __device__ constexpr LayerDef layer_def(int l) {
  switch (l) {
    case L0:  return {64, 1, {BH + 0}, {3}, 0};
    case L1:  return {256, 1, {BH + 1}, {4}, 256};
    case HS0: return {128, 4, {BX, BX + 1, BX + 2, BX + 3}, {4, 4, 4, 4}, 0};
    case V0:  return {256, 5, {BX, BX + 1, BX + 2, BX + 3, BD}, {4, 4, 4, 4, 2}, 256};
    case HS1: return {32, 2, {BH + 2, BH + 3}, {4, 4}, 128};
    case V1:  return {256, 9, {BH, BH + 1, BH + 2, BH + 3, BX, BX + 1, BX + 2, BX + 3, BD}, {4, 4, 4, 4, 4, 4, 4, 4, 2}, 0};
    default:  return {16, 4, {BH, BH + 1, BH + 2, BH + 3}, {4, 4, 4, 4}, 256};
  }
}
__host__ __device__ constexpr int layer_nrows(int l) { return layer_def(l).N > 128 ? 128 : layer_def(l).N; }
__host__ __device__ constexpr int layer_nhalves(int l) { return layer_def(l).N > 128 ? 2 : 1; }
__host__ __device__ constexpr int layer_chunks(int l) { return layer_def(l).nkb * layer_nhalves(l); }
__host__ __device__ constexpr int layer_chunk_bytes(int l) { return layer_nrows(l) * 128; }
__host__ __device__ constexpr int layer_offset(int l) {  // byte offset of the layer's first chunk in the blob
  int o = 0;
  for (int i = 0; i < l; ++i) o += layer_chunks(i) * layer_chunk_bytes(i);
  return o;
}
__host__ __device__ constexpr int bias_offset(int l) {  // float index in the bias section
  int o = 0;
  for (int i = 0; i < l; ++i) o += layer_def(i).N;
  return o;
}
constexpr int kWeightBytes = layer_offset(kNumLayers);
constexpr int kBiasFloats = bias_offset(kNumLayers);
constexpr int kPackedBytes = kWeightBytes + kBiasFloats * 4;

// ----------------------------------------------------------------------------- packing
// value of the (padded) weight matrix of layer l at output row n, input column k
__device__ float packed_weight(const nlb_nerf_mlp_weights_t& w, int l, int n, int k) {
  switch (l) {
    case L0:  return (n < 64 && k < kFeat) ? w.W_d0[n * kFeat + k] : 0.f;
    case L1:  return w.W_d2[n * 64 + k];
    case HS0: return n < 64 ? w.W_s0[n * 256 + k] : (w.W_i0 ? w.W_i0[(n - 64) * 256 + k] : 0.f);   // no intensity head: zeros
    case V0:  return k < 256 ? w.W_v0[n * 283 + k] : (k - 256 < kDir ? w.W_v0[n * 283 + k] : 0.f);
    case HS1:
      if (n < kSem) return k < 64 ? w.W_s2[n * 64 + k] : 0.f;
      if (n == kSem) return (k >= 64 && w.W_i2) ? w.W_i2[k - 64] : 0.f;
      return 0.f;
    case V1:  return k < 512 ? w.W_v1[n * 539 + k] : (k - 512 < kDir ? w.W_v1[n * 539 + k] : 0.f);
    default:  return n < 3 ? w.W_rgb[n * 256 + k] : 0.f;
  }
}
__device__ float packed_bias(const nlb_nerf_mlp_weights_t& w, int l, int n) {
  switch (l) {
    case L0:  return w.b_d0[n];
    case L1:  return w.b_d2[n];
    case HS0: return n < 64 ? w.b_s0[n] : (w.b_i0 ? w.b_i0[n - 64] : 0.f);
    case V0:  return w.b_v0[n];
    case HS1: return n < kSem ? w.b_s2[n] : ((n == kSem && w.b_i2) ? w.b_i2[0] : 0.f);
    case V1:  return w.b_v1[n];
    default:  return n < 3 ? w.b_rgb[n] : 0.f;
  }
}

__global__ void k_pack(nlb_nerf_mlp_weights_t w, uint8_t* __restrict__ blob) {
  for (int l = 0; l < kNumLayers; ++l) {
    const LayerDef d = layer_def(l);
    const int nrows = layer_nrows(l), nh_count = layer_nhalves(l);
    const int total = d.nkb * nh_count * nrows * 64;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int c = e & 63;
      const int r = (e >> 6) % nrows;
      const int chunk = (e >> 6) / nrows;       // kb-major, n-half minor
      const int kb = chunk / nh_count, nh = chunk % nh_count;
      const float v = packed_weight(w, l, nh * 128 + r, kb * 64 + c);
      uint8_t* dst = blob + layer_offset(l) + (size_t)chunk * layer_chunk_bytes(l) + sw128_offset(r, c);
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v);
    }
    float* bias = reinterpret_cast<float*>(blob + kWeightBytes) + bias_offset(l);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < d.N; n += gridDim.x * blockDim.x) bias[n] = packed_bias(w, l, n);
  }
}

// optional timeline of block 0 / first two tiles (dev tool: nlb_debug_set_timeline)
__device__ long long* g_timeline = nullptr;
__device__ __forceinline__ void stamp(int slot) {
  if (g_timeline && blockIdx.x == 0) g_timeline[slot] = clock64();
}

// ----------------------------------------------------------------------------- kernel
struct Smem {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t acc_ready[kNumLayers];
  uint64_t a_ready[6];
  uint32_t tmem_base;
  float bias[kBiasFloats];
};
enum { E_F = 0, E_H0, E_X, E_G, E_H1, E_H2 };

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 values of this thread's row -> bf16 -> A block (swizzled) and, optionally, a
// row-major bf16 copy in global memory (saved for the backward pass / wgrad GEMMs)
__device__ __forceinline__ void store_cols(const float (&v)[32], uint8_t* block, int row, int col_in_block,
                                           __nv_bfloat16* gdst /*this row, first of the 32 columns, or null*/) {
  uint8_t* rowp = block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((col_in_block >> 3) + q) ^ (row & 7);
    uint4 u;
    u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
    if (gdst) reinterpret_cast<uint4*>(gdst)[q] = u;
  }
}

// 32 accumulator columns [c0, c0+32) of this thread's row -> bias (+relu) -> bf16 -> A block
template <bool kRelu>
__device__ __forceinline__ void epi_cols_to_block(uint32_t taddr, const float* __restrict__ bias, uint8_t* block,
                                                  int row, int col_in_block, float* keep0 = nullptr,
                                                  __nv_bfloat16* gdst = nullptr) {
  float v[32];
  tmem_ld32(taddr, v);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] += bias[i];
    if (kRelu) v[i] = fmaxf(v[i], 0.f);
  }
  if (keep0) *keep0 = v[0];
  store_cols(v, block, row, col_in_block, gdst);
}

// ---- 64-column epilogue groups ---------------------------------------------------------
// A warp owns 32 rows of the tile; one group = those rows x one whole operand block (64
// columns, 128 bytes per row).  Global traffic of a group is coalesced through the block
// itself: each thread writes (reads) its own row in shared memory, and after a warp barrier
// the warp moves 4 rows x 128 contiguous bytes per instruction to (from) the row-major
// bf16 matrix.  (Per-thread 64-byte stores at a 512-byte row stride touched 32 lines per
// instruction and kept the epilogue -- the critical path of the tile -- waiting on the LSU.)
__device__ __forceinline__ uint8_t* block_row(uint8_t* block, int row) { return block + (row >> 3) * 1024 + (row & 7) * 128; }

// this thread's 64 values -> bf16 -> its row of the swizzled block
__device__ __forceinline__ void store_row64(const float (&v)[64], uint8_t* block, int row) {
  uint8_t* rowp = block_row(block, row);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 u;
    u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + ((q ^ (row & 7)) * 16)) = u;
  }
}

// rows [wrow0, wrow0+32) of the block -> g[(row) * ld + 0..63] (g = tile base + column offset)
__device__ __forceinline__ void warp_rows_to_global(const uint8_t* block, int wrow0, int lane, __nv_bfloat16* g, int ld,
                                                    int rows_valid) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = wrow0 + i * 4 + (lane >> 3), ch = lane & 7;
    const uint4 u = *reinterpret_cast<const uint4*>(block + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) * 16));
    if (row < rows_valid) *reinterpret_cast<uint4*>(g + (size_t)row * ld + ch * 8) = u;
  }
}

// g[(row) * ld + 0..63] -> rows [wrow0, wrow0+32) of the block (zeros beyond rows_valid)
__device__ __forceinline__ void warp_rows_from_global(uint8_t* block, int wrow0, int lane, const __nv_bfloat16* g, int ld,
                                                      int rows_valid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = wrow0 + i * 4 + (lane >> 3), ch = lane & 7;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) u = __ldg(reinterpret_cast<const uint4*>(g + (size_t)row * ld + ch * 8));
    *reinterpret_cast<uint4*>(block + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) * 16)) = u;
  }
  __syncwarp();
}

// same, asynchronously (no registers, no wait): rows beyond rows_valid are zero-filled; complete with
// cp_async_wait_all() + __syncwarp() before reading
__device__ __forceinline__ void warp_rows_from_global_async(uint8_t* block, int wrow0, int lane, const __nv_bfloat16* g,
                                                            int ld, int rows_valid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = wrow0 + i * 4 + (lane >> 3), ch = lane & 7;
    const bool ok = row < rows_valid;
    cp_async16(block + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) * 16),
               g + (size_t)(ok ? row : 0) * ld + ch * 8, ok);
  }
}

// forward group: 64 accumulator columns of this thread's row -> bias (+relu) -> block (+ saved copy)
template <bool kRelu>
__device__ __forceinline__ void epi_group64(uint32_t taddr, const float* __restrict__ bias, uint8_t* block, int row,
                                            int lane, __nv_bfloat16* gsave, int ld, int rows_valid,
                                            float* keep0 = nullptr) {
  float v[64];
  tmem_ld64(taddr, v);
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    v[i] += bias[i];
    if (kRelu) v[i] = fmaxf(v[i], 0.f);
  }
  if (keep0) *keep0 = v[0];
  store_row64(v, block, row);
  if (gsave) warp_rows_to_global(block, row & ~31, lane, gsave, ld, rows_valid);
}

__device__ __forceinline__ void signal_a_ready(uint64_t* bar) {
  tcgen05_fence_before();
  fence_proxy_async();
  mbar_arrive(bar);
}

// MMAs of forward layer L for one tile: for every (K block, N half) chunk wait for its
// weights in the ring, issue the chunk's K-steps, release the ring slot when they retire;
// finally signal the layer's accumulator.  `c` = running chunk counter (ring position).
template <int L, class SmemT>
__device__ __forceinline__ void issue_layer(SmemT& sm, uint8_t* a_blocks, uint8_t* w_ring, uint32_t tmem, uint32_t& c) {
  constexpr LayerDef d = layer_def(L);
  constexpr int nrows = layer_nrows(L), nh_count = layer_nhalves(L);
  const uint32_t idesc = make_idesc_bf16(128, nrows);
#pragma unroll
  for (int kb = 0; kb < d.nkb; ++kb) {
    const uint64_t adesc = make_desc_sw128(a_blocks + d.a_blk[kb] * kBlockBytes);
#pragma unroll
    for (int nh = 0; nh < nh_count; ++nh, ++c) {
      const uint32_t st = c % kStages;
      mbar_wait(&sm.w_full[st], (c / kStages) & 1);
      tcgen05_fence_after();
      const uint64_t bdesc = make_desc_sw128(w_ring + st * kBlockBytes);
      const uint32_t dcol = tmem + d.tmem_col + nh * 128;
      switch (d.ksteps[kb]) {
        case 1: mma_chunk<1>(dcol, adesc, bdesc, idesc, kb != 0); break;
        case 2: mma_chunk<2>(dcol, adesc, bdesc, idesc, kb != 0); break;
        case 3: mma_chunk<3>(dcol, adesc, bdesc, idesc, kb != 0); break;
        default: mma_chunk<4>(dcol, adesc, bdesc, idesc, kb != 0); break;
      }
      mma_commit(&sm.w_empty[st]);
    }
  }
  mma_commit(&sm.acc_ready[L]);
}

__global__ void __launch_bounds__(kThreads, 1) k_nerf_mlp_fwd(const float* __restrict__ features,
                                                              const float* __restrict__ viewdirs, int M,
                                                              int rows_per_ray, const uint8_t* __restrict__ blob,
                                                              float* __restrict__ o_density, float* __restrict__ o_rgb,
                                                              float* __restrict__ o_sem, float* __restrict__ o_int,
                                                              nlb_nerf_mlp_saved_t sv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_blocks = base;                                  // 9 x 16 KB
  uint8_t* w_ring = base + kNumABlocks * kBlockBytes;        // 4 x 16 KB
  Smem& sm = *reinterpret_cast<Smem*>(w_ring + kStages * kBlockBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    for (int i = 0; i < kNumLayers; ++i) mbar_init(&sm.acc_ready[i], 1);
    for (int i = 0; i < 6; ++i) mbar_init(&sm.a_ready[i], kEpiThreads);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kBiasFloats; i += kThreads)
    sm.bias[i] = reinterpret_cast<const float*>(blob + kWeightBytes)[i];
  if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp >= kProdWarp0) {
    // ===== weight producers: a 1-D bulk copy costs its issuing thread ~800 cycles whatever its
    // size and consecutive copies of one thread do not overlap (tools/bulk_probe.cu), but
    // different warps issue concurrently -> one producer warp per ring stage.
    if (elect_one_sync()) {
      const int my_stage = warp - kProdWarp0;
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int l = 0; l < kNumLayers; ++l) {
          const int n = layer_chunks(l), bytes = layer_chunk_bytes(l);
          const uint8_t* src = blob + layer_offset(l);
#pragma unroll 1
          for (int i = 0; i < n; ++i, ++c) {
            const int st = c % kStages;
            if (st != my_stage) continue;
            mbar_wait_relaxed(&sm.w_empty[st], ((c / kStages) & 1) ^ 1);
            mbar_expect_tx(&sm.w_full[st], bytes);
            bulk_g2s(w_ring + st * kBlockBytes, src + (size_t)i * bytes, bytes, &sm.w_full[st]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: one elected lane, every layer's chunk loop unrolled at compile time
    if (elect_one_sync()) {
      uint32_t c = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int tb = it < 2 ? (int)it * 16 : -1000;
        if (tb >= 0) stamp(tb + 0);
        mbar_wait(&sm.a_ready[E_F], ph);  tcgen05_fence_after(); if (tb >= 0) stamp(tb + 1);
        issue_layer<L0>(sm, a_blocks, w_ring, tmem, c);  if (tb >= 0) stamp(tb + 2);
        mbar_wait(&sm.a_ready[E_H0], ph); tcgen05_fence_after(); if (tb >= 0) stamp(tb + 3);
        issue_layer<L1>(sm, a_blocks, w_ring, tmem, c);  if (tb >= 0) stamp(tb + 4);
        mbar_wait(&sm.a_ready[E_X], ph);  tcgen05_fence_after(); if (tb >= 0) stamp(tb + 5);
        issue_layer<HS0>(sm, a_blocks, w_ring, tmem, c); if (tb >= 0) stamp(tb + 6);
        issue_layer<V0>(sm, a_blocks, w_ring, tmem, c);  if (tb >= 0) stamp(tb + 7);
        mbar_wait(&sm.a_ready[E_G], ph);  tcgen05_fence_after(); if (tb >= 0) stamp(tb + 8);
        issue_layer<HS1>(sm, a_blocks, w_ring, tmem, c); if (tb >= 0) stamp(tb + 9);
        mbar_wait(&sm.a_ready[E_H1], ph); tcgen05_fence_after(); if (tb >= 0) stamp(tb + 10);
        issue_layer<V1>(sm, a_blocks, w_ring, tmem, c);  if (tb >= 0) stamp(tb + 11);
        mbar_wait(&sm.a_ready[E_H2], ph); tcgen05_fence_after(); if (tb >= 0) stamp(tb + 12);
        issue_layer<RGB>(sm, a_blocks, w_ring, tmem, c); if (tb >= 0) stamp(tb + 13);
      }
    }
  } else {
    // ===== epilogue warps 0..7: thread (q, lane) owns row r = 32 q + lane (TMEM lane r); the two
    // warps of a lane quarter (half = 0 / 1) split every layer's columns
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* HB = a_blocks + BH * kBlockBytes;
    uint8_t* XB = a_blocks + BX * kBlockBytes;
    uint8_t* DB = a_blocks + BD * kBlockBytes;
    // staging of a tile's inputs: this thread's feature row -> registers -> H0, and the direction encoding -> D
    auto load_features = [&](int tile_, float (&f)[kFeat]) {
      const int row_ = tile_ * 128 + r;
      if (tile_ < num_tiles && row_ < M) {
        const float4* src = reinterpret_cast<const float4*>(features + (size_t)row_ * kFeat);
#pragma unroll
        for (int q = 0; q < kFeat / 4; ++q) {
          const float4 t = __ldg(src + q);
          f[q * 4] = t.x; f[q * 4 + 1] = t.y; f[q * 4 + 2] = t.z; f[q * 4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < kFeat; ++i) f[i] = 0.f;
      }
    };
    auto store_features = [&](const float (&f)[kFeat]) {
      uint8_t* rowp = HB + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (q < kFeat / 8) {
          u.x = pack_bf16(f[q * 8], f[q * 8 + 1]); u.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
          u.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]); u.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
        }
        *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = u;
      }
    };
    auto stage_dirs = [&](int tile_) {
      const int row_ = tile_ * 128 + r;
      float d[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) d[i] = 0.f;
      if (tile_ < num_tiles && row_ < M) {
        const int ray = row_ / rows_per_ray;
        const float vx = __ldg(viewdirs + 3 * ray), vy = __ldg(viewdirs + 3 * ray + 1), vz = __ldg(viewdirs + 3 * ray + 2);
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
      }
      uint8_t* drow = DB + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (q < 4) {
          u.x = pack_bf16(d[q * 8], d[q * 8 + 1]); u.y = pack_bf16(d[q * 8 + 2], d[q * 8 + 3]);
          u.z = pack_bf16(d[q * 8 + 4], d[q * 8 + 5]); u.w = pack_bf16(d[q * 8 + 6], d[q * 8 + 7]);
        }
        *reinterpret_cast<uint4*>(drow + ((q ^ (r & 7)) * 16)) = u;
      }
    };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int row = tile * 128 + r;
      const bool valid = row < M;
      const int rows_valid = M - tile * 128 < 128 ? M - tile * 128 : 128;
      // ---- features (H0: half 0) and view-direction encoding (D: half 1) of the FIRST tile; every later
      // tile was staged during the previous one (see the end of the loop body)
      if (it == 0) {
        if (half == 0) {
          float f[kFeat];
          load_features(tile, f);
          store_features(f);
          if (sv.f0) warp_rows_to_global(HB, r & ~31, lane, reinterpret_cast<__nv_bfloat16*>(sv.f0) + (size_t)tile * 128 * 64, 64, rows_valid);
        } else {
          stage_dirs(tile);
        }
      }
      signal_a_ready(&sm.a_ready[E_F]);
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 14);

      // ---- L0: h0 = relu(acc + b) -> H1
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 0);
      mbar_wait_warp(&sm.acc_ready[L0], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 1);
      if (half == 0)
        epi_group64<true>(tlane, sm.bias + bias_offset(L0), HB + 1 * kBlockBytes, r, lane,
                          sv.h0 ? reinterpret_cast<__nv_bfloat16*>(sv.h0) + (size_t)tile * 128 * 64 : nullptr, 64, rows_valid);
      signal_a_ready(&sm.a_ready[E_H0]);

      // ---- L1: x = acc + b -> X0..3 ; density = softplus(x[0] - 1)
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 2);
      mbar_wait_warp(&sm.acc_ready[L1], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 3);
      float x0 = 0.f;
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64<false>(tlane + 256 + c0, sm.bias + bias_offset(L1) + c0, XB + (c0 >> 6) * kBlockBytes, r, lane,
                           sv.x ? reinterpret_cast<__nv_bfloat16*>(sv.x) + (size_t)tile * 128 * 256 + c0 : nullptr, 256,
                           rows_valid, c0 == 0 ? &x0 : nullptr);
      signal_a_ready(&sm.a_ready[E_X]);
      if (valid && half == 0) {
        const float xin = x0 - 1.0f;
        o_density[row] = xin > 20.f ? xin : log1pf(expf(xin));
      }

      // ---- HS0: hidden = relu(acc + b) -> H2, H3
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 4);
      mbar_wait_warp(&sm.acc_ready[HS0], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 5);
      {
        const int c0 = half * 64;
        epi_group64<true>(tlane + c0, sm.bias + bias_offset(HS0) + c0, HB + (2 + half) * kBlockBytes, r, lane,
                          sv.g ? reinterpret_cast<__nv_bfloat16*>(sv.g) + (size_t)tile * 128 * 128 + c0 : nullptr, 128,
                          rows_valid);
      }
      signal_a_ready(&sm.a_ready[E_G]);

      // ---- HS1: semantic softmax (19) + intensity
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 6);
      // (every warp waits: the V0 epilogue below overwrites H2/H3, which the HS1 MMAs read)
      mbar_wait_warp(&sm.acc_ready[HS1], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 7);
      if (half == 0) {
        float v[32];
        tmem_ld32(tlane + 128, v);
        const float* b = sm.bias + bias_offset(HS1);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kSem; ++i) { v[i] += b[i]; mx = fmaxf(mx, v[i]); }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kSem; ++i) { v[i] = expf(v[i] - mx); sum += v[i]; }
        const float inv = 1.0f / sum;
        if (valid) {
          if (o_sem) {
#pragma unroll
            for (int i = 0; i < kSem; ++i) o_sem[(size_t)row * kSem + i] = v[i] * inv;
          }
          if (o_int) o_int[row] = v[kSem] + b[kSem];
        }
      }

      // ---- V0: h1 = relu(acc + b) -> H0..3
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 8);
      mbar_wait_warp(&sm.acc_ready[V0], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 9);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64<true>(tlane + 256 + c0, sm.bias + bias_offset(V0) + c0, HB + (c0 >> 6) * kBlockBytes, r, lane,
                          sv.h1 ? reinterpret_cast<__nv_bfloat16*>(sv.h1) + (size_t)tile * 128 * 256 + c0 : nullptr, 256,
                          rows_valid);
      signal_a_ready(&sm.a_ready[E_H1]);

      // ---- V1: h2 = relu(acc + b) -> H0..3
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 10);
      mbar_wait_warp(&sm.acc_ready[V1], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 11);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64<true>(tlane + c0, sm.bias + bias_offset(V1) + c0, HB + (c0 >> 6) * kBlockBytes, r, lane,
                          sv.h2 ? reinterpret_cast<__nv_bfloat16*>(sv.h2) + (size_t)tile * 128 * 256 + c0 : nullptr, 256,
                          rows_valid);
      signal_a_ready(&sm.a_ready[E_H2]);

      // ---- next tile's inputs, while the RGB MMAs run: D is free (the V1 MMAs have completed), the
      // features wait in registers until the RGB MMAs have released H0
      const int next_tile = tile + (int)gridDim.x;
      float fnext[kFeat];
      if (half == 1) stage_dirs(next_tile);
      else load_features(next_tile, fnext);

      // ---- RGB: sigmoid(acc + b) * (1 + 2 pad) - pad
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 12);
      mbar_wait_warp(&sm.acc_ready[RGB], ph);
      tcgen05_fence_after();
      if (it < 2 && r == 0) stamp(64 + (int)it * 16 + 13);
      if (half == 0) {
        store_features(fnext);
        // bf16 copy of the (zero-padded) feature rows: the B operand of density_layer.0's weight gradient
        if (sv.f0 && next_tile < num_tiles)
          warp_rows_to_global(HB, r & ~31, lane, reinterpret_cast<__nv_bfloat16*>(sv.f0) + (size_t)next_tile * 128 * 64, 64,
                              M - next_tile * 128 < 128 ? M - next_tile * 128 : 128);
      } else {
        float v[16];
        tmem_ld16(tlane + 256, v);
        const float* b = sm.bias + bias_offset(RGB);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float s = 1.0f / (1.0f + expf(-(v[i] + b[i])));
            o_rgb[(size_t)row * 3 + i] = s * (1.0f + 2.0f * 0.001f) - 0.001f;
          }
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

// =============================================================================
// Forward, two tiles in flight per CTA (default).  The one-tile kernel above leaves the tensor pipe
// idle during five dependent epilogues and the staging of the next tile (clock64 timeline: 34 K cycles
// per tile against 14 K of MMA issue).  A second 128-row tile cannot be resident (144 KB of operand
// blocks each), but the TRUNK of the next tile can: its features, h0 and x live in the X blocks, which
// the current tile stops reading when its V1 MMAs complete.  So the roles are
//   warps 0-7   main epilogue warps (two per TMEM lane quarter): HS0, V0, V1, RGB of tile i
//   warps 8-11  front warps (one per lane quarter): stage features / view-direction encoding of tile
//               i+1, its L0 and L1 epilogues (-> X, density), and the semantic / intensity head
//               outputs of tile i+1 while the main warps are in its view branch
//   warp 12     MMA issuer: HS0 V0 HS1 V1 (tile i) | L0 L1 (tile i+1) | RGB (tile i)
//   warps 13-16 weight producers (one per ring stage), streaming in that order
// TMEM: V1 / HS0 (0) / HS1 (128) / RGB (224) share columns 0-255, V0 / L0 / L1 share 256-511; every
// reuse is ordered by an mbarrier that the previous reader arrives on (see the hazard notes inline).
namespace v2 {

constexpr int kMainWarps = 8, kFrontWarp0 = 8, kFrontWarps = 4, kMmaWarp2 = 12, kProdWarp2 = 13;
constexpr int kMainThreads = kMainWarps * 32, kFrontThreads = kFrontWarps * 32;
constexpr int kThreads2 = (kProdWarp2 + kStages) * 32;

__host__ __device__ constexpr LayerDef layer_def2(int l) {
  LayerDef d = layer_def(l);
  switch (l) {
    case L0:  d.a_blk[0] = BX + 0; d.tmem_col = 256; break;   // features of the next tile in X0
    case L1:  d.a_blk[0] = BX + 1; d.tmem_col = 256; break;   // h0 in X1
    case HS0: d.tmem_col = 0; break;
    case V0:  d.tmem_col = 256; break;
    case HS1: d.tmem_col = 128; break;
    // h1 / h2 column block j lives in H block hperm(j) = {2, 0, 3, 1}[j]: each main warp then only ever
    // overwrites rows of blocks it wrote itself (g in H2 / H3 by column half), so its global saves can
    // trail its "ready" signal with program order as the only hazard protection
    case V1:  d.tmem_col = 0; d.a_blk[0] = BH + 2; d.a_blk[1] = BH + 0; d.a_blk[2] = BH + 3; d.a_blk[3] = BH + 1; break;
    default:  d.tmem_col = 224; d.a_blk[0] = BH + 2; d.a_blk[1] = BH + 0; d.a_blk[2] = BH + 3; d.a_blk[3] = BH + 1; break;  // RGB
  }
  return d;
}

__host__ __device__ constexpr int hperm(int j) { return j == 0 ? 2 : (j == 1 ? 0 : (j == 2 ? 3 : 1)); }

struct Smem2 {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t acc_ready[kNumLayers];
  uint64_t f_ready, h0_ready, x_ready, hs1_free;   // front warps arrive (128)
  uint64_t g_ready, h1_ready, h2_ready;            // main warps arrive (256)
  uint32_t tmem_base;
  float bias[kBiasFloats];
};

template <int L>
__device__ __forceinline__ void issue_layer2(Smem2& sm, uint8_t* a_blocks, uint8_t* w_ring, uint32_t tmem, uint32_t& c) {
  constexpr LayerDef d = layer_def2(L);
  constexpr int nrows = layer_nrows(L), nh_count = layer_nhalves(L);
  const uint32_t idesc = make_idesc_bf16(128, nrows);
#pragma unroll
  for (int kb = 0; kb < d.nkb; ++kb) {
    const uint64_t adesc = make_desc_sw128(a_blocks + d.a_blk[kb] * kBlockBytes);
#pragma unroll
    for (int nh = 0; nh < nh_count; ++nh, ++c) {
      const uint32_t st = c % kStages;
      mbar_wait(&sm.w_full[st], (c / kStages) & 1);
      tcgen05_fence_after();
      const uint64_t bdesc = make_desc_sw128(w_ring + st * kBlockBytes);
      const uint32_t dcol = tmem + d.tmem_col + nh * 128;
      switch (d.ksteps[kb]) {
        case 1: mma_chunk<1>(dcol, adesc, bdesc, idesc, kb != 0); break;
        case 2: mma_chunk<2>(dcol, adesc, bdesc, idesc, kb != 0); break;
        case 3: mma_chunk<3>(dcol, adesc, bdesc, idesc, kb != 0); break;
        default: mma_chunk<4>(dcol, adesc, bdesc, idesc, kb != 0); break;
      }
      mma_commit(&sm.w_empty[st]);
    }
  }
  mma_commit(&sm.acc_ready[L]);
}

__device__ __forceinline__ void wait_then_fence(uint64_t* bar, uint32_t parity) {
  mbar_wait(bar, parity);
  tcgen05_fence_after();
}

__global__ void __launch_bounds__(kThreads2, 1) k_nerf_mlp_fwd2(const float* __restrict__ features,
                                                                const float* __restrict__ viewdirs, int M,
                                                                int rows_per_ray, const uint8_t* __restrict__ blob,
                                                                float* __restrict__ o_density, float* __restrict__ o_rgb,
                                                                float* __restrict__ o_sem, float* __restrict__ o_int,
                                                                nlb_nerf_mlp_saved_t sv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_blocks = base;                                  // 9 x 16 KB
  uint8_t* w_ring = base + kNumABlocks * kBlockBytes;        // 4 x 16 KB
  Smem2& sm = *reinterpret_cast<Smem2*>(w_ring + kStages * kBlockBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    for (int i = 0; i < kNumLayers; ++i) mbar_init(&sm.acc_ready[i], 1);
    mbar_init(&sm.f_ready, kFrontThreads); mbar_init(&sm.h0_ready, kFrontThreads);
    mbar_init(&sm.x_ready, kFrontThreads); mbar_init(&sm.hs1_free, kFrontThreads);
    mbar_init(&sm.g_ready, kMainThreads); mbar_init(&sm.h1_ready, kMainThreads); mbar_init(&sm.h2_ready, kMainThreads);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kBiasFloats; i += kThreads2)
    sm.bias[i] = reinterpret_cast<const float*>(blob + kWeightBytes)[i];
  if (warp == kMmaWarp2) tmem_alloc(&sm.tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (g_timeline && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_timeline[120] = clock64(); g_timeline[122] = (long long)gt;
    g_timeline[124] = (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  }

  if (warp >= kProdWarp2) {
    // ===== weight producers: chunks in the MMA issuer's consumption order
    if (elect_one_sync()) {
      const int my_stage = warp - kProdWarp2;
      uint32_t c = 0;
      auto stream_layer = [&](int l) {
        const int n = layer_chunks(l), bytes = layer_chunk_bytes(l);
        const uint8_t* src = blob + layer_offset(l);
#pragma unroll 1
        for (int i = 0; i < n; ++i, ++c) {
          const int st = c % kStages;
          if (st != my_stage) continue;
          mbar_wait_relaxed(&sm.w_empty[st], ((c / kStages) & 1) ^ 1);
          mbar_expect_tx(&sm.w_full[st], bytes);
          bulk_g2s(w_ring + st * kBlockBytes, src + (size_t)i * bytes, bytes, &sm.w_full[st]);
        }
      };
      if ((int)blockIdx.x < num_tiles) { stream_layer(L0); stream_layer(L1); }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool has_next = tile + (int)gridDim.x < num_tiles;
        stream_layer(HS0); stream_layer(V0); stream_layer(HS1); stream_layer(V1);
        if (has_next) { stream_layer(L0); stream_layer(L1); }
        stream_layer(RGB);
      }
    }
  } else if (warp == kMmaWarp2) {
    // ===== MMA issuer
    if (elect_one_sync()) {
      uint32_t c = 0, it = 0;
      if ((int)blockIdx.x < num_tiles) {
        wait_then_fence(&sm.f_ready, 0);  issue_layer2<L0>(sm, a_blocks, w_ring, tmem, c);
        wait_then_fence(&sm.h0_ready, 0); issue_layer2<L1>(sm, a_blocks, w_ring, tmem, c);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = it & 1, nph = ph ^ 1;
        const bool has_next = tile + (int)gridDim.x < num_tiles;
        const int tb = (it == 2 || it == 3) ? ((int)it - 2) * 32 : -1;   // dev timeline (nlb_debug_set_timeline)
#define NLB_TS(j) do { if (tb >= 0) stamp(tb + (j)); } while (0)
        wait_then_fence(&sm.x_ready, ph); NLB_TS(0);
        issue_layer2<HS0>(sm, a_blocks, w_ring, tmem, c); NLB_TS(1);
        issue_layer2<V0>(sm, a_blocks, w_ring, tmem, c); NLB_TS(2);
        wait_then_fence(&sm.g_ready, ph); NLB_TS(3);
        issue_layer2<HS1>(sm, a_blocks, w_ring, tmem, c); NLB_TS(4);
        // V1 overwrites columns 0-255: the HS0 accumulator was read before g_ready / h1_ready, the HS1
        // accumulator (128-159) is read by the front warps -> hs1_free
        wait_then_fence(&sm.h1_ready, ph); NLB_TS(5);
        wait_then_fence(&sm.hs1_free, ph); NLB_TS(6);
        issue_layer2<V1>(sm, a_blocks, w_ring, tmem, c); NLB_TS(7);
        if (has_next) {
          // trunk of the next tile (X blocks and columns 256-511 are free: V0 was read before h1_ready,
          // f_ready is only signalled after this tile's V1 MMAs have completed)
          wait_then_fence(&sm.f_ready, nph); NLB_TS(8);
          issue_layer2<L0>(sm, a_blocks, w_ring, tmem, c); NLB_TS(9);
          wait_then_fence(&sm.h0_ready, nph); NLB_TS(10);
          issue_layer2<L1>(sm, a_blocks, w_ring, tmem, c); NLB_TS(11);
        }
        wait_then_fence(&sm.h2_ready, ph); NLB_TS(12);
        issue_layer2<RGB>(sm, a_blocks, w_ring, tmem, c); NLB_TS(13);
#undef NLB_TS
      }
    }
  } else if (warp >= kFrontWarp0) {
    // ===== front warps: thread (q, lane) owns row r = 32 q + lane of the tile it prepares
    const int q = warp - kFrontWarp0;
    const int r = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* XB = a_blocks + BX * kBlockBytes;
    uint8_t* DB = a_blocks + BD * kBlockBytes;
    uint32_t k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
      const uint32_t ph = k & 1;
      const int row = tile * 128 + r;
      const bool valid = row < M;
      const int rows_valid = M - tile * 128 < 128 ? M - tile * 128 : 128;
      const int tb = ((k == 2 || k == 3) && threadIdx.x == kFrontWarp0 * 32) ? 64 + ((int)k - 2) * 16 : -1;
#define NLB_TS(j) do { if (tb >= 0) stamp(tb + (j)); } while (0)
      NLB_TS(0);
      // ---- inputs of this tile into registers (bf16 pairs), before waiting for the blocks
      uint32_t fp[kFeat / 2], dp[16];
      if (valid) {
        const float4* src = reinterpret_cast<const float4*>(features + (size_t)row * kFeat);
#pragma unroll
        for (int i = 0; i < kFeat / 4; ++i) {
          const float4 t = __ldg(src + i);
          fp[2 * i] = pack_bf16(t.x, t.y); fp[2 * i + 1] = pack_bf16(t.z, t.w);
        }
        const int ray = row / rows_per_ray;
        const float vx = __ldg(viewdirs + 3 * ray), vy = __ldg(viewdirs + 3 * ray + 1), vz = __ldg(viewdirs + 3 * ray + 2);
        float d[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) d[i] = 0.f;
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
#pragma unroll
        for (int i = 0; i < 16; ++i) dp[i] = pack_bf16(d[2 * i], d[2 * i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < kFeat / 2; ++i) fp[i] = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) dp[i] = 0u;
      }
      // X and D are read by the previous tile's HS0 / V0 / V1 MMAs: free once V1 has completed
      NLB_TS(1);
      if (k > 0) mbar_wait_warp(&sm.acc_ready[V1], (k - 1) & 1);
      NLB_TS(2);
      {
        uint8_t* frow = block_row(XB, r);
        uint8_t* drow = block_row(DB, r);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 u = make_uint4(0, 0, 0, 0), w = make_uint4(0, 0, 0, 0);
          if (c8 < kFeat / 8) u = make_uint4(fp[c8 * 4], fp[c8 * 4 + 1], fp[c8 * 4 + 2], fp[c8 * 4 + 3]);
          if (c8 < 4) w = make_uint4(dp[c8 * 4], dp[c8 * 4 + 1], dp[c8 * 4 + 2], dp[c8 * 4 + 3]);
          *reinterpret_cast<uint4*>(frow + ((c8 ^ (r & 7)) * 16)) = u;
          *reinterpret_cast<uint4*>(drow + ((c8 ^ (r & 7)) * 16)) = w;
        }
      }
      signal_a_ready(&sm.f_ready); NLB_TS(3);
      // Global copies for the backward pass trail the signals: the blocks stay valid until this same warp
      // overwrites its rows (program order), so the stores run while the tensor pipe works.
      // bf16 copy of the (zero-padded) feature rows: the B operand of density_layer.0's weight gradient
      if (sv.f0) warp_rows_to_global(XB, r & ~31, lane, reinterpret_cast<__nv_bfloat16*>(sv.f0) + (size_t)tile * 128 * 64, 64, rows_valid);

      // ---- L0: h0 = relu(acc + b) -> X1
      mbar_wait_warp(&sm.acc_ready[L0], ph); NLB_TS(4);
      tcgen05_fence_after();
      epi_group64<true>(tlane + 256, sm.bias + bias_offset(L0), XB + 1 * kBlockBytes, r, lane, nullptr, 64, rows_valid);
      signal_a_ready(&sm.h0_ready); NLB_TS(5);
      if (sv.h0) warp_rows_to_global(XB + 1 * kBlockBytes, r & ~31, lane, reinterpret_cast<__nv_bfloat16*>(sv.h0) + (size_t)tile * 128 * 64, 64, rows_valid);

      // ---- L1: x = acc + b -> X0..3 (the L0 / L1 MMAs, readers of X0 / X1, have completed); density
      mbar_wait_warp(&sm.acc_ready[L1], ph); NLB_TS(6);
      tcgen05_fence_after();
      float x0 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64)
        epi_group64<false>(tlane + 256 + c0, sm.bias + bias_offset(L1) + c0, XB + (c0 >> 6) * kBlockBytes, r, lane,
                           nullptr, 256, rows_valid, c0 == 0 ? &x0 : nullptr);
      signal_a_ready(&sm.x_ready); NLB_TS(7);
      if (sv.x) {
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 64)
          warp_rows_to_global(XB + (c0 >> 6) * kBlockBytes, r & ~31, lane,
                              reinterpret_cast<__nv_bfloat16*>(sv.x) + (size_t)tile * 128 * 256 + c0, 256, rows_valid);
      }
      if (valid) {
        const float xin = x0 - 1.0f;
        o_density[row] = xin > 20.f ? xin : log1pf(expf(xin));
      }

      NLB_TS(8);
      // ---- HS1 of this tile (the main warps are in its view branch): semantic softmax (19) + intensity
      mbar_wait_warp(&sm.acc_ready[HS1], ph); NLB_TS(9);
      tcgen05_fence_after();
      {
        float v[32];
        tmem_ld32(tlane + 128, v);
        tcgen05_fence_before();
        mbar_arrive(&sm.hs1_free);
        const float* b = sm.bias + bias_offset(HS1);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kSem; ++i) { v[i] += b[i]; mx = fmaxf(mx, v[i]); }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kSem; ++i) { v[i] = expf(v[i] - mx); sum += v[i]; }
        const float inv = 1.0f / sum;
        if (valid) {
          if (o_sem) {
#pragma unroll
            for (int i = 0; i < kSem; ++i) o_sem[(size_t)row * kSem + i] = v[i] * inv;
          }
          if (o_int) o_int[row] = v[kSem] + b[kSem];
        }
      }
      NLB_TS(10);
#undef NLB_TS
    }
  } else {
    // ===== main epilogue warps 0..7: thread (q, lane) owns row r = 32 q + lane; the two warps of a
    // lane quarter (half = 0 / 1) split every layer's columns
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* HB = a_blocks + BH * kBlockBytes;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int row = tile * 128 + r;
      const bool valid = row < M;
      const int rows_valid = M - tile * 128 < 128 ? M - tile * 128 : 128;
      const int tb = ((it == 2 || it == 3) && threadIdx.x == 0) ? 96 + ((int)it - 2) * 12 : -1;
#define NLB_TS(j) do { if (tb >= 0) stamp(tb + (j)); } while (0)
      // ---- HS0: hidden = relu(acc + b) -> H2, H3  (H is free: every warp waited for the previous RGB)
      mbar_wait_warp(&sm.acc_ready[HS0], ph); NLB_TS(0);
      tcgen05_fence_after();
      {
        const int c0 = half * 64;
        epi_group64<true>(tlane + c0, sm.bias + bias_offset(HS0) + c0, HB + (2 + half) * kBlockBytes, r, lane, nullptr, 128,
                          rows_valid);
      }
      signal_a_ready(&sm.g_ready); NLB_TS(1);
      // (global copies trail the signals, see the front warps)
      if (sv.g) warp_rows_to_global(HB + (2 + half) * kBlockBytes, r & ~31, lane,
                                    reinterpret_cast<__nv_bfloat16*>(sv.g) + (size_t)tile * 128 * 128 + half * 64, 128, rows_valid);

      // ---- V0: h1 = relu(acc + b) -> H0..3 (overwrites H2 / H3, which the HS1 MMAs read)
      mbar_wait_warp(&sm.acc_ready[HS1], ph); NLB_TS(2);
      mbar_wait_warp(&sm.acc_ready[V0], ph); NLB_TS(3);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64<true>(tlane + 256 + c0, sm.bias + bias_offset(V0) + c0, HB + hperm(c0 >> 6) * kBlockBytes, r, lane,
                          nullptr, 256, rows_valid);
      signal_a_ready(&sm.h1_ready); NLB_TS(4);
      if (sv.h1) {
#pragma unroll 1
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
          warp_rows_to_global(HB + hperm(c0 >> 6) * kBlockBytes, r & ~31, lane,
                              reinterpret_cast<__nv_bfloat16*>(sv.h1) + (size_t)tile * 128 * 256 + c0, 256, rows_valid);
      }

      NLB_TS(5);
      // ---- V1: h2 = relu(acc + b) -> H0..3
      mbar_wait_warp(&sm.acc_ready[V1], ph); NLB_TS(6);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64<true>(tlane + c0, sm.bias + bias_offset(V1) + c0, HB + hperm(c0 >> 6) * kBlockBytes, r, lane,
                          nullptr, 256, rows_valid);
      signal_a_ready(&sm.h2_ready); NLB_TS(7);
      if (sv.h2) {
#pragma unroll 1
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
          warp_rows_to_global(HB + hperm(c0 >> 6) * kBlockBytes, r & ~31, lane,
                              reinterpret_cast<__nv_bfloat16*>(sv.h2) + (size_t)tile * 128 * 256 + c0, 256, rows_valid);
      }

      NLB_TS(8);
      // ---- RGB: sigmoid(acc + b) * (1 + 2 pad) - pad   (every warp waits: the next HS0 epilogue
      // overwrites H2 / H3, which the RGB MMAs read)
      mbar_wait_warp(&sm.acc_ready[RGB], ph); NLB_TS(9);
      tcgen05_fence_after();
      if (half == 1) {
        float v[16];
        tmem_ld16(tlane + 224, v);
        const float* b = sm.bias + bias_offset(RGB);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float s = 1.0f / (1.0f + expf(-(v[i] + b[i])));
            o_rgb[(size_t)row * 3 + i] = s * (1.0f + 2.0f * 0.001f) - 0.001f;
          }
        }
      }
      tcgen05_fence_before();
      NLB_TS(10);
#undef NLB_TS
    }
  }
  __syncthreads();
  if (g_timeline && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_timeline[121] = clock64(); g_timeline[123] = (long long)gt;
  }
  if (warp == kMmaWarp2) tmem_dealloc(tmem, 512);
}

constexpr size_t kSmemBytes2 = 1024 + (kNumABlocks + kStages) * (size_t)kBlockBytes + sizeof(Smem2);

}  // namespace v2

// =============================================================================
// Backward data-gradient chain (same machinery, transposed weights).
//   dc   = g_rgb * 1.002 * s(1-s)                       [16]   (S block)
//   dzv1 = (dc  . Wrgb)        * [h2 > 0]               [256]  (P blocks)
//   dh1 | dxa = dzv1 . Wv1[:, h1 | x]                   [512]
//   dzv0 = dh1 * [h1 > 0]                               [256]  (Q blocks)
//   dx  += dzv0 . Wv0[:, x]
//   dhs1 = softmax' (g_sem) | g_int                     [32]   (S block)
//   dzg  = (dhs1 . Whs1) * [g > 0]                      [128]  (Q0,Q1)
//   dx  += dzg . Whs0 ;  dx[0] += g_density * sigmoid(x0 - 1)  (P blocks)
//   dz0  = (dx . W1) * [h0 > 0]                         [64]   (Q2)
//   df   = dz0 . W0                                     [40]   -> grad_features
// Every dZ is also written row-major bf16 for the weight-gradient GEMMs.
namespace bwd {

constexpr int BP = 0, BQ = 4, BS = 8;
enum { B_RGB = 0, B_V1, B_V0, B_HS1, B_HS0, B_L1, B_L0, kNumBLayers };
struct BLayerDef {
  int N, nkb;
  int a_blk[4];
  int ksteps[4];
  int tmem_col[4];  // per 128-column part of N
  bool accumulate;  // add onto the existing accumulator (dx gathers three terms)
};
__host__ __device__ constexpr BLayerDef blayer_def(int l) {
  switch (l) {
    case B_RGB: return {256, 1, {BS}, {1}, {256, 384, 0, 0}, false};
    // 256-wide quantities staged in P (dzv1, dx): column block j lives in P block pperm(j) = {0, 2, 1, 3}[j], so
    // the warps of column half h only ever touch P(h) and P(h + 2) -- the g mask / dzg (P(h)) included -- and
    // their global saves can trail the "ready" signal with program order as the only hazard protection
    case B_V1:  return {512, 4, {BP, BP + 2, BP + 1, BP + 3}, {4, 4, 4, 4}, {256, 384, 0, 128}, false};
    case B_V0:  return {256, 4, {BQ, BQ + 1, BQ + 2, BQ + 3}, {4, 4, 4, 4}, {0, 128, 0, 0}, true};
    case B_HS1: return {128, 1, {BS}, {2}, {256, 0, 0, 0}, false};
    case B_HS0: return {256, 2, {BP, BP + 1}, {4, 4}, {0, 128, 0, 0}, true};   // dzg lives in P0,P1 (free after B_V1)
    case B_L1:  return {64, 4, {BP, BP + 2, BP + 1, BP + 3}, {4, 4, 4, 4}, {256, 0, 0, 0}, false};
    default:    return {48, 1, {BS}, {4}, {320, 0, 0, 0}, false};                  // dz0 lives in S (free after B_HS1)
  }
}
__host__ __device__ constexpr int pperm(int j) { return j == 1 ? 2 : (j == 2 ? 1 : j); }
__host__ __device__ constexpr int bl_nrows(int l) { return blayer_def(l).N > 128 ? 128 : blayer_def(l).N; }
__host__ __device__ constexpr int bl_nparts(int l) { return (blayer_def(l).N + 127) / 128; }
__host__ __device__ constexpr int bl_chunks(int l) { return blayer_def(l).nkb * bl_nparts(l); }
__host__ __device__ constexpr int bl_chunk_bytes(int l) { return bl_nrows(l) * 128; }
__host__ __device__ constexpr int bl_offset(int l) {
  int o = 0;
  for (int i = 0; i < l; ++i) o += bl_chunks(i) * bl_chunk_bytes(i);
  return o;
}
constexpr int kPackedTBytes = bl_offset(kNumBLayers);

// B operand of backward layer l: row n = index of the layer INPUT being differentiated,
// column k = index of the pre-activation gradient being contracted
__device__ float packed_weight_t(const nlb_nerf_mlp_weights_t& w, int l, int n, int k) {
  switch (l) {
    case B_RGB: return k < 3 ? w.W_rgb[k * 256 + n] : 0.f;                       // n: h2 unit
    case B_V1:  return w.W_v1[k * 539 + n];                                       // n < 512: h1 | x
    case B_V0:  return w.W_v0[k * 283 + n];                                       // n < 256: x
    case B_HS1:                                                                   // n: hidden unit (128)
      if (n < 64) return k < kSem ? w.W_s2[k * 64 + n] : 0.f;
      return k == kSem ? w.W_i2[n - 64] : 0.f;
    case B_HS0: return k < 64 ? w.W_s0[k * 256 + n] : w.W_i0[(k - 64) * 256 + n]; // n: x col
    case B_L1:  return w.W_d2[k * 64 + n];                                        // n: h0 unit
    default:    return n < kFeat ? w.W_d0[k * kFeat + n] : 0.f;                   // n: feature
  }
}

__global__ void k_pack_t(nlb_nerf_mlp_weights_t w, uint8_t* __restrict__ blob) {
  for (int l = 0; l < kNumBLayers; ++l) {
    const BLayerDef d = blayer_def(l);
    const int nrows = bl_nrows(l), np_count = bl_nparts(l);
    const int total = d.nkb * np_count * nrows * 64;
    const int kvalid = (l == B_RGB) ? 16 : (l == B_HS1 ? 32 : 64 * d.nkb);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int c = e & 63;
      const int r = (e >> 6) % nrows;
      const int chunk = (e >> 6) / nrows;  // kb-major, n-part minor
      const int kb = chunk / np_count, np = chunk % np_count;
      const int k = kb * 64 + c;
      const float v = k < kvalid ? packed_weight_t(w, l, np * 128 + r, k) : 0.f;
      uint8_t* dst = blob + bl_offset(l) + (size_t)chunk * bl_chunk_bytes(l) + sw128_offset(r, c);
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(v);
    }
  }
}

struct BSmem {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t acc_ready[kNumBLayers];
  uint64_t a_ready[kNumBLayers];
  uint32_t tmem_base;
};

// 32 accumulator columns -> (optional) ReLU mask from a saved bf16 activation -> block + global.
// `act_row` / `gdst` always point at dereferenceable memory (the caller clamps the row
// index); `write_g` predicates the global store.
template <bool kMask>
__device__ __forceinline__ void epi_masked(uint32_t taddr, const __nv_bfloat16* __restrict__ act_row, uint8_t* block,
                                           int row, int col_in_block, __nv_bfloat16* gdst, bool write_g,
                                           float add0 = 0.f) {
  uint4 m[4];
  if (kMask) {
    const uint4* a4 = reinterpret_cast<const uint4*>(act_row);
#pragma unroll
    for (int q = 0; q < 4; ++q) m[q] = __ldg(a4 + q);
  }
  float v[32];
  tmem_ld32(taddr, v);
  v[0] += add0;
  if (kMask) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // post-ReLU activations are >= 0: a unit was active iff its bf16 bits are non-zero (ignoring -0)
      if ((m[q].x & 0x7FFFu) == 0) v[q * 8 + 0] = 0.f;
      if ((m[q].x & 0x7FFF0000u) == 0) v[q * 8 + 1] = 0.f;
      if ((m[q].y & 0x7FFFu) == 0) v[q * 8 + 2] = 0.f;
      if ((m[q].y & 0x7FFF0000u) == 0) v[q * 8 + 3] = 0.f;
      if ((m[q].z & 0x7FFFu) == 0) v[q * 8 + 4] = 0.f;
      if ((m[q].z & 0x7FFF0000u) == 0) v[q * 8 + 5] = 0.f;
      if ((m[q].w & 0x7FFFu) == 0) v[q * 8 + 6] = 0.f;
      if ((m[q].w & 0x7FFF0000u) == 0) v[q * 8 + 7] = 0.f;
    }
  }
  uint8_t* rowp = block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((col_in_block >> 3) + q) ^ (row & 7);
    uint4 u;
    u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
    if (write_g) reinterpret_cast<uint4*>(gdst)[q] = u;
  }
}

// MMAs of backward layer L for one tile (see mlp::issue_layer).
template <int L, class SmemT>
__device__ __forceinline__ void issue_blayer(SmemT& sm, uint8_t* a_blocks, uint8_t* w_ring, uint32_t tmem, uint32_t& c,
                                             uint32_t ph) {
  constexpr BLayerDef d = blayer_def(L);
  constexpr int nrows = bl_nrows(L), np_count = bl_nparts(L);
  mbar_wait(&sm.a_ready[L], ph);
  tcgen05_fence_after();
  const uint32_t idesc = make_idesc_bf16(128, nrows);
#pragma unroll
  for (int kb = 0; kb < d.nkb; ++kb) {
    const uint64_t adesc = make_desc_sw128(a_blocks + d.a_blk[kb] * kBlockBytes);
#pragma unroll
    for (int np = 0; np < np_count; ++np, ++c) {
      const uint32_t st = c % kStages;
      mbar_wait(&sm.w_full[st], (c / kStages) & 1);
      tcgen05_fence_after();
      const uint64_t bdesc = make_desc_sw128(w_ring + st * kBlockBytes);
      const uint32_t dcol = tmem + d.tmem_col[np];
      const bool acc0 = d.accumulate || kb != 0;
      switch (d.ksteps[kb]) {
        case 1: mma_chunk<1>(dcol, adesc, bdesc, idesc, acc0); break;
        case 2: mma_chunk<2>(dcol, adesc, bdesc, idesc, acc0); break;
        case 3: mma_chunk<3>(dcol, adesc, bdesc, idesc, acc0); break;
        default: mma_chunk<4>(dcol, adesc, bdesc, idesc, acc0); break;
      }
      mma_commit(&sm.w_empty[st]);
    }
  }
  mma_commit(&sm.acc_ready[L]);
}

// backward group: 64 accumulator columns of this thread's row -> (optional) ReLU mask from
// the saved bf16 activation -> block + row-major bf16 copy for the weight-gradient GEMMs.
// The activation rows are first staged, coalesced, into the destination block itself (it
// is free: it is about to be overwritten), each thread then reads its own row from there.
//   act / gdst = tile base + column offset of the saved activation / output matrices.
//   kPrefetched: the activation rows are already in the block (warp_rows_from_global_async at the start
//   of the tile, completed by the caller) -- the global-load latency of the two widest epilogues is then
//   hidden behind the first GEMMs instead of being paid per group on the tile's critical path.
template <bool kMask, bool kPrefetched = false>
__device__ __forceinline__ void epi_group64_masked(uint32_t taddr, const __nv_bfloat16* __restrict__ act,
                                                   uint8_t* block, int row, int lane, __nv_bfloat16* gdst, int ld_act,
                                                   int ld_out, int rows_valid, float add0 = 0.f) {
  if (kMask && !kPrefetched) warp_rows_from_global(block, row & ~31, lane, act, ld_act, rows_valid);
  float v[64];
  tmem_ld64(taddr, v);
  v[0] += add0;
  if (kMask) {
    const uint8_t* rowp = block_row(block, row);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 m = *reinterpret_cast<const uint4*>(rowp + ((q ^ (row & 7)) * 16));
      // post-ReLU activations are >= 0: a unit was active iff its bf16 bits are non-zero (ignoring -0)
      if ((m.x & 0x7FFFu) == 0) v[q * 8 + 0] = 0.f;
      if ((m.x & 0x7FFF0000u) == 0) v[q * 8 + 1] = 0.f;
      if ((m.y & 0x7FFFu) == 0) v[q * 8 + 2] = 0.f;
      if ((m.y & 0x7FFF0000u) == 0) v[q * 8 + 3] = 0.f;
      if ((m.z & 0x7FFFu) == 0) v[q * 8 + 4] = 0.f;
      if ((m.z & 0x7FFF0000u) == 0) v[q * 8 + 5] = 0.f;
      if ((m.w & 0x7FFFu) == 0) v[q * 8 + 6] = 0.f;
      if ((m.w & 0x7FFF0000u) == 0) v[q * 8 + 7] = 0.f;
    }
  }
  store_row64(v, block, row);
  if (gdst) warp_rows_to_global(block, row & ~31, lane, gdst, ld_out, rows_valid);
}

__global__ void __launch_bounds__(kThreads, 1) k_nerf_mlp_bwd(nlb_nerf_mlp_grad_in_t gi, nlb_nerf_mlp_saved_t sv, int M,
                                                              const uint8_t* __restrict__ blob,
                                                              float* __restrict__ grad_features,
                                                              nlb_nerf_mlp_grad_out_t go) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_blocks = base;
  uint8_t* w_ring = base + kNumABlocks * kBlockBytes;
  BSmem& sm = *reinterpret_cast<BSmem*>(w_ring + kStages * kBlockBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    for (int i = 0; i < kNumBLayers; ++i) { mbar_init(&sm.acc_ready[i], 1); mbar_init(&sm.a_ready[i], kEpiThreads); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp >= kProdWarp0) {
    if (elect_one_sync()) {
      const int my_stage = warp - kProdWarp0;
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int l = 0; l < kNumBLayers; ++l) {
          const int n = bl_chunks(l), bytes = bl_chunk_bytes(l);
          const uint8_t* src = blob + bl_offset(l);
#pragma unroll 1
          for (int i = 0; i < n; ++i, ++c) {
            const int st = c % kStages;
            if (st != my_stage) continue;
            mbar_wait_relaxed(&sm.w_empty[st], ((c / kStages) & 1) ^ 1);
            mbar_expect_tx(&sm.w_full[st], bytes);
            bulk_g2s(w_ring + st * kBlockBytes, src + (size_t)i * bytes, bytes, &sm.w_full[st]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one_sync()) {
      uint32_t c = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        issue_blayer<B_RGB>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_V1>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_V0>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_HS1>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_HS0>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_L1>(sm, a_blocks, w_ring, tmem, c, ph);
        issue_blayer<B_L0>(sm, a_blocks, w_ring, tmem, c, ph);
      }
    }
  } else {
    const int half = warp >> 2;  // the two warps of a TMEM lane quarter split every layer's columns
    const int ld_v1 = go.ld_v1 ? go.ld_v1 : 256, ld_v0 = go.ld_v0 ? go.ld_v0 : 256, ld_g = go.ld_g ? go.ld_g : 128;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* PB = a_blocks + BP * kBlockBytes;
    uint8_t* QB = a_blocks + BQ * kBlockBytes;
    uint8_t* SB = a_blocks + BS * kBlockBytes;
    auto bf = [](void* p) { return reinterpret_cast<__nv_bfloat16*>(p); };
    auto cbf = [](const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); };
    float dc_rgb[3] = {0.f, 0.f, 0.f}, dc_g[3] = {0.f, 0.f, 0.f};
    auto load_dc_inputs = [&](int row_, bool valid_) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        dc_rgb[i] = (valid_ && gi.g_rgb) ? __ldg(gi.rgb + (size_t)row_ * 3 + i) : 0.f;
        dc_g[i] = (valid_ && gi.g_rgb) ? __ldg(gi.g_rgb + (size_t)row_ * 3 + i) : 0.f;
      }
    };
    // this warp's 32 rows x this half's two 64-column blocks of a saved [M,256] activation -> blocks dst[0..3]
    auto prefetch_mask = [&](uint8_t* dst, const void* act, int tile_, bool permuted) {
      if (tile_ >= num_tiles) return;
      const int rv = M - tile_ * 128 < 128 ? M - tile_ * 128 : 128;
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        warp_rows_from_global_async(dst + (permuted ? pperm(c0 >> 6) : (c0 >> 6)) * kBlockBytes, r & ~31, lane,
                                    cbf(act) + (size_t)tile_ * 128 * 256 + c0, 256, rv);
    };
    // global copy of this warp's rows of a staged 64-column block (for the weight-gradient GEMMs); issued AFTER
    // the block's "ready" signal so that the stores run under the next GEMM
    auto save_block = [&](const uint8_t* block, void* g, size_t trow_, int ld, int c0, int rv) {
      if (g) warp_rows_to_global(block, r & ~31, lane, bf(g) + trow_ * ld + c0, ld, rv);
    };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int row = tile * 128 + r;
      const bool valid = row < M;
      const int rows_valid = M - tile * 128 < 128 ? M - tile * 128 : 128;
      const size_t trow = (size_t)tile * 128;  // first row of the tile
      uint8_t* srow = SB + (r >> 3) * 1024 + (r & 7) * 128;
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 0);
      // ---- dc -> S (cols 0..2): half 0 (inputs were loaded at the end of the previous tile)
      if (half == 0) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (it == 0) load_dc_inputs(row, valid);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float s = (dc_rgb[i] + 0.001f) / (1.0f + 2.0f * 0.001f);
          v[i] = dc_g[i] * (1.0f + 2.0f * 0.001f) * s * (1.0f - s);
        }
        store_cols(v, SB, r, 0, nullptr);
        if (valid && go.d_rgb) {
          uint4* dst = reinterpret_cast<uint4*>(bf(go.d_rgb) + (size_t)row * 16);
          dst[0] = *reinterpret_cast<uint4*>(srow + ((0 ^ (r & 7)) * 16));
          dst[1] = make_uint4(0, 0, 0, 0);
        }
      }
      signal_a_ready(&sm.a_ready[B_RGB]);

      // ReLU masks of the two widest epilogues: h2 -> P, h1 -> Q.  For the first tile they are fetched
      // here; for every later tile they were issued during the previous one, as soon as its last reader
      // of Q (B_V0) / P (B_L1) had completed, so they are resident when the tile starts.
      if (it == 0) {
        prefetch_mask(QB, sv.h1, tile, false);
        prefetch_mask(PB, sv.h2, tile, true);
      }

      // ---- dzv1 = dh2 * [h2 > 0] -> P0..3
      mbar_wait_warp(&sm.acc_ready[B_RGB], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 2);
      tcgen05_fence_after();
      cp_async_wait_all();
      __syncwarp();
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64_masked<true, true>(tlane + 256 + c0, cbf(sv.h2) + trow * 256 + c0, PB + pperm(c0 >> 6) * kBlockBytes, r, lane,
                                 nullptr, 256, ld_v1, rows_valid);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 3);
      signal_a_ready(&sm.a_ready[B_V1]);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        save_block(PB + pperm(c0 >> 6) * kBlockBytes, go.d_v1, trow, ld_v1, c0, rows_valid);

      // ---- d(sem logits) | d(intensity) -> S (cols 0..19): half 1, while B_V1 runs; S is free: B_RGB completed above
      if (half == 1) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        if (valid) {
          if (gi.g_semantic) {
            float p[kSem], g[kSem], dot = 0.f;
#pragma unroll
            for (int i = 0; i < kSem; ++i) {
              p[i] = __ldg(gi.semantic + (size_t)row * kSem + i);
              g[i] = __ldg(gi.g_semantic + (size_t)row * kSem + i);
              dot = fmaf(p[i], g[i], dot);
            }
#pragma unroll
            for (int i = 0; i < kSem; ++i) v[i] = p[i] * (g[i] - dot);
          }
          if (gi.g_intensity) v[kSem] = __ldg(gi.g_intensity + row);
        }
        store_cols(v, SB, r, 0, (valid && go.d_hs1) ? bf(go.d_hs1) + (size_t)row * 32 : nullptr);
      }
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 6);
      signal_a_ready(&sm.a_ready[B_HS1]);

      // ---- dzv0 = dh1 * [h1 > 0] -> Q0..3   (dh1 in accB, dx partial stays in accA)
      mbar_wait_warp(&sm.acc_ready[B_V1], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 4);
      tcgen05_fence_after();
      // B_V1 has consumed P: the ReLU mask of the sem | intensity hidden layer goes to P0 / P1 now and is
      // there when the HS1 epilogue needs it
      warp_rows_from_global_async(PB + half * kBlockBytes, r & ~31, lane, cbf(sv.g) + trow * 128 + half * 64, 128, rows_valid);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        epi_group64_masked<true, true>(tlane + 256 + c0, cbf(sv.h1) + trow * 256 + c0, QB + (c0 >> 6) * kBlockBytes, r, lane,
                                 nullptr, 256, ld_v0, rows_valid);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 5);
      signal_a_ready(&sm.a_ready[B_V0]);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        save_block(QB + (c0 >> 6) * kBlockBytes, go.d_v0, trow, ld_v0, c0, rows_valid);


      // ---- dzg = dg * [g > 0] -> P0,P1
      mbar_wait_warp(&sm.acc_ready[B_HS1], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 7);
      tcgen05_fence_after();
      cp_async_wait_all();
      __syncwarp();
      // B_HS1 has consumed S: the ReLU mask of density_layer.0 goes there for the last epilogue
      if (half == 0) warp_rows_from_global_async(SB, r & ~31, lane, cbf(sv.h0) + trow * 64, 64, rows_valid);
      // B_V0 (ahead of B_HS1 on the pipe) has consumed Q, which this tile does not touch again
      prefetch_mask(QB, sv.h1, tile + (int)gridDim.x, false);
      {
        const int c0 = half * 64;
        epi_group64_masked<true, true>(tlane + 256 + c0, cbf(sv.g) + trow * 128 + c0, PB + half * kBlockBytes, r, lane,
                                       nullptr, 128, ld_g, rows_valid);
      }
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 8);
      signal_a_ready(&sm.a_ready[B_HS0]);
      save_block(PB + half * kBlockBytes, go.d_g, trow, ld_g, half * 64, rows_valid);

      // ---- dx = accA (+ density term on column 0) -> P0..3
      mbar_wait_warp(&sm.acc_ready[B_HS0], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 9);
      tcgen05_fence_after();
      {
        float dterm = 0.f;
        if (valid && gi.g_density) dterm = __ldg(gi.g_density + row) * (1.0f - expf(-__ldg(gi.density + row)));
#pragma unroll 1
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
          epi_group64_masked<false>(tlane + c0, nullptr, PB + pperm(c0 >> 6) * kBlockBytes, r, lane,
                                    nullptr, 256, 256, rows_valid, c0 == 0 ? dterm : 0.f);
      }
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 10);
      signal_a_ready(&sm.a_ready[B_L1]);
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64)
        save_block(PB + pperm(c0 >> 6) * kBlockBytes, go.d_x, trow, 256, c0, rows_valid);

      // ---- dz0 = dh0 * [h0 > 0] -> S
      mbar_wait_warp(&sm.acc_ready[B_L1], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 11);
      tcgen05_fence_after();
      if (half == 0) {
        cp_async_wait_all();
        __syncwarp();
        epi_group64_masked<true, true>(tlane + 256, cbf(sv.h0) + trow * 64, SB, r, lane, nullptr, 64, 64, rows_valid);
      }
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 12);
      signal_a_ready(&sm.a_ready[B_L0]);
      if (half == 0) save_block(SB, go.d_h0, trow, 64, 0, rows_valid);
      // B_L1 has consumed P (dx): next tile's h2 mask
      prefetch_mask(PB, sv.h2, tile + (int)gridDim.x, true);

      // next tile's rgb / g_rgb (dc staging opens the tile with nothing to overlap their latency)
      if (half == 0) {
        const int nrow = (tile + (int)gridDim.x) * 128 + r;
        load_dc_inputs(nrow, nrow < M && tile + (int)gridDim.x < num_tiles);
      }
      // ---- grad_features = accB[64:112) (40 valid columns)
      mbar_wait_warp(&sm.acc_ready[B_L0], ph);
      if (it < 2 && warp == 0 && lane == 0) mlp::stamp((int)it * 20 + 13);
      tcgen05_fence_after();
      if (half == 0) {
        float v[32];
        tmem_ld32(tlane + 320, v);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(grad_features + (size_t)row * kFeat);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      } else {
        float t[16];
        tmem_ld16(tlane + 352, t);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(grad_features + (size_t)row * kFeat);
          dst[8] = make_float4(t[0], t[1], t[2], t[3]);
          dst[9] = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

constexpr size_t kBSmemBytes = 1024 + (kNumABlocks + kStages) * (size_t)kBlockBytes + sizeof(BSmem);

}  // namespace bwd

constexpr size_t kSmemBytes = 1024 + (kNumABlocks + kStages) * (size_t)kBlockBytes + sizeof(Smem);

}  // namespace mlp
}  // namespace nlb

using namespace nlb;

extern "C" size_t nlb_nerf_mlp_packed_bytes(void) { return mlp::kPackedBytes; }

extern "C" int nlb_nerf_mlp_pack(const nlb_nerf_mlp_weights_t* w, void* packed, void* stream) {
  if (!w || !packed) { nlb_set_error("nerf_mlp_pack: null pointer"); return NLB_EINVAL; }
  const void* const* p = reinterpret_cast<const void* const*>(w);
  for (size_t i = 0; i < sizeof(*w) / sizeof(void*); ++i) {
    // the intensity head is optional in the reference (Config.use_intensity)
    if (!p[i] && !(i >= 8 && i < 12)) { nlb_set_error("nerf_mlp_pack: null weight pointer %zu", i); return NLB_EINVAL; }
  }
  // Config.use_intensity=False: the intensity head's rows of the fused sem | intensity layers are packed as zeros
  // and nlb_nerf_mlp_forward is called with intensity == NULL (inference; the training kernels need the head)
  if ((!w->W_i0) != (!w->b_i0) || (!w->W_i0) != (!w->W_i2) || (!w->W_i0) != (!w->b_i2)) {
    nlb_set_error("nerf_mlp_pack: the intensity head must be given completely or not at all");
    return NLB_EINVAL;
  }
  mlp::k_pack<<<148, 256, 0, (cudaStream_t)stream>>>(*w, reinterpret_cast<uint8_t*>(packed));
  return nlb_check_launch("nerf_mlp_pack");
}

extern "C" int nlb_nerf_mlp_forward(const float* features, const float* viewdirs, int M, int rows_per_ray,
                                    const void* packed, float* density, float* rgb, float* semantic, float* intensity,
                                    const nlb_nerf_mlp_saved_t* saved, void* stream) {
  if (M == 0) return NLB_OK;
  if (!features || !viewdirs || !packed || !density || !rgb) { nlb_set_error("nerf_mlp_forward: null pointer"); return NLB_EINVAL; }
  if (rows_per_ray < 1) { nlb_set_error("nerf_mlp_forward: rows_per_ray must be >= 1"); return NLB_EINVAL; }
  if (reinterpret_cast<uintptr_t>(features) & 15 || reinterpret_cast<uintptr_t>(packed) & 15) {
    nlb_set_error("nerf_mlp_forward: features / packed weights must be 16-byte aligned");
    return NLB_EINVAL;
  }
  static bool attr_set[64] = {false};   // per device: function attributes are device state
  static const bool legacy = [] { const char* e = getenv("NLB_MLP_FWD_LEGACY"); return e && e[0] == '1'; }();  // A/B timing
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(mlp::k_nerf_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp::kSmemBytes);
    cudaFuncSetAttribute(mlp::v2::k_nerf_mlp_fwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp::v2::kSmemBytes2);
    attr_set[dev] = true;
  }
  const int sms = nlb_sm_count();
  const int tiles = (M + 127) / 128;
  const int grid = tiles < sms ? tiles : sms;
  nlb_nerf_mlp_saved_t sv = {};
  if (saved) sv = *saved;
  if (legacy)
    mlp::k_nerf_mlp_fwd<<<grid, mlp::kThreads, mlp::kSmemBytes, (cudaStream_t)stream>>>(
        features, viewdirs, M, rows_per_ray, reinterpret_cast<const uint8_t*>(packed), density, rgb, semantic, intensity, sv);
  else
    mlp::v2::k_nerf_mlp_fwd2<<<grid, mlp::v2::kThreads2, mlp::v2::kSmemBytes2, (cudaStream_t)stream>>>(
        features, viewdirs, M, rows_per_ray, reinterpret_cast<const uint8_t*>(packed), density, rgb, semantic, intensity, sv);
  return nlb_check_launch("nerf_mlp_forward");
}

extern "C" size_t nlb_nerf_mlp_packed_transposed_bytes(void) { return mlp::bwd::kPackedTBytes; }

extern "C" int nlb_nerf_mlp_pack_transposed(const nlb_nerf_mlp_weights_t* w, void* packed_t, void* stream) {
  if (!w || !packed_t) { nlb_set_error("nerf_mlp_pack_transposed: null pointer"); return NLB_EINVAL; }
  const void* const* p = reinterpret_cast<const void* const*>(w);
  for (size_t i = 0; i < sizeof(*w) / sizeof(void*); ++i)
    if (!p[i]) { nlb_set_error("nerf_mlp_pack_transposed: null weight pointer %zu", i); return NLB_EINVAL; }
  mlp::bwd::k_pack_t<<<148, 256, 0, (cudaStream_t)stream>>>(*w, reinterpret_cast<uint8_t*>(packed_t));
  return nlb_check_launch("nerf_mlp_pack_transposed");
}

extern "C" int nlb_nerf_mlp_backward(const nlb_nerf_mlp_grad_in_t* gin, const nlb_nerf_mlp_saved_t* saved, int M,
                                     const void* packed_t, float* grad_features, const nlb_nerf_mlp_grad_out_t* gout,
                                     void* stream) {
  if (M == 0) return NLB_OK;
  if (!gin || !saved || !packed_t || !grad_features || !gout) { nlb_set_error("nerf_mlp_backward: null pointer"); return NLB_EINVAL; }
  if (!saved->h0 || !saved->g || !saved->h1 || !saved->h2) { nlb_set_error("nerf_mlp_backward: the activations saved by the forward are required"); return NLB_EINVAL; }
  if (!gout->d_rgb || !gout->d_v1 || !gout->d_v0 || !gout->d_hs1 || !gout->d_g || !gout->d_x || !gout->d_h0) {
    nlb_set_error("nerf_mlp_backward: every pre-activation gradient buffer of nlb_nerf_mlp_grad_out_t is required");
    return NLB_EINVAL;
  }
  if ((gout->ld_v1 && gout->ld_v1 < 256) || (gout->ld_v0 && gout->ld_v0 < 256) || (gout->ld_g && gout->ld_g < 128)) {
    nlb_set_error("nerf_mlp_backward: leading dimensions smaller than the matrices");
    return NLB_EINVAL;
  }
  if ((gin->g_rgb && !gin->rgb) || (gin->g_semantic && !gin->semantic) || (gin->g_density && !gin->density)) {
    nlb_set_error("nerf_mlp_backward: forward outputs are required next to their gradients");
    return NLB_EINVAL;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(mlp::bwd::k_nerf_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp::bwd::kBSmemBytes);
    attr_set[dev] = true;
  }
  const int sms = nlb_sm_count();
  const int tiles = (M + 127) / 128;
  const int grid = tiles < sms ? tiles : sms;
  mlp::bwd::k_nerf_mlp_bwd<<<grid, mlp::kThreads, mlp::bwd::kBSmemBytes, (cudaStream_t)stream>>>(
      *gin, *saved, M, reinterpret_cast<const uint8_t*>(packed_t), grad_features, *gout);
  return nlb_check_launch("nerf_mlp_backward");
}

extern "C" int nlb_debug_set_timeline(void* buf) {
  long long* p = reinterpret_cast<long long*>(buf);
  cudaError_t e = cudaMemcpyToSymbol(nlb::mlp::g_timeline, &p, sizeof(p));
  return e == cudaSuccess ? 0 : NLB_ECUDA;
}
