// Stage-3 ray-drop U-Net, inference (`R/` = NeRF_LiDAR/NeRF_Lidar_code/): R/src/unet/unet_model.py:6-47,
// R/src/unet/unet_parts.py:8-77 -- DoubleConv (3x3 conv, no bias -> BatchNorm (eval) -> ReLU, twice), Down (2x2 max pool +
// DoubleConv), Up (bilinear x2 upsampling with align_corners or a 2x2 stride-2 transposed convolution, concatenation
// with the skip tensor, DoubleConv), OutConv (1x1 conv with bias) -- on the 32 x 1024 range-image features.
//
// fp32 throughout, as the reference computes it (the drop mask is a hard threshold on these logits).  The 3x3
// convolution is a register-tiled direct convolution: a block owns 64 output channels x (4 rows x 32 columns) of
// pixels, a thread 8 channels x 4 pixels; input channels are walked in chunks of 8 through shared memory; the
// concatenation of an Up block is never materialised (the kernel reads its input channels from two tensors); the
// folded BatchNorm scale / shift and the ReLU are the epilogue.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/nlb200.h"

namespace nlb {
namespace unet {

constexpr int kOcTile = 64, kRows = 4, kCols = 32, kIcChunk = 8;
constexpr int kConvThreads = 256;   // 32 pixel threads (4 px each) x 8 channel threads (8 oc each)

// out[n, oc, y, x] = act(scale[oc] * sum_{ic,ky,kx} in[n, ic, y+ky-1, x+kx-1] * W[oc, ic, ky, kx] + shift[oc])
// `in` = channels [0, CA) of inA followed by [0, CB) of inB (torch.cat([x2, x1], dim=1) of Up.forward).
__global__ void __launch_bounds__(kConvThreads) k_conv3x3(const float* __restrict__ inA, int CA, const float* __restrict__ inB,
                                                          int CB, const float* __restrict__ W, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int OC, int H, int Wd, int relu,
                                                          float* __restrict__ out, int ksplit, float* __restrict__ partial) {
  __shared__ float s_in[kIcChunk][kRows + 2][kCols + 2];
  // row stride 68: the staging stores walk (c, k) fastest (the weights' memory order) and would otherwise hit one bank
  __shared__ __align__(16) float s_w[kIcChunk][9][kOcTile + 4];
  const int C = CA + CB;
  const int tiles_x = (Wd + kCols - 1) / kCols;
  const int tile = blockIdx.x, tx0 = (tile % tiles_x) * kCols, ty0 = (tile / tiles_x) * kRows;
  const int oc0 = blockIdx.y * kOcTile, n = blockIdx.z / ksplit, ks = blockIdx.z - n * ksplit;
  // split over the input channels (ksplit > 1): this block sums the chunks [ic_begin, ic_end) into its own slice of
  // `partial`; k_conv_finish adds the slices in order and applies the epilogue
  const int chunks = (C + kIcChunk - 1) / kIcChunk, per = (chunks + ksplit - 1) / ksplit;
  const int ic_begin = ks * per * kIcChunk, ic_end = min(C, (ks + 1) * per * kIcChunk);
  const int pt = threadIdx.x & 31, ct = threadIdx.x >> 5;      // pixel thread, channel thread
  const int py = pt >> 3, px = (pt & 7) * 4;                    // this thread's 4 pixels: row py, columns px..px+3
  float acc[8][4];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
  const size_t plane = (size_t)H * Wd;
  for (int ic0 = ic_begin; ic0 < ic_end; ic0 += kIcChunk) {
    __syncthreads();
    for (int e = threadIdx.x; e < kIcChunk * (kRows + 2) * (kCols + 2); e += kConvThreads) {
      const int c = e / ((kRows + 2) * (kCols + 2)), r = (e / (kCols + 2)) % (kRows + 2), q = e % (kCols + 2);
      const int ic = ic0 + c, y = ty0 + r - 1, x = tx0 + q - 1;
      float v = 0.f;
      if (ic < ic_end && y >= 0 && y < H && x >= 0 && x < Wd) {
        const float* src = ic < CA ? inA + ((size_t)n * CA + ic) * plane : inB + ((size_t)n * CB + (ic - CA)) * plane;
        v = __ldg(src + (size_t)y * Wd + x);
      }
      s_in[c][r][q] = v;
    }
    // W[oc][ic0 .. ic0+8)[3][3] is 72 contiguous floats per output channel: consecutive threads read consecutive
    // addresses (with o fastest every thread touched its own 32-byte sector: 147 KB of L2 sectors per chunk and
    // block, ~9 TB/s over the chip at the FMA rate -- the kernel was L2-bound on its own weights)
    for (int e = threadIdx.x; e < kIcChunk * 9 * kOcTile; e += kConvThreads) {
      const int o = e / (kIcChunk * 9), r = e - o * (kIcChunk * 9);
      const int c = r / 9, k = r - c * 9;
      const int ic = ic0 + c, oc = oc0 + o;
      s_w[c][k][o] = (ic < ic_end && oc < OC) ? __ldg(W + ((size_t)oc * C + ic0) * 9 + r) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < kIcChunk; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float v[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] = s_in[c][py + ky][px + q];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][ky * 3 + kx][ct * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][ky * 3 + kx][ct * 8 + 4]);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int o = 0; o < 8; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = fmaf(w[o], v[p + kx], acc[o][p]);
        }
      }
    }
  }
  const int y = ty0 + py;
  if (y >= H) return;
  if (ksplit > 1) {
    float* base = partial + (size_t)ks * (gridDim.z / ksplit) * OC * plane;   // slice ks of [ksplit][N, OC, H, W]
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int oc = oc0 + ct * 8 + o;
      if (oc >= OC) continue;
      float* dst = base + (((size_t)n * OC + oc) * H + y) * Wd + tx0 + px;
      if (tx0 + px + 3 < Wd && (Wd & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (tx0 + px + p < Wd) dst[p] = acc[o][p];
      }
    }
    return;
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int oc = oc0 + ct * 8 + o;
    if (oc >= OC) continue;
    const float sc = scale ? __ldg(scale + oc) : 1.f, sh = shift ? __ldg(shift + oc) : 0.f;
    float* dst = out + (((size_t)n * OC + oc) * H + y) * Wd + tx0 + px;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (tx0 + px + p < Wd) {
        const float r = fmaf(acc[o][p], sc, sh);
        dst[p] = relu ? fmaxf(r, 0.f) : r;
      }
    }
  }
}

// ----------------------------------------------------------------------------- 3x3 convolution on tcgen05 (TF32)
// torch's default for convolutions is TF32 (torch.backends.cudnn.allow_tf32 = True), so what the reference's U-Net
// computes on a GPU is this arithmetic: fp32 storage, operands read with a 10-bit mantissa, fp32 accumulation.
// Implicit GEMM per CTA: M = 128 pixels (1 x 128 or 2 x 64 of the image), N = NT output channels, K = 32 input
// channels x 9 taps per chunk.  One pipeline stage = one (chunk, tap): A[128 px][32 ch] built by the 128 staging
// threads from a raw halo tile ([32][R+2][CW+2] floats, fetched one chunk ahead with 4-byte cp.async, zero fill at the
// borders), B[NT][32] = a pre-packed, pre-swizzled block of the weights fetched with one bulk copy; both are
// SWIZZLE_128B K-major operand blocks (32 fp32 = one 128-byte row); 4 tcgen05.mma (K = 8) per stage, accumulator in
// TMEM.  Warps 0-3: staging + epilogue (folded BatchNorm + ReLU, or a raw partial slice when the input channels are
// split over CTAs), warp 4: MMA issue, warps 5-7: weight copies, one warp per ring slot (a bulk copy keeps its
// issuing thread busy for ~800 cycles whatever its size -- tools/bulk_probe.cu -- against 256 cycles of MMA per
// stage: with ONE copy thread the ring ran at the copy-issue rate, 3.7 us per chunk).
constexpr int kTcM = 128, kTcKc = 32, kTcStages = 3;
constexpr int kTcThreads = 128 + 32 + 32 * kTcStages;
constexpr int kTcABytes = kTcM * kTcKc * 4;   // 16 KB

template <int NT, int CW>
struct TcShape {
  static constexpr int R = kTcM / CW;
  static constexpr int kPlane = (R + 2) * (CW + 2);      // floats per channel of the raw tile
  static constexpr int kRawFloats = kTcKc * kPlane;
  static constexpr int kBBytes = NT * kTcKc * 4;
  static constexpr int kStageBytes = kTcABytes + kBBytes;
  static constexpr size_t kSmem = 1024 + (size_t)kTcStages * kStageBytes + 2 * (size_t)kRawFloats * 4 + 256;
  static constexpr int kRawPer = (kPlane + 127) / 128;   // raw-tile elements per staging thread and channel
};

__host__ __device__ inline int packed_offset_floats(int o, int c) {   // (row, channel) inside a [NT][32] fp32 block
  return (o >> 3) * 256 + (o & 7) * 32 + (((c >> 2) ^ (o & 7)) << 2) + (c & 3);
}

// W[OC][C][3][3] -> blocks [oc tile][chunk][tap][NT x 32] in operand layout
__global__ void k_pack_conv_weights(const float* __restrict__ W, int OC, int C, int NT, float* __restrict__ packed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)OC * C * 9;
  if (i >= total) return;
  const int k = (int)(i % 9), ic = (int)((i / 9) % C), oc = (int)(i / ((size_t)9 * C));
  const int chunks = C / kTcKc;
  const size_t block = ((size_t)(oc / NT) * chunks + ic / kTcKc) * 9 + k;
  packed[block * ((size_t)NT * kTcKc) + packed_offset_floats(oc % NT, ic % kTcKc)] = W[i];
}

// W[C][OC][2][2] (ConvTranspose2d) -> blocks [virtual-channel tile][chunk][NT x 32], v = (dy * 2 + dx) * OC + oc
__global__ void k_pack_convT_weights(const float* __restrict__ W, int C, int OC, int NT, float* __restrict__ packed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)C * OC * 4;
  if (i >= total) return;
  const int t4 = (int)(i % 4), oc = (int)((i / 4) % OC), ic = (int)(i / ((size_t)4 * OC));
  const int v = t4 * OC + oc, chunks = C / kTcKc;
  const size_t block = (size_t)(v / NT) * chunks + ic / kTcKc;
  packed[block * ((size_t)NT * kTcKc) + packed_offset_floats(v % NT, ic % kTcKc)] = W[i];
}

// kTransposed: ConvTranspose2d(kernel 2, stride 2) as the same implicit GEMM with ONE tap (the centre) and 4 x oc_real
// virtual output channels v = (dy * 2 + dx) * oc_real + oc, scattered to out[oc, 2y + dy, 2x + dx] + bias (`shift`).
template <int NT, int CW, bool kTransposed>
__global__ void __launch_bounds__(kTcThreads, 1) k_conv3x3_tf32(const float* __restrict__ inA, int CA,
                                                                const float* __restrict__ inB, int CB,
                                                                const float* __restrict__ packed,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift, int OC, int H, int Wd,
                                                                float* __restrict__ out, int ksplit,
                                                                float* __restrict__ partial, int oc_real) {
  using namespace umma;
  using Sh = TcShape<NT, CW>;
  constexpr int kTaps = kTransposed ? 1 : 9;
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw_ + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = base;
  float* raw = reinterpret_cast<float*>(base + kTcStages * Sh::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 2 * Sh::kRawFloats);
  uint64_t* full = bars;                 // [kTcStages]: 128 staging arrivals + the weight copy's bytes
  uint64_t* empty = bars + kTcStages;    // [kTcStages]: the stage's MMAs have completed
  uint64_t* acc_bar = bars + 2 * kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 1);

  const int C = CA + CB;
  const int warp = threadIdx.x >> 5;
  const int tiles_x = Wd / CW;
  const int tx0 = (blockIdx.x % tiles_x) * CW, ty0 = (blockIdx.x / tiles_x) * Sh::R;
  const int oc_tile = blockIdx.y, oc0 = oc_tile * NT;
  const int n = blockIdx.z / ksplit, ks = blockIdx.z - n * ksplit;
  const int chunks_all = C / kTcKc, per = (chunks_all + ksplit - 1) / ksplit;
  const int chunk_begin = ks * per, chunk_end = min(chunks_all, chunk_begin + per);
  const int n_chunks = max(chunk_end - chunk_begin, 0);
  const int n_stages = n_chunks * kTaps;
  const size_t plane = (size_t)H * Wd;

  if (threadIdx.x == 0) {
    for (int s_ = 0; s_ < kTcStages; ++s_) {
      mbar_init(&full[s_], 128 + 1);
      mbar_init(&empty[s_], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, NT);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ---------------------------------------------------------------- staging threads: pixel m of the tile
    const int m = threadIdx.x;
    const int pr = m / CW, pxl = m - pr * CW;
    // this thread's share of one raw channel plane: (smem offset, global offset, in range), the same for every channel
    int so[Sh::kRawPer], go[Sh::kRawPer];
    bool ok[Sh::kRawPer];
#pragma unroll
    for (int k = 0; k < Sh::kRawPer; ++k) {
      const int e = m + 128 * k;
      const int rr = e / (CW + 2), q = e - rr * (CW + 2);
      const int y = ty0 + rr - 1, x = tx0 + q - 1;
      so[k] = e < Sh::kPlane ? e : -1;
      ok[k] = e < Sh::kPlane && y >= 0 && y < H && x >= 0 && x < Wd;
      go[k] = ok[k] ? y * Wd + x : 0;
    }
    auto fetch = [&](int chunk, int buf) {
      const int ic0 = chunk * kTcKc;
      const float* src = ic0 < CA ? inA + ((size_t)n * CA + ic0) * plane : inB + ((size_t)n * CB + (ic0 - CA)) * plane;
      float* dst = raw + buf * Sh::kRawFloats;
#pragma unroll 4
      for (int c = 0; c < kTcKc; ++c) {
#pragma unroll
        for (int k = 0; k < Sh::kRawPer; ++k)
          if (so[k] >= 0) cp_async4_zfill(dst + c * Sh::kPlane + so[k], src + (size_t)c * plane + go[k], ok[k]);
      }
      cp_async_commit();
    };
    if (n_chunks > 0) fetch(chunk_begin, 0);
    int stage = 0;
    for (int ci = 0; ci < n_chunks; ++ci) {
      if (ci + 1 < n_chunks) {
        fetch(chunk_begin + ci + 1, (ci + 1) & 1);
        cp_async_wait_group<1>();
      } else {
        cp_async_wait_group<0>();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // chunk ci of the raw tile is complete for all staging threads
      const float* rt = raw + (ci & 1) * Sh::kRawFloats;
#pragma unroll 1
      for (int tap = 0; tap < kTaps; ++tap, ++stage) {
        const int slot = stage % kTcStages;
        mbar_wait_warp(&empty[slot], ((stage / kTcStages) & 1) ^ 1);
        const int ky = kTransposed ? 1 : tap / 3, kx = kTransposed ? 1 : tap - ky * 3;
        const float* srcp = rt + (pr + ky) * (CW + 2) + pxl + kx;
        uint8_t* arow = ring + slot * Sh::kStageBytes + (m >> 3) * 1024 + (m & 7) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v;
          v.x = srcp[(4 * j + 0) * Sh::kPlane];
          v.y = srcp[(4 * j + 1) * Sh::kPlane];
          v.z = srcp[(4 * j + 2) * Sh::kPlane];
          v.w = srcp[(4 * j + 3) * Sh::kPlane];
          *reinterpret_cast<float4*>(arow + ((j ^ (m & 7)) << 4)) = v;
        }
        fence_proxy_async();
        mbar_arrive(&full[slot]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // everyone is done with this raw buffer before it is refilled
    }
    // ---------------------------------------------------------------- epilogue: TMEM lane m = pixel m
    const int y = ty0 + pr, x = tx0 + pxl;
    if (n_stages > 0) {
      mbar_wait_warp(acc_bar, 0);
      tcgen05_fence_after();
    }
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    float* dst = ksplit > 1 ? partial + (size_t)ks * (gridDim.z / ksplit) * OC * plane : out;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      if (n_stages > 0) {
        tmem_ld32(tl + c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (y < H) {
        if constexpr (kTransposed) {
          const int t4 = (oc0 + c0) / oc_real, ocb = oc0 + c0 - t4 * oc_real;   // a 32-column group has one tap
          float* o = out + (((size_t)n * oc_real + ocb) * (2 * H) + 2 * y + (t4 >> 1)) * (2 * Wd) + 2 * x + (t4 & 1);
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)j * (4 * plane)] = v[j] + __ldg(shift + ocb + j);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int oc = oc0 + c0 + j;
            float r = v[j];
            if (ksplit == 1) r = fmaxf(fmaf(r, __ldg(scale + oc), __ldg(shift + oc)), 0.f);
            dst[(((size_t)n * OC + oc) * H + y) * Wd + x] = r;
          }
        }
      }
    }
    tcgen05_fence_before();
  } else if (warp == 4) {
    // ---------------------------------------------------------------- MMA issue
    const uint32_t idesc = make_idesc_tf32(kTcM, NT);
    for (int stage = 0; stage < n_stages; ++stage) {
      const int slot = stage % kTcStages;
      mbar_wait_warp(&full[slot], (stage / kTcStages) & 1);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint64_t da = make_desc_sw128(ring + slot * Sh::kStageBytes);
        const uint64_t db = make_desc_sw128(ring + slot * Sh::kStageBytes + kTcABytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma_tf32_ss(tmem, da + kk * 2, db + kk * 2, idesc, stage > 0 || kk > 0);
        mma_commit(&empty[slot]);
        if (stage == n_stages - 1) mma_commit(acc_bar);
      }
      __syncwarp();
    }
    tcgen05_fence_before();
  } else {
    // ---------------------------------------------------------------- weight copies: warp 5 + s serves ring slot s
    if ((threadIdx.x & 31) == 0) {
      for (int stage = warp - 5; stage < n_stages; stage += kTcStages) {
        const int slot = stage % kTcStages;
        mbar_wait_relaxed(&empty[slot], ((stage / kTcStages) & 1) ^ 1);
        const int chunk = chunk_begin + stage / kTaps, tap = stage % kTaps;
        const float* src = packed + (((size_t)oc_tile * chunks_all + chunk) * kTaps + tap) * ((size_t)NT * kTcKc);
        mbar_expect_tx(&full[slot], Sh::kBBytes);
        bulk_g2s(ring + slot * Sh::kStageBytes + kTcABytes, src, Sh::kBBytes, &full[slot]);
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, NT);
  }
}

// Sum of the ksplit partial slices (in slice order: deterministic) + folded BatchNorm + ReLU.
__global__ void k_conv_finish(const float* __restrict__ partial, int ksplit, size_t slice /*N*OC*plane*/, size_t plane, int OC,
                              const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                              float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= slice) return;
  float acc = __ldg(partial + i);
  for (int k = 1; k < ksplit; ++k) acc += __ldg(partial + (size_t)k * slice + i);
  const int oc = (int)((i / plane) % OC);
  const float r = fmaf(acc, scale ? __ldg(scale + oc) : 1.f, shift ? __ldg(shift + oc) : 0.f);
  out[i] = relu ? fmaxf(r, 0.f) : r;
}

__global__ void k_maxpool2(const float* __restrict__ in, int planes, int H, int Wd, float* __restrict__ out) {
  const int Ho = H / 2, Wo = Wd / 2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)planes * Ho * Wo) return;
  const int x = i % Wo, y = (i / Wo) % Ho;
  const size_t p = i / ((size_t)Wo * Ho);
  const float* s = in + (p * H + 2 * y) * Wd + 2 * x;
  out[i] = fmaxf(fmaxf(__ldg(s), __ldg(s + 1)), fmaxf(__ldg(s + Wd), __ldg(s + Wd + 1)));
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True): src = dst * (in - 1) / (out - 1)
__global__ void k_upsample2_bilinear(const float* __restrict__ in, int planes, int H, int Wd, float* __restrict__ out) {
  const int Ho = 2 * H, Wo = 2 * Wd;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)planes * Ho * Wo) return;
  const int x = i % Wo, y = (i / Wo) % Ho;
  const size_t p = i / ((size_t)Wo * Ho);
  const float ry = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f, rx = Wo > 1 ? (float)(Wd - 1) / (float)(Wo - 1) : 0.f;
  const float sy = ry * y, sx = rx * x;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < Wd - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float* s = in + p * H * Wd;
  const float v00 = __ldg(s + (size_t)y0 * Wd + x0), v01 = __ldg(s + (size_t)y0 * Wd + x1);
  const float v10 = __ldg(s + (size_t)y1 * Wd + x0), v11 = __ldg(s + (size_t)y1 * Wd + x1);
  out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// nn.ConvTranspose2d(C, OC, kernel_size=2, stride=2): out[n, oc, 2y+dy, 2x+dx] = b[oc] + sum_ic in[n, ic, y, x] W[ic, oc, dy, dx]
__global__ void k_convtranspose2(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                 int C, int OC, int H, int Wd, float* __restrict__ out) {
  const int Ho = 2 * H, Wo = 2 * Wd;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= (size_t)OC * Ho * Wo) return;
  const int xo = i % Wo, yo = (i / Wo) % Ho, oc = i / ((size_t)Wo * Ho);
  const int x = xo >> 1, y = yo >> 1, k = (yo & 1) * 2 + (xo & 1);
  float acc = bias ? __ldg(bias + oc) : 0.f;
  const float* s = in + (size_t)n * C * H * Wd + (size_t)y * Wd + x;
  for (int ic = 0; ic < C; ++ic) acc = fmaf(__ldg(s + (size_t)ic * H * Wd), __ldg(W + ((size_t)ic * OC + oc) * 4 + k), acc);
  out[(size_t)n * OC * Ho * Wo + i] = acc;
}

// OutConv: 1x1 convolution with bias
__global__ void k_conv1x1(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias, int C,
                          int OC, size_t plane, float* __restrict__ out, int sigmoid) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (i >= plane) return;
  for (int oc = 0; oc < OC; ++oc) {
    float acc = bias ? __ldg(bias + oc) : 0.f;
    for (int ic = 0; ic < C; ++ic) acc = fmaf(__ldg(in + ((size_t)n * C + ic) * plane + i), __ldg(W + (size_t)oc * C + ic), acc);
    out[((size_t)n * OC + oc) * plane + i] = sigmoid ? 1.0f / (1.0f + expf(-acc)) : acc;
  }
}

// Split of the input-channel loop: the layers below full resolution have 8-128 tiles for 148 SMs (measured before
// the split: 64 -> 64 at 32 x 1024, 256 blocks, 123 us; 512 -> 512 at 2 x 64, 16 blocks, 561 us for HALF the flops).
// Enough slices for ~4 blocks per SM, at most `max_split` (the partial buffer holds 128 floats per full-resolution
// pixel: a level-k tensor is 64 / 2^k of them, so up to 2^(k+1) slices fit).
static int pick_ksplit(int blocks, int C, size_t slice_floats, size_t partial_floats) {
  static const int want = 4 * nlb_sm_count();
  int ks = (want + blocks - 1) / blocks;
  const int chunks = (C + kIcChunk - 1) / kIcChunk;
  if (ks > chunks) ks = chunks;
  if ((size_t)ks * slice_floats > partial_floats) ks = (int)(partial_floats / slice_floats);
  if (ks < 1) ks = 1;
  // no empty slices: ceil(chunks / ks) chunks per slice must leave the last slice non-empty
  while (ks > 1 && (ks - 1) * ((chunks + ks - 1) / ks) >= chunks) --ks;
  return ks;
}

// The tensor-core path needs whole 32-channel chunks on both inputs, whole output-channel tiles and an image the
// 128-pixel tile divides (1 x 128 or 2 x 64); everything else (the first layer: 6 input channels) stays on k_conv3x3.
static bool tc_eligible(int CA, int CB, int OC, int H, int Wd) {
  return CA % kTcKc == 0 && CB % kTcKc == 0 && OC % 64 == 0 && (Wd % 128 == 0 || Wd == 64) && H >= 1;
}
static int tc_oc_tile(int OC) { return OC % 128 == 0 ? 128 : 64; }

template <int NT, int CW>
static int launch_tf32(const float* inA, int CA, const float* inB, int CB, const nlb_unet_conv_t& L, int OC, int N, int H,
                       int Wd, float* out, float* partial, size_t partial_floats, cudaStream_t st, int oc_real = 0) {
  using Sh = TcShape<NT, CW>;
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    if (cudaFuncSetAttribute(k_conv3x3_tf32<NT, CW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sh::kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(k_conv3x3_tf32<NT, CW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sh::kSmem) != cudaSuccess) {
      nlb_set_error("unet: cannot reserve %zu bytes of shared memory", Sh::kSmem);
      return NLB_ECUDA;
    }
    attr[dev] = true;
  }
  const int tiles = (Wd / CW) * ((H + Sh::R - 1) / Sh::R), oc_tiles = OC / NT;
  const size_t slice = (size_t)N * OC * H * Wd;
  // one CTA per SM (200 KB of shared memory): slices of the input channels until every SM has a CTA (a layer with
  // that many tiles already is not split: no partial slices, no finishing kernel)
  const int blocks = tiles * oc_tiles * N, chunks = (CA + CB) / kTcKc;
  int ksplit = (nlb_sm_count() * 5 / 4 + blocks - 1) / blocks;
  if (ksplit > chunks) ksplit = chunks;
  if ((size_t)ksplit * slice > partial_floats) ksplit = (int)(partial_floats / slice);
  if (ksplit < 1) ksplit = 1;
  while (ksplit > 1 && (ksplit - 1) * ((chunks + ksplit - 1) / ksplit) >= chunks) --ksplit;
  if (oc_real > 0) {   // transposed convolution: OC = 4 * oc_real virtual channels, never split
    k_conv3x3_tf32<NT, CW, true><<<dim3(tiles, oc_tiles, N), kTcThreads, Sh::kSmem, st>>>(
        inA, CA, inB, CB, L.packed, nullptr, L.shift, OC, H, Wd, out, 1, nullptr, oc_real);
    return nlb_check_launch("unet convtranspose tf32");
  }
  k_conv3x3_tf32<NT, CW, false><<<dim3(tiles, oc_tiles, N * ksplit), kTcThreads, Sh::kSmem, st>>>(
      inA, CA, inB, CB, L.packed, L.scale, L.shift, OC, H, Wd, out, ksplit, partial, 0);
  if (int e = nlb_check_launch("unet conv3x3 tf32")) return e;
  if (ksplit > 1) {
    k_conv_finish<<<(unsigned)((slice + 255) / 256), 256, 0, st>>>(partial, ksplit, slice, (size_t)H * Wd, OC, L.scale, L.shift, 1, out);
    return nlb_check_launch("unet conv finish");
  }
  return NLB_OK;
}

static int conv3x3_tf32(const float* inA, int CA, const float* inB, int CB, const nlb_unet_conv_t& L, int OC, int N, int H,
                        int Wd, float* out, float* partial, size_t partial_floats, cudaStream_t st, int oc_real = 0) {
  const bool wide = Wd % 128 == 0;
  const int nt = tc_oc_tile(oc_real > 0 ? oc_real : OC);
  if (nt == 128)
    return wide ? launch_tf32<128, 128>(inA, CA, inB, CB, L, OC, N, H, Wd, out, partial, partial_floats, st, oc_real)
                : launch_tf32<128, 64>(inA, CA, inB, CB, L, OC, N, H, Wd, out, partial, partial_floats, st, oc_real);
  return wide ? launch_tf32<64, 128>(inA, CA, inB, CB, L, OC, N, H, Wd, out, partial, partial_floats, st, oc_real)
              : launch_tf32<64, 64>(inA, CA, inB, CB, L, OC, N, H, Wd, out, partial, partial_floats, st, oc_real);
}

// ConvTranspose2d(C, OC, 2, 2) on the same kernel: in[N,C,H,W] -> out[N,OC,2H,2W]
static int convtranspose_tf32(const float* in, int C, const float* packed, const float* bias, int OC, int N, int H, int Wd,
                              float* out, cudaStream_t st) {
  nlb_unet_conv_t L{nullptr, nullptr, bias, packed};
  return conv3x3_tf32(in, C, nullptr, 0, L, 4 * OC, N, H, Wd, out, nullptr, 0, st, OC);
}

static int conv3x3(const float* inA, int CA, const float* inB, int CB, const nlb_unet_conv_t& L, int OC, int N, int H, int Wd,
                   float* out, float* partial, size_t partial_floats, cudaStream_t st) {
  if (!inA || !L.weight || !L.scale || !L.shift || !out || (CB > 0 && !inB)) { nlb_set_error("unet: null pointer in a 3x3 layer"); return NLB_EINVAL; }
  const int tiles = ((Wd + kCols - 1) / kCols) * ((H + kRows - 1) / kRows);
  const int oc_tiles = (OC + kOcTile - 1) / kOcTile;
  const size_t plane = (size_t)H * Wd, slice = (size_t)N * OC * plane;
  static const bool kNoSplit = getenv("NLB_UNET_NO_SPLIT") != nullptr;   // A/B timing
  if (L.packed && tc_eligible(CA, CB, OC, H, Wd) && partial)
    return conv3x3_tf32(inA, CA, inB, CB, L, OC, N, H, Wd, out, partial, partial_floats, st);
  const int ksplit = (partial && !kNoSplit) ? pick_ksplit(tiles * oc_tiles * N, CA + CB, slice, partial_floats) : 1;
  dim3 grid(tiles, oc_tiles, N * ksplit);
  k_conv3x3<<<grid, kConvThreads, 0, st>>>(inA, CA, inB, CB, L.weight, L.scale, L.shift, OC, H, Wd, 1, out, ksplit, partial);
  if (int e = nlb_check_launch("unet conv3x3")) return e;
  if (ksplit > 1) {
    k_conv_finish<<<(unsigned)((slice + 255) / 256), 256, 0, st>>>(partial, ksplit, slice, plane, OC, L.scale, L.shift, 1, out);
    return nlb_check_launch("unet conv finish");
  }
  return NLB_OK;
}

}  // namespace unet
}  // namespace nlb

using namespace nlb;

// activations of one forward pass, in floats: x1..x5 (encoder), one pooled / upsampled scratch, one mid tensor and
// one decoder output per level (sized for the finest level), for N images
extern "C" size_t nlb_unet_workspace_bytes(int N, int H, int W) {
  const size_t px = (size_t)N * H * W;
  // x1 64, x2 128/4, x3 256/16, x4 512/64, x5 1024/256 pixel-equivalents of floats + four scratch tensors (pooled /
  // upsampled input, mid tensor, two ping-pong decoder outputs) of at most 64 channels at full resolution (96 reserved)
  // + the partial sums of the input-channel split (128)
  return (64 + 32 + 16 + 8 + 4 + 4 * 96 + 128) * px * sizeof(float) + 1024;
}

extern "C" int nlb_unet_forward(const float* image, const nlb_unet_weights_t* w, int N, int Cin, int H, int W, float* logits,
                                float* regression, float* workspace, void* stream) {
  if (N == 0) return NLB_OK;
  if (N < 0 || Cin < 1 || !image || !w || !logits || !workspace) { nlb_set_error("unet_forward: bad argument"); return NLB_EINVAL; }
  if (H < 16 || W < 16 || H % 16 || W % 16) {
    nlb_set_error("unet_forward: H and W must be multiples of 16 (four 2x2 poolings without padding), got %d x %d", H, W);
    return NLB_EUNSUPPORTED;
  }
  if (w->n_classes < 1 || !w->outc_weight) { nlb_set_error("unet_forward: the output layer is required"); return NLB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const int bil = w->bilinear != 0;
  const int f = bil ? 2 : 1;
  const size_t px = (size_t)N * H * W;
  float* x1 = workspace;                 // [N, 64, H, W]
  float* x2 = x1 + 64 * px;              // [N, 128, H/2, W/2]
  float* x3 = x2 + 32 * px;              // [N, 256, H/4, W/4]
  float* x4 = x3 + 16 * px;              // [N, 512, H/8, W/8]
  float* x5 = x4 + 8 * px;               // [N, 1024/f, H/16, W/16]
  float* s0 = x5 + 4 * px;               // scratch: pooled / upsampled input
  float* s1 = s0 + 96 * px;              // scratch: mid tensor of a DoubleConv
  float* s2 = s1 + 96 * px;              // scratch: decoder level outputs (ping-pong)
  float* s3 = s2 + 96 * px;
  float* part = s3 + 96 * px;            // partial sums of the input-channel split
  const size_t part_floats = 128 * px;
  auto double_conv = [&](const float* inA, int CA, const float* inB, int CB, const nlb_unet_conv_t* L, int mid, int oc, int h,
                         int wd, float* out) -> int {
    if (int e = unet::conv3x3(inA, CA, inB, CB, L[0], mid, N, h, wd, s1, part, part_floats, st)) return e;
    return unet::conv3x3(s1, mid, nullptr, 0, L[1], oc, N, h, wd, out, part, part_floats, st);
  };
  auto pool = [&](const float* in, int C, int h, int wd) {
    const size_t total = (size_t)N * C * (h / 2) * (wd / 2);
    unet::k_maxpool2<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, N * C, h, wd, s0);
  };
  // encoder
  if (int e = double_conv(image, Cin, nullptr, 0, w->inc, 64, 64, H, W, x1)) return e;
  pool(x1, 64, H, W);
  if (int e = double_conv(s0, 64, nullptr, 0, w->down[0], 128, 128, H / 2, W / 2, x2)) return e;
  pool(x2, 128, H / 2, W / 2);
  if (int e = double_conv(s0, 128, nullptr, 0, w->down[1], 256, 256, H / 4, W / 4, x3)) return e;
  pool(x3, 256, H / 4, W / 4);
  if (int e = double_conv(s0, 256, nullptr, 0, w->down[2], 512, 512, H / 8, W / 8, x4)) return e;
  pool(x4, 512, H / 8, W / 8);
  if (int e = double_conv(s0, 512, nullptr, 0, w->down[3], 1024 / f, 1024 / f, H / 16, W / 16, x5)) return e;
  // decoder: Up(in, out): x1' = up(x_low); DoubleConv(cat[x_skip, x1'], out, mid)
  const float* low = x5;
  int c_low = 1024 / f;
  const float* skips[4] = {x4, x3, x2, x1};
  const int c_skip[4] = {512, 256, 128, 64};
  const int c_out[4] = {512 / f, 256 / f, 128 / f, 64};
  int h = H / 16, wd = W / 16;
  float* outs[2] = {s2, s3};
  for (int u = 0; u < 4; ++u) {
    int c_up;
    if (bil) {
      c_up = c_low;
      const size_t total = (size_t)N * c_low * (2 * h) * (2 * wd);
      unet::k_upsample2_bilinear<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(low, N * c_low, h, wd, s0);
    } else {
      c_up = c_low / 2;
      if (!w->up_weight[u]) { nlb_set_error("unet_forward: transposed-convolution weights of up%d are required", u + 1); return NLB_EINVAL; }
      if (w->up_packed[u] && w->up_bias[u] && unet::tc_eligible(c_low, 0, c_up, h, wd)) {
        if (int e = unet::convtranspose_tf32(low, c_low, w->up_packed[u], w->up_bias[u], c_up, N, h, wd, s0, st)) return e;
      } else {
        const size_t total = (size_t)c_up * (2 * h) * (2 * wd);
        unet::k_convtranspose2<<<dim3((unsigned)((total + 255) / 256), N), 256, 0, st>>>(low, w->up_weight[u], w->up_bias[u],
                                                                                        c_low, c_up, h, wd, s0);
      }
    }
    h *= 2; wd *= 2;
    const int cin = c_skip[u] + c_up;
    const int mid = bil ? cin / 2 : c_out[u];
    float* dst = outs[u & 1];   // never aliases its inputs: s0 (upsampled), s1 (mid), the skip tensor, `low` = the other one
    if (int e = double_conv(skips[u], c_skip[u], s0, c_up, w->up[u], mid, c_out[u], h, wd, dst)) return e;
    low = dst;
    c_low = c_out[u];
  }
  unet::k_conv1x1<<<dim3((unsigned)(((size_t)H * W + 255) / 256), N), 256, 0, st>>>(low, w->outc_weight, w->outc_bias, 64,
                                                                                    w->n_classes, (size_t)H * W, logits, 0);
  if (regression) {   // UNet(regression=True): reg = sigmoid(outr(x)) (unet_model.py:45-46)
    if (!w->outr_weight) { nlb_set_error("unet_forward: a regression output needs the outr layer"); return NLB_EINVAL; }
    unet::k_conv1x1<<<dim3((unsigned)(((size_t)H * W + 255) / 256), N), 256, 0, st>>>(low, w->outr_weight, w->outr_bias, 64, 1,
                                                                                      (size_t)H * W, regression, 1);
  }
  return nlb_check_launch("unet_forward");
}

// Operand-layout copy of a 3x3 layer's weights for the TF32 tensor-core path (nlb_unet_conv_t.packed): same number
// of floats as the weights.  Layers the path does not serve (C not a multiple of 32, OC not a multiple of 64) need
// none: returns NLB_EUNSUPPORTED and the caller leaves `packed` NULL.
extern "C" int nlb_unet_pack_conv(const float* weight, int OC, int C, float* packed, void* stream) {
  if (!weight || !packed) { nlb_set_error("unet_pack_conv: null pointer"); return NLB_EINVAL; }
  if (OC < 64 || OC % 64 || C < unet::kTcKc || C % unet::kTcKc) {
    nlb_set_error("unet_pack_conv: the TF32 path needs OC %% 64 == 0 and C %% 32 == 0 (got %d, %d)", OC, C);
    return NLB_EUNSUPPORTED;
  }
  const size_t total = (size_t)OC * C * 9;
  unet::k_pack_conv_weights<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weight, OC, C, unet::tc_oc_tile(OC), packed);
  return nlb_check_launch("unet_pack_conv");
}

// The same for a ConvTranspose2d(C, OC, kernel 2, stride 2) weight [C, OC, 2, 2] (nlb_unet_weights_t.up_packed).
extern "C" int nlb_unet_pack_convtranspose(const float* weight, int C, int OC, float* packed, void* stream) {
  if (!weight || !packed) { nlb_set_error("unet_pack_convtranspose: null pointer"); return NLB_EINVAL; }
  if (OC < 64 || OC % 64 || C < unet::kTcKc || C % unet::kTcKc) {
    nlb_set_error("unet_pack_convtranspose: the TF32 path needs OC %% 64 == 0 and C %% 32 == 0 (got %d, %d)", OC, C);
    return NLB_EUNSUPPORTED;
  }
  const size_t total = (size_t)C * OC * 4;
  unet::k_pack_convT_weights<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weight, C, OC, unet::tc_oc_tile(OC), packed);
  return nlb_check_launch("unet_pack_convtranspose");
}
