// Dev probe: tcgen05.mma with MN-MAJOR shared-memory operands, as the weight-gradient kernel uses them.
//   D[Mdim, N] = A^T B,  A = [K, Mdim] row-major, B = [K, N] row-major  (K = the sample index m)
// Operands sit in shared memory exactly as the other kernels keep [rows][64] bf16 blocks (128-byte rows,
// SWIZZLE_128B, 8-row / 1024-byte atoms) -- here the ROWS are K and the 64 columns are 64 M/N elements, which is
// the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = byte distance of the
// next 64-element block along M/N, SBO = byte distance of the next 8 K-rows (1024).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn_probe umma_mn_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include "../nerf_lidar_b200/csrc/umma.cuh"

using namespace nlb::umma;

__device__ __forceinline__ uint64_t desc_mn(const void* p, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t a = smem_u32(p);
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// KROWS = K (rows per block, multiple of 16, <= 128); Mdim in {64,128}; N multiple of 16 (Mdim=128) / 8
template <int Mdim, int N, int KROWS>
__global__ void __launch_bounds__(160) probe(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                             float* __restrict__ out, int swap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int BLK = KROWS * 128;            // bytes of one [KROWS][64] block
  constexpr int MA = (Mdim + 63) / 64, NB = (N + 63) / 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + MA * BLK;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  // stage: row k of A -> row k of every A block (64 columns each), 16-byte chunks swizzled by (k & 7)
  for (int e = threadIdx.x; e < KROWS * (Mdim / 8); e += blockDim.x) {
    const int k = e / (Mdim / 8), ch = e % (Mdim / 8);
    const uint4 v = *reinterpret_cast<const uint4*>(a + (size_t)k * Mdim + ch * 8);
    uint8_t* blk = sA + (ch / 8) * BLK;
    *reinterpret_cast<uint4*>(blk + (k >> 3) * 1024 + (k & 7) * 128 + (((ch & 7) ^ (k & 7)) * 16)) = v;
  }
  for (int e = threadIdx.x; e < KROWS * (N / 8); e += blockDim.x) {
    const int k = e / (N / 8), ch = e % (N / 8);
    const uint4 v = *reinterpret_cast<const uint4*>(b + (size_t)k * N + ch * 8);
    uint8_t* blk = sB + (ch / 8) * BLK;
    *reinterpret_cast<uint4*>(blk + (k >> 3) * 1024 + (k & 7) * 128 + (((ch & 7) ^ (k & 7)) * 16)) = v;
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 4 && elect_one_sync()) {
    const uint32_t idesc = make_idesc_bf16(Mdim, N) | (1u << 15) | (1u << 16);
    const uint32_t lbo = swap ? 1024 : BLK, sbo = swap ? BLK : 1024;
    for (int kk = 0; kk < KROWS / 16; ++kk)
      mma_bf16_ss(tmem, desc_mn(sA + kk * 2048, lbo, sbo), desc_mn(sB + kk * 2048, lbo, sbo), idesc, kk != 0);
    mma_commit(&bar_mma);
  }
  __syncwarp();
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;   // TMEM lane; for Mdim = 64 rows 0..63 live in lanes 0..63? (checked by the host)
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = v[j];
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

template <int Mdim, int N, int KROWS>
static int run() {
  std::vector<float> A((size_t)KROWS * Mdim), B((size_t)KROWS * N);
  srand(Mdim + N + KROWS);
  for (auto& x : A) x = (rand() % 17 - 8) / 8.0f;
  for (auto& x : B) x = (rand() % 13 - 6) / 4.0f;
  std::vector<__nv_bfloat16> ha(A.size()), hb(B.size());
  for (size_t i = 0; i < A.size(); ++i) ha[i] = __float2bfloat16(A[i]);
  for (size_t i = 0; i < B.size(); ++i) hb[i] = __float2bfloat16(B[i]);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ((Mdim + 63) / 64 + (N + 63) / 64) * (size_t)KROWS * 128 + 1024;
  cudaFuncSetAttribute(probe<Mdim, N, KROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int ok_any = 0;
  for (int swap = 0; swap < 1; ++swap) {   // (the swapped reading faults: LBO = slab stride, SBO = 1024 it is)
    cudaMemset(dout, 0xff, 128 * N * 4);
    probe<Mdim, N, KROWS><<<1, 160, smem>>>(da, db, dout, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d K=%d swap=%d CUDA error %s\n", Mdim, N, KROWS, swap, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(128 * N);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < Mdim; ++r)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < KROWS; ++k) ref += (double)A[(size_t)k * Mdim + r] * B[(size_t)k * N + n];
        maxerr = fmax(maxerr, fabs(ref - out[(size_t)r * N + n]));
      }
    printf("M=%d N=%d K=%d %s max abs err %.4g %s\n", Mdim, N, KROWS, swap ? "LBO=1024,SBO=block" : "LBO=block,SBO=1024",
           maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    if (maxerr < 1e-3 && swap == 0) ok_any = 1;
  }
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return ok_any ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<128, 256, 128>();
  bad += run<128, 64, 64>();
  bad += run<128, 128, 32>();
  bad += run<128, 16, 128>();
  bad += run<128, 32, 128>();
  // (M = 64 places the accumulator rows on other TMEM lanes than this probe reads back: the kernels use M = 128 only)
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
