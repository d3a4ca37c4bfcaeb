// Dev probe: sustained tcgen05.mma issue/execute rate of one CTA per SM for the shapes
// nerf_mlp.cu uses, with and without concurrent bulk copies into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nerf_lidar_b200/csrc/umma.cuh"
using namespace nlb::umma;

struct Res { long long cycles; long long ns; };

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// mode bit0: N=256 (else 128); bit1: concurrent bulk copies (4 producer lanes); bit2: commit after every 4 MMAs
__global__ void __launch_bounds__(288, 1) probe(const uint8_t* __restrict__ gsrc, int rounds, int mode, Res* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                 // 4 x 16 KB
  uint8_t* sB = base + 4 * 16384;     // 2 x 16 KB ([256][64])
  uint8_t* ring = base + 6 * 16384;   // 4 x 16 KB copy targets
  __shared__ uint64_t bar_done, bar_dummy, bar_cp[4];
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done, 1); mbar_init(&bar_dummy, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_cp[i], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 4) tmem_alloc(&tmem_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_s;
  const bool n256 = mode & 1, copies = mode & 2, commit4 = mode & 4;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n256 ? 256 : 128);
    long long t0 = clock64(), g0 = gtimer();
    for (int r = 0; r < rounds; ++r) {
      const uint64_t adesc = make_desc_sw128(sA + (r & 3) * 16384);
      const uint64_t bdesc = make_desc_sw128(sB);
      for (int kk = 0; kk < 4; ++kk) mma_bf16_ss(tmem + (r & 1) * 256, adesc + kk * 2, bdesc + kk * 2, idesc, kk != 0);
      if (commit4) mma_commit(&bar_dummy);
    }
    mma_commit(&bar_done);
    long long t1 = clock64();
    mbar_wait(&bar_done, 0);
    long long t2 = clock64(), g2 = gtimer();
    if (blockIdx.x == 0) { res[0].cycles = t1 - t0; res[0].ns = 0; res[1].cycles = t2 - t0; res[1].ns = g2 - g0; }
  } else if (warp >= 5 && lane == 0 && copies) {
    const int st = warp - 5;
    // each producer streams `rounds/4` chunks of 16 KB into its ring slot, waiting for each to land
    for (int i = 0; i < rounds / 4; ++i) {
      mbar_expect_tx(&bar_cp[st], 16384);
      bulk_g2s(ring + st * 16384, gsrc + (size_t)((i * 4 + st) % 32) * 16384, 16384, &bar_cp[st]);
      mbar_wait(&bar_cp[st], i & 1);
    }
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
  uint8_t* gsrc; cudaMalloc(&gsrc, 32 * 16384); cudaMemset(gsrc, 0, 32 * 16384);
  Res* res; cudaMallocManaged(&res, 2 * sizeof(Res));
  const size_t smem = 10 * 16384 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rounds = 400;
  for (int grid : {1, 148}) for (int mode = 0; mode < 8; ++mode) {
    probe<<<grid, 288, smem>>>(gsrc, rounds, mode, res);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    const int n = (mode & 1) ? 256 : 128;
    const double mmas = rounds * 4.0;
    printf("grid %3d N=%d copies=%d commit4=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA, %.1f ns/MMA -> clock %.0f MHz, %.0f MAC/cyc/SM\n",
           grid, n, !!(mode & 2), !!(mode & 4), res[0].cycles / mmas, res[1].cycles / mmas, res[1].ns / mmas,
           res[1].cycles * 1e3 / res[1].ns, 128.0 * n * 16 * mmas / res[1].cycles);
  }
  return 0;
}
