// Dev probe: one 128 x N x K tile on tcgen05 to validate the shared-memory operand
// layout (K-major, SWIZZLE_128B), the descriptors, TMEM alloc / ld and the bulk copy
// used by nerf_mlp.cu.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include "../nerf_lidar_b200/csrc/umma.cuh"

using namespace nlb::umma;

// D[128, N] = A[128, K] * B[N, K]^T ; A, B packed as 16 KB blocks [rows<=128][64] bf16 SW128.
template <int N, int KBLOCKS>
__global__ void __launch_bounds__(160) probe(const __nv_bfloat16* __restrict__ a_blocks,
                                             const __nv_bfloat16* __restrict__ b_blocks, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // KBLOCKS * 16 KB
  uint8_t* sB = smem + KBLOCKS * 16384;       // KBLOCKS * 16 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, 2 * KBLOCKS * 16384);
    for (int k = 0; k < KBLOCKS; ++k) {
      bulk_g2s(sA + k * 16384, a_blocks + (size_t)k * 8192, 16384, &bar_load);
      bulk_g2s(sB + k * 16384, b_blocks + (size_t)k * 8192, 16384, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int k = 0; k < KBLOCKS; ++k)
      for (int kk = 0; kk < 4; ++kk)
        mma_bf16_ss(tmem, make_desc_sw128(sA + k * 16384) + kk * 2, make_desc_sw128(sB + k * 16384) + kk * 2, idesc,
                    (k | kk) != 0);
    mma_commit(&bar_mma);
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) out[(size_t)row * N + c0 + j] = v[j];
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// host-side packing of a [rows, K] row-major matrix into SW128 K-major 16 KB blocks
static void pack(const std::vector<float>& m, int rows, int K, std::vector<__nv_bfloat16>& out, int kblocks) {
  out.assign((size_t)kblocks * 8192, __float2bfloat16(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      int kb = k / 64, kc = k % 64;
      size_t byte = sw128_offset(r, kc);
      out[(size_t)kb * 8192 + byte / 2] = __float2bfloat16(m[(size_t)r * K + k]);
    }
}

template <int N, int KBLOCKS>
static int run() {
  const int K = KBLOCKS * 64;
  std::vector<float> A(128 * K), B((size_t)N * K);
  srand(1);
  for (auto& x : A) x = (rand() % 17 - 8) / 8.0f;
  for (auto& x : B) x = (rand() % 13 - 6) / 4.0f;
  std::vector<__nv_bfloat16> pa, pb;
  pack(A, 128, K, pa, KBLOCKS);
  pack(B, N, K, pb, KBLOCKS);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, pa.size() * 2);
  cudaMalloc(&db, pb.size() * 2);
  cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, pa.data(), pa.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice);
  size_t smem = 2 * KBLOCKS * 16384 + 1024;
  cudaFuncSetAttribute(probe<N, KBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<N, KBLOCKS><<<1, 160, smem>>>(da, db, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
  std::vector<float> out(128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)r * K + k] * B[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(ref - out[(size_t)r * N + n]));
    }
  printf("N=%d K=%d max abs err %.4g %s\n", N, K, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return maxerr < 1e-3 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<128, 1>();
  bad += run<128, 2>();
  bad += run<64, 1>();
  bad += run<32, 2>();
  bad += run<16, 4>();
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
