// Dev probe: latency / overlap of 1-D bulk async copies (UBLKCP) global -> shared on B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nerf_lidar_b200/csrc/umma.cuh"
using namespace nlb::umma;

__global__ void probe(const uint8_t* __restrict__ src, long long* out, int bytes, int inflight, int rounds, size_t span, int nprod, int lanes_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars_all[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&bars_all[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const bool producer = lanes_mode ? (threadIdx.x < nprod) : ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nprod);
  const int pid = lanes_mode ? threadIdx.x : (threadIdx.x >> 5);
  if (producer) {
    uint8_t* smem_p = smem + (size_t)pid * inflight * bytes;
    uint64_t* bars = bars_all + pid * 4;
    // warm L2
    size_t off = ((size_t)(blockIdx.x * 8 + pid) * 7919 * 4096) % span;
    long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t phase[16] = {0};
    for (int r = 0; r < rounds; ++r) {
      if (issued - done == inflight) {
        int s = done % inflight;
        mbar_wait(&bars[s], phase[s]); phase[s] ^= 1; ++done;
      }
      int s = issued % inflight;
      mbar_expect_tx(&bars[s], bytes);
      bulk_g2s(smem_p + s * bytes, src + (off + (size_t)issued * bytes) % span, bytes, &bars[s]);
      ++issued;
    }
    while (done < issued) { int s = done % inflight; mbar_wait(&bars[s], phase[s]); phase[s] ^= 1; ++done; }
    long long t1 = clock64();
    if (blockIdx.x == 0 && pid == 0) out[0] = t1 - t0;
  }
}

int main() {
  const size_t span = 600 * 1024;  // like the packed MLP weights: L2 resident
  uint8_t* src; cudaMalloc(&src, span + (1 << 20)); cudaMemset(src, 1, span + (1 << 20));
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int cfgs[][4] = {{16384, 2, 1, 0}, {16384, 2, 2, 0}, {16384, 2, 4, 0}, {16384, 2, 4, 1}, {8192, 2, 8, 0}, {8192, 2, 8, 1}, {16384, 1, 8, 0}, {65536, 2, 1, 0}, {49152, 4, 1, 0}};
  for (int grid : {1, 148}) {
    for (auto& c : cfgs) {
      int bytes = c[0], inflight = c[1], nprod = c[2], lanes = c[3], rounds = 64;
      for (int rep = 0; rep < 2; ++rep) probe<<<grid, 256, (size_t)nprod * inflight * bytes>>>(src, out, bytes, inflight, rounds, span, nprod, lanes);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d  chunk %6d B  in-flight %2d  producers %d (%s) : %7.1f cycles/chunk/producer  %6.1f B/cycle/SM  %s\n", grid, bytes, inflight, nprod,
             lanes ? "lanes" : "warps", (double)cyc / rounds, (double)bytes * rounds * nprod / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
