// Dev probe: L1 / L2 cost of the access patterns of the hash-grid kernels on B200 -- scattered 4 / 8 / 16-byte
// gathers and reductions from an L2-resident 8 MB table (one hashed level), as a function of how the 32 lanes of
// one instruction share 128-byte lines and 32-byte sectors.  Answers what the pair-lane kernels rely on (two
// lanes of an instruction in the same line = one L1 wavefront; in the same sector = one L2 request) and what a
// pair-lane version of the dense-level scatter could gain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu && tools/gather_probe
// Patterns (per warp instruction): 0 = 32 lanes, 32 random lines; 1 = lane pairs share an aligned 8-byte pair;
// 2 = lane pairs share a 128-byte line, different sectors; 3 = lane quads share a line; 4 = fully coalesced.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr uint32_t kRows = 1u << 21;  // floats: 8 MB, L2-resident
constexpr int kIters = 64;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// row index (in units of `width` floats) of lane `lane` for instruction `it` of warp `w`
__device__ __forceinline__ uint32_t row_of(int pattern, uint32_t w, int it, int lane, uint32_t rows) {
  const uint32_t base = mix(w * 977u + it * 131071u);
  switch (pattern) {
    case 0: return mix(base + lane * 2654435761u) % rows;
    case 1: return ((mix(base + (lane >> 1) * 2654435761u) % rows) & ~1u) | (lane & 1);
    case 2: {  // same 128-byte line, sectors 16 bytes.. apart by half a line
      const uint32_t line_rows = 32;  // rows per 128 bytes at 4 bytes per row (scaled by the caller for wider rows)
      const uint32_t r = (mix(base + (lane >> 1) * 2654435761u) % rows) & ~(line_rows - 1);
      return r + (lane & 1) * (line_rows / 2) + ((lane >> 1) & 7);
    }
    case 3: {
      const uint32_t r = (mix(base + (lane >> 2) * 2654435761u) % rows) & ~31u;
      return r + (lane & 3) * 8 + ((lane >> 2) & 7);
    }
    default: return (base % (rows - 32)) + lane;
  }
}

template <int W>  // floats per access: 1, 2, 4
__global__ void k_gather(const float* __restrict__ table, float* out, int pattern) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rows = kRows / W;
  float acc = 0.f;
#pragma unroll 4
  for (int it = 0; it < kIters; ++it) {
    uint32_t r = row_of(pattern, w, it, lane, rows);
    if (pattern == 2 || pattern == 3) r = (r / W) % rows;  // keep the line / sector sharing at wider rows
    const float* p = table + (size_t)r * W;
    if (W == 1) acc += __ldg(p);
    if (W == 2) { float2 v = __ldg(reinterpret_cast<const float2*>(p)); acc += v.x + v.y; }
    if (W == 4) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); acc += v.x + v.y + v.z + v.w; }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int W>
__global__ void k_red(float* __restrict__ table, int pattern) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rows = kRows / W;
#pragma unroll 4
  for (int it = 0; it < kIters; ++it) {
    uint32_t r = row_of(pattern, w, it, lane, rows);
    if (pattern == 2 || pattern == 3) r = (r / W) % rows;
    float* p = table + (size_t)r * W;
    if (W == 1) atomicAdd(p, 1.0f);
    if (W == 2) atomicAdd(reinterpret_cast<float2*>(p), make_float2(1.f, 1.f));
    if (W == 4) atomicAdd(reinterpret_cast<float4*>(p), make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

template <class F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  float *table, *out;
  cudaMalloc(&table, (size_t)kRows * 4);
  cudaMemset(table, 0, (size_t)kRows * 4);
  cudaMalloc(&out, 4);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = sms * 16, threads = 256;
  const double lane_accesses = (double)blocks * threads * kIters;
  const char* names[] = {"32 random lines", "pairs share 8 bytes", "pairs share a line", "quads share a line", "coalesced"};
  printf("%d SMs, %.0f MHz; %d blocks x %d threads x %d accesses per lane, 8 MB table\n", sms, khz / 1e3, blocks, threads, kIters);
  for (int width : {1, 2, 4}) {
    for (int pattern = 0; pattern < 5; ++pattern) {
      float g = 0.f, r = 0.f;
      if (width == 1) { g = time_ms([&] { k_gather<1><<<blocks, threads>>>(table, out, pattern); }); r = time_ms([&] { k_red<1><<<blocks, threads>>>(table, pattern); }); }
      if (width == 2) { g = time_ms([&] { k_gather<2><<<blocks, threads>>>(table, out, pattern); }); r = time_ms([&] { k_red<2><<<blocks, threads>>>(table, pattern); }); }
      if (width == 4) { g = time_ms([&] { k_gather<4><<<blocks, threads>>>(table, out, pattern); }); r = time_ms([&] { k_red<4><<<blocks, threads>>>(table, pattern); }); }
      const double cyc = (double)khz * 1e3 * 1e-3;  // cycles per ms
      printf("%2d B  %-20s gather %7.3f ms = %5.2f lane-accesses/clk/SM   red %7.3f ms = %5.2f lane-accesses/clk/SM\n", width * 4,
             names[pattern], g, lane_accesses / (g * cyc * sms), r, lane_accesses / (r * cyc * sms));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
