// Write-pattern microbenchmark behind the canvas kernel's design (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pattern_bench tools/pattern_bench.cu
// Writes a (128, 64, 160000) float32 canvas (5.24 GB) with: A a linear float4 fill; B the canvas
// kernel's pattern (CTA = 128 cells x 64 channel rows of 512 B); C = B behind a dependent 512 B
// load + __syncthreads (the map read); D = B with plain st.global; E = persistent B (6 CTAs per SM
// looping over tiles, next tile's map word prefetched).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define NCELL 160000
#define C 64
#define B 128

__device__ __forceinline__ void st_na(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void k_linear(float4* out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool LOAD, bool NA>
__global__ void __launch_bounds__(256) k_tile(float* out, const int* map) {
  __shared__ int idx[128];
  const int tiles = NCELL / 128;
  const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
  float z = 0.f;
  if (LOAD) {
    if (threadIdx.x < 128) idx[threadIdx.x] = map[(size_t)b * NCELL + t * 128 + threadIdx.x];
    __syncthreads();
    z = idx[threadIdx.x & 127] > 0 ? 1.f : 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dst = out + (size_t)b * C * NCELL + (size_t)t * 128;
  for (int c = warp; c < C; c += 8) {
    float4* p = reinterpret_cast<float4*>(dst + (size_t)c * NCELL) + lane;
    if (NA) st_na(p, make_float4(z, z, z, z)); else *p = make_float4(z, z, z, z);
  }
}

__global__ void __launch_bounds__(256) k_persistent(float* out, const int* map) {
  const int tiles = NCELL / 128, total = tiles * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  int nxt = 0;
  if (tile < total && threadIdx.x < 128) nxt = map[(size_t)(tile / tiles) * NCELL + (tile % tiles) * 128 + threadIdx.x];
  for (; tile < total; tile += gridDim.x) {
    const int cur = nxt;
    const int tn = tile + gridDim.x;
    if (tn < total && threadIdx.x < 128) nxt = map[(size_t)(tn / tiles) * NCELL + (tn % tiles) * 128 + threadIdx.x];
    const int any = __syncthreads_or(cur > 0);
    const float z = any ? 1.f : 0.f;
    const int b = tile / tiles, t = tile - b * tiles;
    float* dst = out + (size_t)b * C * NCELL + (size_t)t * 128;
    for (int c = warp; c < C; c += 8) st_na(reinterpret_cast<float4*>(dst + (size_t)c * NCELL) + lane, make_float4(z, z, z, z));
  }
}

template <typename F>
static void run(const char* name, F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaDeviceSynchronize();
  float best = 1e9f, sum = 0.f;
  for (int i = 0; i < 10; ++i) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best; sum += ms;
  }
  const double bytes = (double)B * C * NCELL * 4;
  printf("%-44s avg %.4f ms  best %.4f ms  %.0f GB/s (avg)  %s\n", name, sum / 10, best, bytes / (sum / 10) / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; int* map;
  const size_t n = (size_t)B * C * NCELL;
  cudaMalloc(&out, n * 4);
  cudaMalloc(&map, (size_t)B * NCELL * 4);
  cudaMemset(map, 0xff, (size_t)B * NCELL * 4);
  const int grid_tiles = B * (NCELL / 128);
  run("A linear float4 fill", [&] { k_linear<<<148 * 16, 256>>>((float4*)out, n / 4); });
  run("B tile pattern, st.L1::no_allocate", [&] { k_tile<false, true><<<grid_tiles, 256>>>(out, map); });
  run("D tile pattern, plain st.global", [&] { k_tile<false, false><<<grid_tiles, 256>>>(out, map); });
  run("C tile pattern + map load + barrier", [&] { k_tile<true, true><<<grid_tiles, 256>>>(out, map); });
  run("E persistent tiles (6/SM), map prefetched", [&] { k_persistent<<<148 * 6, 256>>>(out, map); });
  run("E persistent tiles (4/SM), map prefetched", [&] { k_persistent<<<148 * 4, 256>>>(out, map); });
  run("E persistent tiles (8/SM), map prefetched", [&] { k_persistent<<<148 * 8, 256>>>(out, map); });
  return 0;
}
