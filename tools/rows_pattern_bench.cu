// Write-pattern microbenchmark behind vx_bins_kernel<decorate> (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/rows_pattern_bench tools/rows_pattern_bench.cu
// Output = 1,116,000 decorated pillar rows of 60 x 9 float32 (2,160 B each, 2.41 GB), written the way the bins
// kernel writes them - one CTA of 256 threads per 128 consecutive rows, a warp per pair of rows, the zero tail
// of a row from registers and its live head (n x 36 B) after it - but with NOTHING else in the kernel: no ranking,
// no gathers, no decoration.  Variants:
//   A  linear float4 fill of the same bytes                                 (the write ceiling)
//   B  the bins pattern, 4 CTAs/SM (52 KB of shared memory reserved, as in the library)
//   C  the same, 8 CTAs/SM (no shared memory)
//   D  the bins pattern with every row written as ONE contiguous sweep (head and tail by the same 32 lanes, in order)
//   E  B with a dependent 1 us "gather" (a global load chain) in front of every head
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ROWS 1116000
#define PER 540           // floats per row
#define NV 128

__device__ __forceinline__ void st_na(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void k_linear(float4* out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    st_na(out + i, make_float4(0.f, 0.f, 0.f, 0.f));
}

__device__ __forceinline__ int row_points(int row) { return 1 + (int)(((unsigned)row * 2654435761u) >> 29) + ((row & 15) == 0 ? 9 : 0); }  // 1..8 (+9): mean 5.1

template <int MODE>
__global__ void __launch_bounds__(256) k_rows(float* out, const float4* pts, int npts, int* num, int4* coords, int delay_clk) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * NV;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE >= 4) {   // the slot ranking of the bins kernel: ~6 us in which the CTA stores nothing
    const long long t0 = clock64();
    while (clock64() - t0 < delay_clk) {}
    __syncthreads();
  }
  for (int v = warp * 2; v < NV; v += 16) {
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + v + h;
      if (row >= ROWS) return;
      const int n = row_points(row);
      const int nd = n * 9;
      float4* d4 = reinterpret_cast<float4*>(out + (size_t)row * PER);
      float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
      if (MODE == 1) {   // one contiguous sweep
        for (int j = lane; j < PER / 4; j += 32) st_na(d4 + j, j < ((nd + 3) >> 2) ? val : z4);
        continue;
      }
      for (int j = ((nd + 3) >> 2) + lane; j < PER / 4; j += 32) st_na(d4 + j, z4);
      if (MODE == 2) {   // a dependent load chain in front of the head (stands in for the point gather)
        int idx = (row * 97 + lane) % npts;
        float4 q = __ldg(pts + idx);
        idx = (int)(fabsf(q.x) * 1e6f) % npts;
        q = __ldg(pts + idx);
        val.x += q.y;
      }
      if (MODE == 3 || MODE == 5) {
        if (lane == 0) { num[row] = n; coords[row] = make_int4(0, 0, row & 255, row >> 8); }
      }
      if (MODE == 5) {   // heads through a shared-memory stage, as the library does
        float* st = sm + warp * 544;
        for (int i = lane; i < nd; i += 32) st[i] = val.x + i;
        __syncwarp();
        for (int j = lane; j < ((nd + 3) >> 2); j += 32) st_na(d4 + j, reinterpret_cast<float4*>(st)[j]);
        __syncwarp();
      } else {
        for (int j = lane; j < ((nd + 3) >> 2); j += 32) st_na(d4 + j, val);
      }
    }
  }
  if (npts < 0 && sm[0] == 123.f) out[0] = 0.f;   // keeps the shared-memory reservation alive
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
  const double bytes = (double)ROWS * PER * 4;
  printf("%-64s mean %.3f ms  best %.3f ms  %.2f TB/s   %s\n", name, sum / 10, best, bytes / (sum / 10 * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; float4* pts; int* num; int4* coords;
  const size_t n = (size_t)ROWS * PER;
  const int npts = 6800000;
  cudaMalloc(&out, n * 4);
  cudaMalloc(&pts, (size_t)npts * 16);
  cudaMemset(pts, 0x3c, (size_t)npts * 16);
  cudaMalloc(&num, (size_t)ROWS * 4);
  cudaMalloc(&coords, (size_t)ROWS * 16);
  cudaFuncSetAttribute(k_rows<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248);
  cudaFuncSetAttribute(k_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248);
  cudaFuncSetAttribute(k_rows<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int grid = (ROWS + NV - 1) / NV;
  cudaFuncSetAttribute(k_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248);
  cudaFuncSetAttribute(k_rows<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248);
  run("A linear float4 fill", [&] { k_linear<<<148 * 16, 256>>>((float4*)out, n / 4); });
  run("B bins pattern (tail, then head), 4 CTAs/SM", [&] { k_rows<0><<<grid, 256, 53248>>>(out, pts, npts, num, coords, 0); });
  run("C bins pattern, 8 CTAs/SM", [&] { k_rows<0><<<grid, 256, 0>>>(out, pts, npts, num, coords, 0); });
  run("D rows as one contiguous sweep, 8 CTAs/SM", [&] { k_rows<1><<<grid, 256, 0>>>(out, pts, npts, num, coords, 0); });
  run("E bins pattern + dependent load chain before the head, 4 CTAs/SM", [&] { k_rows<2><<<grid, 256, 53248>>>(out, pts, npts, num, coords, 0); });
  run("F bins pattern + num_points / coords stores, 4 CTAs/SM", [&] { k_rows<3><<<grid, 256, 53248>>>(out, pts, npts, num, coords, 0); });
  for (int d : {4000, 8000, 12000, 20000}) {
    char name[96];
    snprintf(name, sizeof(name), "G bins pattern behind %d idle clocks per CTA, 4 CTAs/SM", d);
    run(name, [&] { k_rows<4><<<grid, 256, 53248>>>(out, pts, npts, num, coords, d); });
  }
  run("H = G(10000) + small stores + staged heads, 4 CTAs/SM", [&] { k_rows<5><<<grid, 256, 53248>>>(out, pts, npts, num, coords, 10000); });
  run("H at 3 CTAs/SM", [&] { k_rows<5><<<grid, 256, 70000>>>(out, pts, npts, num, coords, 10000); });
  return 0;
}
