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

// ---- TMA store variants: the SM only issues copies, the copy engine streams the rows out of shared memory.
// F: per tile, one warp issues 64 one-dimensional bulk stores of 512 B (cp.async.bulk.global.shared::cta) from a
//    zero buffer, behind the map load.  G: per tile ONE two-dimensional tensor store (box 128 cells x 64 rows =
//    32 KB) from a ring of NBUF shared-memory tiles; a buffer is reused once its bulk group has been read.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store_1d(void* g, const void* s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(128) k_tma1d(float* out, const int* map) {
  __shared__ __align__(128) float zero[128];
  const int tiles = NCELL / 128, total = tiles * B;
  zero[threadIdx.x] = 0.f;
  fence_async_smem();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // every warp owns a quarter of the CTA's tiles
  for (int tile = blockIdx.x * 4 + warp; tile < total; tile += gridDim.x * 4) {
    const int b = tile / tiles, t = tile - b * tiles;
    const int4 m = *reinterpret_cast<const int4*>(map + (size_t)b * NCELL + t * 128 + lane * 4);
    const int any = __any_sync(0xffffffffu, (m.x & m.y & m.z & m.w) >= 0);
    float* dst = out + (size_t)b * C * NCELL + (size_t)t * 128;
    if (!any || true) {
      bulk_store_1d(dst + (size_t)lane * NCELL, zero, 512);
      bulk_store_1d(dst + (size_t)(lane + 32) * NCELL, zero, 512);
      bulk_commit();
    }
  }
  bulk_wait_read<0>();
}

#include <cuda.h>
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* tm, int x, int y, const void* s) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(x), "r"(y),
               "r"(smem_u32(s))
               : "memory");
}

template <int NBUF, int ROWS>
__global__ void __launch_bounds__(128) k_tma2d(const __grid_constant__ CUtensorMap tm, const int* map) {
  extern __shared__ __align__(128) float buf[];   // NBUF x ROWS x 128 floats
  const int per = ROWS * 128;
  const int tiles = NCELL / 128, total = tiles * B * (C / ROWS);
  for (int i = threadIdx.x; i < NBUF * per; i += 128) buf[i] = 0.f;
  fence_async_smem();
  __syncthreads();
  int it = 0;
  for (int job = blockIdx.x; job < total; job += gridDim.x, ++it) {
    const int tile = job / (C / ROWS), part = job - tile * (C / ROWS);
    const int b = tile / tiles, t = tile - b * tiles;
    float* cur = buf + (it % NBUF) * per;
    const int mv = threadIdx.x < 128 ? map[(size_t)b * NCELL + t * 128 + threadIdx.x] : -1;
    if (threadIdx.x == 0) bulk_wait_read<NBUF - 1>();     // the buffer's previous store has been read
    __syncthreads();
    if (mv >= 0) cur[threadIdx.x] = 1.f;                  // stand-in for the feature scatter
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tensor_store_2d(&tm, t * 128, b * C + part * ROWS, cur);
      bulk_commit();
    }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();
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
  for (int per_sm : {2, 4, 8, 16}) {
    char name[96];
    snprintf(name, sizeof(name), "F 1D bulk stores 64 x 512 B, %d CTAs/SM", per_sm);
    run(name, [&] { k_tma1d<<<148 * per_sm, 128>>>(out, map); });
  }
  {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    auto make = [&](int rows) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {NCELL, (cuuint64_t)B * C};
      cuuint64_t strides[1] = {(cuuint64_t)NCELL * 4};
      cuuint32_t box[2] = {128, (cuuint32_t)rows};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return tm;
    };
    CUtensorMap tm64 = make(64), tm16 = make(16);
    cudaFuncSetAttribute(k_tma2d<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 64 * 512);
    cudaFuncSetAttribute(k_tma2d<6, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 64 * 512);
    cudaFuncSetAttribute(k_tma2d<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * 512);
    run("G 2D tensor store 32 KB, 3 buffers, 2 CTAs/SM", [&] { k_tma2d<3, 64><<<148 * 2, 128, 3 * 64 * 512>>>(tm64, map); });
    run("G 2D tensor store 32 KB, 6 buffers, 1 CTA/SM", [&] { k_tma2d<6, 64><<<148, 128, 6 * 64 * 512>>>(tm64, map); });
    run("G 2D tensor store 8 KB, 4 buffers, 6 CTAs/SM", [&] { k_tma2d<4, 16><<<148 * 6, 128, 4 * 16 * 512>>>(tm16, map); });
  }
  return 0;
}
