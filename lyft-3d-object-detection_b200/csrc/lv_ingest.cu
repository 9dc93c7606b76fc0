// Multi-sweep ingest (sm_100a): raw 5-float sweeps -> one (N,4|5) float32 cloud in the key
// frame, the step immediately in front of VoxelGenerator.generate.
//
// LV_INGEST_SECOND  second/second/data/nuscenes_dataset.py:196-223
//     intensity /= 255 (float32); sweeps other than the key sweep: xyz = xyz @ R.T computed in
//     float64 and stored to float32, THEN xyz += T computed in float64 and stored to float32
//     (two roundings, :216-218); column 4 := ts - sweep_ts (0 for the key sweep); the result
//     keeps columns [0,1,2,4] (out_cols == 4) or all five.
// LV_INGEST_DEVKIT  nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py:99-137
//     every sweep: xyz = (M4x4 . [x;y;z;1])[:3] in float64, stored once (:195); remove_close
//     (:153-165): rows with |x| < r and |y| < r are dropped - here they become NaN rows, which
//     every downstream kernel ignores, so row counts stay static; columns [x,y,z,intensity]
//     (+ time lag as column 4 when out_cols == 5).
//
// HBM layout: raw (N_total, 5) float32 rows, sweeps back to back (host offset table); out
// (N_total, out_cols) float32.  One CTA = one tile of 1024 rows: a full, 16-byte aligned tile
// is staged in shared memory by one TMA bulk copy (20 KB) and read back conflict-free (row
// stride 5 words); the output leaves as 128-bit stores when out_cols == 4.
// Algorithmic bytes: 20 read + 4*out_cols written per point.
#include "lv_common.cuh"

#define ING_THREADS 256
#define ING_ITEMS 4
#define ING_TILE (ING_THREADS * ING_ITEMS)

struct IngestParams {
  const float* raw;
  float* out;
  int64_t n;
  const int64_t* sweep_off;   // device [S+1]
  const double* tm;           // device [S][12]: SECOND: R row-major (9) + T (3); DEVKIT: rows 0..2 of the 4x4
  const float* lag;           // device [S]
  const uint8_t* has_tm;      // device [S]
  int n_sweeps, mode, out_cols, use_tma;
  float close_radius;         // < 0: off
};

__device__ __forceinline__ int ing_find(const int64_t* __restrict__ offs, int lo, int hi, int64_t i) {
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(offs + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(ING_THREADS) ingest_sweeps_kernel(IngestParams p) {
  __shared__ __align__(128) float tile[ING_TILE * 5];
  __shared__ __align__(8) uint64_t bar;
  const int64_t row0 = (int64_t)blockIdx.x * ING_TILE;
  const float* src = p.raw + row0 * 5;
  const bool staged = p.use_tma && row0 + ING_TILE <= p.n;   // tile starts are 16-byte aligned with the base
  if (staged) {
    if (threadIdx.x == 0) {
      lv_mbar_init(&bar, 1);
      lv_mbar_init_fence();
      lv_mbar_expect_tx(&bar, ING_TILE * 20);
      lv_tma_load_1d(tile, src, ING_TILE * 20, &bar);
    }
    __syncthreads();
  }
  // sweeps covered by the tile (usually one)
  const int64_t last = (row0 + ING_TILE < p.n ? row0 + ING_TILE : p.n) - 1;
  const int sa = ing_find(p.sweep_off, 0, p.n_sweeps, row0);
  const int sb = __ldg(p.sweep_off + sa + 1) > last ? sa : ing_find(p.sweep_off, sa, p.n_sweeps, last);
  if (staged) lv_mbar_wait(&bar, 0);
#pragma unroll
  for (int r = 0; r < ING_ITEMS; ++r) {
    const int li = r * ING_THREADS + threadIdx.x;
    const int64_t i = row0 + li;
    if (i >= p.n) break;
    float v[5];
    if (staged) {
#pragma unroll
      for (int c = 0; c < 5; ++c) v[c] = tile[li * 5 + c];
    } else {
#pragma unroll
      for (int c = 0; c < 5; ++c) v[c] = __ldg(src + li * 5 + c);
    }
    const int s = sa == sb ? sa : ing_find(p.sweep_off, sa, sb + 1, i);
    float x = v[0], y = v[1], z = v[2];
    if (__ldg(p.has_tm + s)) {
      const double* M = p.tm + (size_t)s * 12;
      const double X = x, Y = y, Z = z;
      if (p.mode == LV_INGEST_SECOND) {
        // :216-217  (N,3) f32 @ (3,3) f64 -> f64, stored to f32;  :218  += T in f64, stored to f32
        x = (float)fma(__ldg(M + 2), Z, fma(__ldg(M + 1), Y, __ldg(M + 0) * X));
        y = (float)fma(__ldg(M + 5), Z, fma(__ldg(M + 4), Y, __ldg(M + 3) * X));
        z = (float)fma(__ldg(M + 8), Z, fma(__ldg(M + 7), Y, __ldg(M + 6) * X));
        x = (float)__dadd_rn((double)x, __ldg(M + 9));
        y = (float)__dadd_rn((double)y, __ldg(M + 10));
        z = (float)__dadd_rn((double)z, __ldg(M + 11));
      } else {
        // data_classes.py:195  M[:3,:4] . [x;y;z;1] in f64, one store
        x = (float)(fma(__ldg(M + 2), Z, fma(__ldg(M + 1), Y, __ldg(M + 0) * X)) + __ldg(M + 3));
        y = (float)(fma(__ldg(M + 6), Z, fma(__ldg(M + 5), Y, __ldg(M + 4) * X)) + __ldg(M + 7));
        z = (float)(fma(__ldg(M + 10), Z, fma(__ldg(M + 9), Y, __ldg(M + 8) * X)) + __ldg(M + 11));
      }
    }
    const float lag = __ldg(p.lag + s);
    float o[5];
    if (p.mode == LV_INGEST_SECOND) {
      const float inten = __fdiv_rn(v[3], 255.0f);   // :203 / :215
      o[0] = x; o[1] = y; o[2] = z;
      if (p.out_cols == 4) { o[3] = lag; } else { o[3] = inten; o[4] = lag; }
    } else {
      if (p.close_radius >= 0.f && fabsf(x) < p.close_radius && fabsf(y) < p.close_radius) {
        x = y = z = __int_as_float(0x7fc00000);      // remove_close (:153-165): a NaN row
      }
      o[0] = x; o[1] = y; o[2] = z; o[3] = v[3]; o[4] = lag;
    }
    if (p.out_cols == 4) {
      lv_st_stream_f4(reinterpret_cast<float4*>(p.out) + i, make_float4(o[0], o[1], o[2], o[3]));
    } else {
      float* q = p.out + i * 5;
#pragma unroll
      for (int c = 0; c < 5; ++c) q[c] = o[c];
    }
  }
}

extern "C" int lv_ingest_sweeps(lv_handle* h, const float* d_raw, int32_t n_sweeps, const int64_t* h_sweep_offsets,
                                const double* h_sweep_tm, const uint8_t* h_has_tm, const float* h_time_lag,
                                int32_t mode, float close_radius, int32_t out_cols, float* d_out, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_ingest_sweeps: null handle");
  LV_REQUIRE(n_sweeps >= 0 && h_sweep_offsets, "lv_ingest_sweeps: bad sweep table");
  LV_REQUIRE(mode == LV_INGEST_SECOND || mode == LV_INGEST_DEVKIT, "lv_ingest_sweeps: bad mode %d", mode);
  LV_REQUIRE(out_cols == 4 || out_cols == 5, "lv_ingest_sweeps: out_cols must be 4 or 5, got %d", out_cols);
  if (n_sweeps == 0) return LV_OK;
  LV_REQUIRE(h_sweep_offsets[0] == 0, "lv_ingest_sweeps: sweep_offsets[0] must be 0");
  for (int s = 0; s < n_sweeps; ++s)
    LV_REQUIRE(h_sweep_offsets[s + 1] >= h_sweep_offsets[s], "lv_ingest_sweeps: sweep_offsets must be non-decreasing");
  const int64_t n = h_sweep_offsets[n_sweeps];
  if (n == 0) return LV_OK;
  LV_REQUIRE(d_raw && d_out && h_time_lag, "lv_ingest_sweeps: null pointer");
  LV_REQUIRE(out_cols != 4 || (reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "lv_ingest_sweeps: out must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  std::vector<double> tm((size_t)n_sweeps * 12, 0.0);
  std::vector<uint8_t> has((size_t)n_sweeps, 0);
  for (int s = 0; s < n_sweeps; ++s) {
    has[s] = (h_sweep_tm != nullptr) && (h_has_tm == nullptr || h_has_tm[s]);
    if (!has[s]) continue;
    const double* src = h_sweep_tm + (size_t)s * 16;   // row-major 4x4 per sweep
    if (mode == LV_INGEST_SECOND) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) tm[(size_t)s * 12 + r * 3 + c] = src[r * 4 + c];   // R
      for (int r = 0; r < 3; ++r) tm[(size_t)s * 12 + 9 + r] = src[r * 4 + 3];         // T
    } else {
      for (int k = 0; k < 12; ++k) tm[(size_t)s * 12 + k] = src[k];
    }
  }
  const void *d_off, *d_tm, *d_lag, *d_has;
  LV_CHECK(h->ing_offsets.sync(h_sweep_offsets, sizeof(int64_t) * (n_sweeps + 1), stream, &d_off));
  LV_CHECK(h->ing_tm.sync(tm.data(), sizeof(double) * tm.size(), stream, &d_tm));
  LV_CHECK(h->ing_lag.sync(h_time_lag, sizeof(float) * n_sweeps, stream, &d_lag));
  LV_CHECK(h->ing_has.sync(has.data(), has.size(), stream, &d_has));
  IngestParams p;
  p.raw = d_raw; p.out = d_out; p.n = n;
  p.sweep_off = (const int64_t*)d_off; p.tm = (const double*)d_tm; p.lag = (const float*)d_lag;
  p.has_tm = (const uint8_t*)d_has;
  p.n_sweeps = n_sweeps; p.mode = mode; p.out_cols = out_cols;
  p.use_tma = !h->disable_tma && (reinterpret_cast<uintptr_t>(d_raw) & 15) == 0;
  p.close_radius = close_radius;
  const int64_t tiles = lv_div_up(n, ING_TILE);
  LV_REQUIRE(tiles < (1ll << 31), "lv_ingest_sweeps: too many points");
  ingest_sweeps_kernel<<<(unsigned)tiles, ING_THREADS, 0, stream>>>(p);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}
