// Block-filtering voxelizer variant: VoxelGeneratorV2(block_filtering=True, block_factor,
// block_size, height_threshold) as configured by second/second/configs/nuscenes/all.fhd.config:9-12 and passed at
// second/second/builder/voxel_builder.py:28-31.  PARITY UNPINNED: the arithmetic lives in spconv 1.x
// (points_to_voxel_3d_with_filtering), which is not in the reference tree; the rule restated in
// oracle/voxel_oracle.c (lvo_points_to_voxel_filtered) is what these kernels reproduce bit for bit:
//   every stored point (slot t < num_points[v]) updates the z-min / z-max of its (y / bf, x / bf) block;
//   a voxel survives iff over the block_size x block_size window of blocks around its own
//   height = max - min satisfies height > height_threshold && height < height_high_threshold;
//   the survivors keep their first-come order.
//
// The filter runs behind the ordinary voxelizer (lv_voxel.cu), on its padded per-frame output: the stored points
// of a voxel are exactly the points that updated the block ranges in the serial loop.
//   VF1 vf_fill      block ranges <- (+99999999, -99999999) as order-preserving u32 keys
//   VF2 vf_minmax    one thread per stored slot: atomicMin / atomicMax of the z key (L2 atomics)
//   VF3 vf_mask      one thread per voxel: window scan -> keep flag
//   VF4 vf_scan      one CTA per frame: stable exclusive scan of the flags -> destination row, new voxel_num
//   VF5 vf_compact   one warp per surviving row: 128-bit streaming copy into the caller's arrays (+ zero tail)
#include "lv_common.cuh"

namespace {

constexpr int VF_THREADS = 256;

__device__ __forceinline__ unsigned vf_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float vf_unkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct FilterParams {
  const float* voxels;        // (F, V, T, C)
  const int32_t* coords;      // (F, V, 3) zyx
  const int32_t* num_points;  // (F, V)
  const int32_t* voxel_num;   // (F)
  float* out_voxels;
  int32_t* out_coords;
  int32_t* out_num_points;
  int32_t* out_voxel_num;
  int32_t* out_mask;          // optional (F, V): keep flags over the unfiltered voxels
  unsigned* mins;             // (frames in flight, BH, BW) keys
  unsigned* maxs;
  int32_t* dst;               // (F, V) destination row or -1
  int V, T, C, BH, BW, bf, bs, zero_tail;
  float lo_thr, hi_thr;
  int f0;                     // first frame of this sub-batch
};

__global__ void __launch_bounds__(VF_THREADS) vf_fill_kernel(unsigned* mins, unsigned* maxs, int64_t n) {
  const unsigned kmin = vf_key(99999999.0f), kmax = vf_key(-99999999.0f);
  for (int64_t i = (int64_t)blockIdx.x * VF_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * VF_THREADS) {
    mins[i] = kmin;
    maxs[i] = kmax;
  }
}

// grid (slot blocks, frames of the sub-batch)
template <bool C4>
__global__ void __launch_bounds__(VF_THREADS) vf_minmax_kernel(FilterParams p) {
  const int f = p.f0 + blockIdx.y;
  int vnum = p.voxel_num[f];
  vnum = vnum < 0 ? 0 : (vnum > p.V ? p.V : vnum);
  const int64_t slots = (int64_t)vnum * p.T;
  const int64_t base = (int64_t)f * p.V;
  unsigned* mins = p.mins + (int64_t)blockIdx.y * p.BH * p.BW;
  unsigned* maxs = p.maxs + (int64_t)blockIdx.y * p.BH * p.BW;
  for (int64_t s = (int64_t)blockIdx.x * VF_THREADS + threadIdx.x; s < slots; s += (int64_t)gridDim.x * VF_THREADS) {
    const int v = (int)(s / p.T), t = (int)(s - (int64_t)v * p.T);
    float z;
    if (C4) z = lv_ld_stream_f4(reinterpret_cast<const float4*>(p.voxels) + base * p.T + s).z;
    else z = p.voxels[(base * p.T + s) * p.C + 2];
    if (t >= p.num_points[base + v]) continue;
    const int32_t* c = p.coords + (base + v) * 3;
    const int b = (c[1] / p.bf) * p.BW + c[2] / p.bf;
    const unsigned k = vf_key(z);
    atomicMin(mins + b, k);
    atomicMax(maxs + b, k);
  }
}

__global__ void __launch_bounds__(VF_THREADS) vf_mask_kernel(FilterParams p) {
  const int f = p.f0 + blockIdx.y;
  int vnum = p.voxel_num[f];
  vnum = vnum < 0 ? 0 : (vnum > p.V ? p.V : vnum);
  const int64_t base = (int64_t)f * p.V;
  const unsigned* mins = p.mins + (int64_t)blockIdx.y * p.BH * p.BW;
  const unsigned* maxs = p.maxs + (int64_t)blockIdx.y * p.BH * p.BW;
  for (int v = blockIdx.x * VF_THREADS + threadIdx.x; v < vnum; v += gridDim.x * VF_THREADS) {
    const int32_t* c = p.coords + (base + v) * 3;
    const int by = c[1] / p.bf, bx = c[2] / p.bf;
    int y0 = by - p.bs / 2, y1 = by + p.bs - p.bs / 2, x0 = bx - p.bs / 2, x1 = bx + p.bs - p.bs / 2;
    y0 = y0 < 0 ? 0 : y0;
    x0 = x0 < 0 ? 0 : x0;
    y1 = y1 > p.BH ? p.BH : y1;
    x1 = x1 > p.BW ? p.BW : x1;
    unsigned kmn = mins[by * p.BW + bx], kmx = maxs[by * p.BW + bx];
    for (int j = y0; j < y1; ++j)
      for (int k = x0; k < x1; ++k) {
        kmn = min(kmn, mins[j * p.BW + k]);
        kmx = max(kmx, maxs[j * p.BW + k]);
      }
    const float height = __fsub_rn(vf_unkey(kmx), vf_unkey(kmn));
    p.dst[base + v] = (height > p.lo_thr && height < p.hi_thr) ? 1 : 0;
  }
}

// one CTA of 1024 threads per frame: flags -> destination rows (stable), kept count.  Warp w owns a contiguous
// span of voxels and walks it in coalesced rounds of 32 (ballot + popc), so one scan of the 32 warp totals orders
// everything.
__global__ void __launch_bounds__(1024) vf_scan_kernel(FilterParams p) {
  __shared__ int s_warp[32];
  const int f = p.f0 + blockIdx.x;
  int vnum = p.voxel_num[f];
  vnum = vnum < 0 ? 0 : (vnum > p.V ? p.V : vnum);
  const int64_t base = (int64_t)f * p.V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int span = (((vnum + 31) >> 5) + 31) & ~31;          // voxels per warp, a multiple of 32
  const int lo = min(warp * span, vnum), hi = min(lo + span, vnum);
  int mine = 0;
#pragma unroll 4
  for (int v0 = lo; v0 < hi; v0 += 32) {
    const int v = v0 + lane;
    mine += __popc(__ballot_sync(0xffffffffu, v < hi && p.dst[base + v] != 0));
  }
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  int run = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    const int t = s_warp[w];
    if (w < warp) run += t;
    total += t;
  }
#pragma unroll 4
  for (int v0 = lo; v0 < hi; v0 += 32) {
    const int v = v0 + lane;
    const int keep = (v < hi && p.dst[base + v] != 0) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (v < hi) {
      p.dst[base + v] = keep ? run + __popc(bal & ((1u << lane) - 1u)) : -1;
      if (p.out_mask) p.out_mask[base + v] = keep;
    }
    run += __popc(bal);
  }
  if (p.out_mask)
    for (int v = vnum + threadIdx.x; v < p.V; v += 1024) p.out_mask[base + v] = 0;
  if (threadIdx.x == 0) p.out_voxel_num[f] = total;
}

// grid (row blocks, frames): a warp per source row
__global__ void __launch_bounds__(VF_THREADS) vf_compact_kernel(FilterParams p, int vec4) {
  const int f = p.f0 + blockIdx.y;
  int vnum = p.voxel_num[f];
  vnum = vnum < 0 ? 0 : (vnum > p.V ? p.V : vnum);
  const int64_t base = (int64_t)f * p.V;
  const int per = p.T * p.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = VF_THREADS / 32;
  for (int v = blockIdx.x * wpb + warp; v < vnum; v += gridDim.x * wpb) {
    const int d = p.dst[base + v];
    if (d < 0) continue;
    const float* src = p.voxels + (base + v) * per;
    float* out = p.out_voxels + (base + d) * per;
    if (vec4) {
      for (int i = lane; i < per / 4; i += 32)
        lv_st_stream_f4(reinterpret_cast<float4*>(out) + i, lv_ld_stream_f4(reinterpret_cast<const float4*>(src) + i));
    } else {
      for (int i = lane; i < per; i += 32) out[i] = src[i];
    }
    if (lane < 3) p.out_coords[(base + d) * 3 + lane] = p.coords[(base + v) * 3 + lane];
    if (lane == 3) p.out_num_points[base + d] = p.num_points[base + v];
  }
  if (!p.zero_tail) return;
  // rows [kept, V): generate_multi_gpu padding.  The kept count is final (vf_scan ran before this launch).
  const int kept = p.out_voxel_num[f];
  for (int v = kept + blockIdx.x * wpb + warp; v < p.V; v += gridDim.x * wpb) {
    float* out = p.out_voxels + (base + v) * per;
    for (int i = lane; i < per; i += 32) out[i] = 0.f;
    if (lane < 3) p.out_coords[(base + v) * 3 + lane] = 0;
    if (lane == 3) p.out_num_points[base + v] = 0;
  }
}

}  // namespace

extern "C" int lv_voxel_block_filter(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                                     int32_t n_frames, const float* d_voxels, const int32_t* d_coords,
                                     const int32_t* d_num_points, const int32_t* d_voxel_num, float* d_out_voxels,
                                     int32_t* d_out_coords, int32_t* d_out_num_points, int32_t* d_out_voxel_num,
                                     int32_t* d_out_mask, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_voxel_block_filter: null handle");
  LV_REQUIRE(cfg && flt, "lv_voxel_block_filter: null config");
  LV_REQUIRE(n_frames >= 0, "lv_voxel_block_filter: negative frame count");
  LV_REQUIRE(cfg->num_features >= 3 && cfg->max_points > 0 && cfg->max_voxels > 0, "lv_voxel_block_filter: bad config");
  LV_REQUIRE(flt->block_factor > 0 && flt->block_size > 0, "lv_voxel_block_filter: block_factor and block_size must be > 0");
  int32_t grid[3];
  LV_CHECK(lv_voxel_grid_size(cfg, grid));
  LV_REQUIRE(grid[0] > 0 && grid[1] > 0, "lv_voxel_block_filter: empty grid");
  // spconv asserts the same in VoxelGeneratorV2.__init__
  LV_REQUIRE(grid[0] % flt->block_factor == 0 && grid[1] % flt->block_factor == 0,
             "lv_voxel_block_filter: grid %d x %d is not divisible by block_factor %d", grid[0], grid[1], flt->block_factor);
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_coords && d_num_points && d_voxel_num && d_out_voxels && d_out_coords && d_out_num_points &&
                 d_out_voxel_num, "lv_voxel_block_filter: null pointer");
  LV_REQUIRE(d_voxels != d_out_voxels && d_coords != d_out_coords && d_num_points != d_out_num_points,
             "lv_voxel_block_filter: the compaction is not in place, outputs must not alias inputs");
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));

  FilterParams p;
  memset(&p, 0, sizeof(p));
  p.voxels = d_voxels; p.coords = d_coords; p.num_points = d_num_points; p.voxel_num = d_voxel_num;
  p.out_voxels = d_out_voxels; p.out_coords = d_out_coords; p.out_num_points = d_out_num_points;
  p.out_voxel_num = d_out_voxel_num; p.out_mask = d_out_mask;
  p.V = cfg->max_voxels; p.T = cfg->max_points; p.C = cfg->num_features;
  p.BH = grid[1] / flt->block_factor; p.BW = grid[0] / flt->block_factor;
  p.bf = flt->block_factor; p.bs = flt->block_size; p.zero_tail = cfg->zero_tail;
  p.lo_thr = flt->height_threshold; p.hi_thr = flt->height_high_threshold;
  const int64_t blocks = (int64_t)p.BH * p.BW;
  LV_REQUIRE(blocks < (1ll << 28), "lv_voxel_block_filter: block grid too large");
  int64_t fif = (128ll << 20) / (blocks * 8);   // block ranges of the frames in flight: <= 128 MB
  if (fif < 1) fif = 1;
  if (fif > n_frames) fif = n_frames;
  LV_CHECK(h->flt_ranges.ensure((size_t)fif * blocks * 8, stream));
  LV_CHECK(h->flt_dst.ensure((size_t)n_frames * p.V * 4, stream));
  p.mins = h->flt_ranges.as<unsigned>();
  p.maxs = p.mins + fif * blocks;
  p.dst = h->flt_dst.as<int32_t>();
  const bool c4 = p.C == 4 && (reinterpret_cast<uintptr_t>(d_voxels) & 15) == 0;
  const int vec4 = ((p.T * p.C) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(d_voxels) | reinterpret_cast<uintptr_t>(d_out_voxels)) & 15) == 0;
  for (int f0 = 0; f0 < n_frames; f0 += (int)fif) {
    const int nf = (int)(f0 + fif <= n_frames ? fif : n_frames - f0);
    p.f0 = f0;
    int gx = (int)lv_div_up(lv_div_up((int64_t)nf * blocks, VF_THREADS), 4);
    if (gx > h->num_sms * 8) gx = h->num_sms * 8;
    vf_fill_kernel<<<gx, VF_THREADS, 0, stream>>>(p.mins, p.maxs, (int64_t)nf * blocks);
    LV_LAUNCH_CHECK(h);
    int sx = (int)lv_div_up((int64_t)p.V * p.T, VF_THREADS * 4);
    const int cap = (int)lv_div_up((int64_t)h->num_sms * 8, nf);
    if (sx > cap) sx = cap;
    if (sx < 1) sx = 1;
    if (c4) vf_minmax_kernel<true><<<dim3((unsigned)sx, (unsigned)nf), VF_THREADS, 0, stream>>>(p);
    else vf_minmax_kernel<false><<<dim3((unsigned)sx, (unsigned)nf), VF_THREADS, 0, stream>>>(p);
    LV_LAUNCH_CHECK(h);
    int mx = (int)lv_div_up(p.V, VF_THREADS);
    if (mx > cap) mx = cap;
    if (mx < 1) mx = 1;
    vf_mask_kernel<<<dim3((unsigned)mx, (unsigned)nf), VF_THREADS, 0, stream>>>(p);
    LV_LAUNCH_CHECK(h);
    vf_scan_kernel<<<nf, 1024, 0, stream>>>(p);
    LV_LAUNCH_CHECK(h);
    int cx = (int)lv_div_up(p.V, VF_THREADS / 32);
    if (cx > cap) cx = cap;
    if (cx < 1) cx = 1;
    vf_compact_kernel<<<dim3((unsigned)cx, (unsigned)nf), VF_THREADS, 0, stream>>>(p, vec4);
    LV_LAUNCH_CHECK(h);
  }
  return LV_OK;
}

extern "C" int lv_voxelize_filtered(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                                    const float* d_points, int32_t n_frames, const int64_t* h_frame_offsets,
                                    float* d_voxels, int32_t* d_coords, int32_t* d_num_points, int32_t* d_voxel_num,
                                    int32_t* d_mask, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_voxelize_filtered: null handle");
  LV_REQUIRE(cfg && flt, "lv_voxelize_filtered: null config");
  LV_REQUIRE(cfg->overflow_mode == LV_OVERFLOW_CONTINUE,
             "lv_voxelize_filtered: block filtering exists only with the spconv >= 1.1 `continue` overflow rule");
  LV_REQUIRE(n_frames >= 0, "lv_voxelize_filtered: negative frame count");
  if (n_frames == 0) return LV_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  const size_t F = n_frames, V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features;
  LV_CHECK(h->flt_tmp[0].ensure(F * V * T * C * 4, stream));
  LV_CHECK(h->flt_tmp[1].ensure(F * V * 3 * 4, stream));
  LV_CHECK(h->flt_tmp[2].ensure(F * V * 4, stream));
  LV_CHECK(h->flt_tmp[3].ensure(F * 4, stream));
  lv_voxel_config inner = *cfg;
  inner.zero_tail = 0;  // the filter reads only rows < voxel_num
  LV_CHECK(lv_voxelize(h, &inner, d_points, n_frames, h_frame_offsets, h->flt_tmp[0].as<float>(),
                       h->flt_tmp[1].as<int32_t>(), h->flt_tmp[2].as<int32_t>(), h->flt_tmp[3].as<int32_t>(), stream_));
  return lv_voxel_block_filter(h, cfg, flt, n_frames, h->flt_tmp[0].as<float>(), h->flt_tmp[1].as<int32_t>(),
                               h->flt_tmp[2].as<int32_t>(), h->flt_tmp[3].as<int32_t>(), d_voxels, d_coords,
                               d_num_points, d_voxel_num, d_mask, stream_);
}
