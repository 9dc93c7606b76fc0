// BEV rasteriser kernels (sm_100a).
//
// Replaces generating-dataset/generating_train_bev.py:47-104 (+ :213 quantise,
// + deeplab_v3_baseline/dataset/dataset.py:83-106 concat) and the sensor->car
// transform lyft_dataset_sdk/utils/data_classes.py:188-195.  Semantics in
// include/lyft_voxel.h and SURVEY.md Appendix A.1.
//
// Data layout in HBM
//   points   : (N_total, stride) float32, frames/sweeps ("segments") back to back
//   counts   : workspace u32 [frames_in_flight][S0*S1*S2], indexed (c1*S1 + c0)*S2 + c2,
//              all-zero between calls (the finalize kernel re-zeroes what it reads);
//              up to 192 MB in flight (128 frames of 336x336x3: 173 MB, more than the 126 MB L2 -
//              fewer, larger launches measured faster than L2 residency, see lv_bev_rasterize); only
//              the touched sectors ever move, tracked by a dirty bitmap (1 bit per 4 counts).
//   outputs  : per frame (S0,S1,S2) f32 raw / f32 normalised / u8, (S2+3,S0,S1) f32 CHW
//
// Kernels
//   bev_hist_kernel      one thread per point: optional fp64 4x4, fp64 un-fused affine,
//                        trunc, bounds, warp-aggregated (__match_any_sync) RED.ADD.
//                        Algorithmic bytes: stride*4 per point read.
//   bev_finalize_*       one pass over the counts of the frames in flight: emits every
//                        requested output and clears the counts.
//                        Algorithmic bytes: 4 (raw) [+4 norm] [+1 u8] [+ (S2+3)*4/S2 chw + 1 map] per cell.
#include <algorithm>
#include <vector>

#include "lv_common.cuh"

struct BevParams {
  const float* pts;
  int stride;
  const int64_t* seg_offsets;  // device, global segment table
  const int32_t* seg_frame;    // device or null
  const double* seg_tm;        // device or null (16 doubles per segment)
  int seg_lo, seg_hi;          // segments of this sub-batch [seg_lo, seg_hi)
  int64_t pt_begin, pt_end;    // points of this sub-batch
  int frame_base;              // first frame of this sub-batch
  double m0, m1, m2, t0, t1, t2;
  int S0, S1, S2;
  unsigned cells;
  unsigned* counts;
  int u16;                     // counts are 16-bit, two per word (every frame of the call has < 65,536 points, so a
                               // half can never carry into its neighbour): half the footprint - 128 frames of 336x336x3
                               // are 87 MB instead of 173 MB and stay in the 126 MB L2
  unsigned* dirty;             // 1 bit per 4 consecutive counts ("quad"): set by the first hit of a cell
  int use_tma;                 // point rows 16-byte aligned: full tiles are staged by TMA bulk copies
  // zero fill of the dense outputs of the sub-batch (pass A of the finalize), spread over the tiles of this kernel:
  // it depends on nothing, and the histogram leaves the store path idle (28 % of the DRAM peak, waiting for ATOMs)
  float4* z_raw;               // or null
  float4* z_norm;              // or null
  uchar4* z_u8;                // or null
  int64_t z_quads;             // quads (4 cells) to clear
  int z_per_tile;              // quads cleared per tile of points (ceil); 0 = the finalize kernel streams the zeros itself
};

__device__ __forceinline__ int bev_find_segment(const int64_t* __restrict__ offs, int lo, int hi, int64_t i) {
  // largest s in [lo, hi) with offs[s] <= i   (offs[lo] <= i < offs[hi] holds)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offs + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}
// the same for a warp-uniform i: a 32-ary search, one coalesced load per round (all lanes must call)
__device__ __forceinline__ int bev_find_segment_warp(const int64_t* __restrict__ offs, int lo, int hi, int64_t i, int lane) {
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int cand = lo + lane * step;
    const bool ok = cand < hi && __ldg(offs + cand) <= i;   // monotone in lane; lane 0 always holds
    const int j = 31 - __clz(__ballot_sync(0xffffffffu, ok));
    lo += j * step;
    hi = lo + step < hi ? lo + step : hi;
  }
  return lo;
}

#define BEV_THREADS 256
#define BEV_WARPS (BEV_THREADS / 32)
#define BEV_ITEMS 4
#define BEV_WTILE (32 * BEV_ITEMS)          // 128 points per warp
#define BEV_TILE (BEV_THREADS * BEV_ITEMS)  // 1024 points per tile
#define BEV_STAGES 4                         // tiles in flight per CTA (TMA pipeline depth)

// Persistent CTAs walk the tiles of the sub-batch (tile t = absolute points [t*1024, (t+1)*1024));
// a warp owns a contiguous span of 128 points of the tile.  STRIDE 4 / 5: tiles that lie
// entirely inside the sub-batch are staged in shared memory by one TMA bulk copy each
// (cp.async.bulk + mbarrier, four stages: the copies of the CTA's next three tiles are in
// flight while the current one is processed) and read back with conflict-free LDS (a row stride of 5 words
// is coprime with the 32 banks).  STRIDE 0: any row stride, direct global loads.
template <int STRIDE>
__global__ void __launch_bounds__(BEV_THREADS, 4) bev_hist_kernel(BevParams p) {
  extern __shared__ __align__(128) float tiles[];  // [BEV_STAGES][BEV_TILE * STRIDE]
  __shared__ __align__(8) uint64_t bar[BEV_STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t_lo = p.pt_begin / BEV_TILE, t_hi = (p.pt_end + BEV_TILE - 1) / BEV_TILE;
  const bool tma = STRIDE != 0 && p.use_tma;
  constexpr unsigned TILE_FLOATS = BEV_TILE * (STRIDE ? STRIDE : 1);
  if (tma) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < BEV_STAGES; ++s) lv_mbar_init(&bar[s], 1);
      lv_mbar_init_fence();
    }
    __syncthreads();
  }
  unsigned phase = 0;  // bit s = parity the next fill of stage s completes with
  int64_t t = t_lo + blockIdx.x;
  const int64_t t_step = gridDim.x;
  auto staged = [&](int64_t tt) { return tma && tt < t_hi && tt * BEV_TILE >= p.pt_begin && (tt + 1) * BEV_TILE <= p.pt_end; };
  auto issue = [&](int64_t tt, int s) {
    lv_mbar_expect_tx(&bar[s], TILE_FLOATS * 4);
    lv_tma_load_1d(tiles + (size_t)s * TILE_FLOATS, p.pts + tt * (int64_t)TILE_FLOATS, TILE_FLOATS * 4, &bar[s]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < BEV_STAGES - 1; ++s)
      if (staged(t + s * t_step)) issue(t + s * t_step, s);
  }
  // results of the ATOMs of the previous tile (was the cell empty?): looked at one tile later,
  // when they have long arrived
  unsigned pend_key[BEV_ITEMS], pend_old[BEV_ITEMS];
  // ... in fact two tiles later: most counts miss in L2 (173 MB of counts per 128 frames), so an ATOM's answer
  // takes a DRAM round trip (measured: BEV stage 0.135 -> 0.130 ms; answering with REDs only - an unconditional
  // second RED for the dirty bit - is slower, 0.145 ms)
  unsigned pend2_key[BEV_ITEMS], pend2_old[BEV_ITEMS];
#pragma unroll
  for (int r = 0; r < BEV_ITEMS; ++r) { pend2_key[r] = 0xffffffffu; pend2_old[r] = 1; }
#pragma unroll
  for (int r = 0; r < BEV_ITEMS; ++r) { pend_key[r] = 0xffffffffu; pend_old[r] = 1; }
  for (unsigned k = 0; t < t_hi; t += t_step, ++k) {
    const int s = k % BEV_STAGES;
    // stage (k-1) % STAGES was released by the barrier that ended the previous iteration
    if (threadIdx.x == 0 && staged(t + (BEV_STAGES - 1) * t_step))
      issue(t + (BEV_STAGES - 1) * t_step, (k + BEV_STAGES - 1) % BEV_STAGES);
    // this tile's share of the zero fill: stores that wait for nothing, issued before anything is waited for
    if (p.z_per_tile) {
      const int64_t q_lo = (t - t_lo) * p.z_per_tile;
      const int n = (int)(p.z_quads - q_lo < p.z_per_tile ? p.z_quads - q_lo : p.z_per_tile);
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = threadIdx.x; q < n; q += BEV_THREADS) {
        if (p.z_raw) lv_st_stream_f4(p.z_raw + q_lo + q, z4);
        if (p.z_norm) lv_st_stream_f4(p.z_norm + q_lo + q, z4);
        if (p.z_u8) p.z_u8[q_lo + q] = make_uchar4(0, 0, 0, 0);
      }
    }
    // segments of the warp's span (uniform across the warp), looked up while the copy is in flight
    const int64_t span0 = t * BEV_TILE + warp * BEV_WTILE;
    const int64_t v0 = span0 < p.pt_begin ? p.pt_begin : span0;
    const int64_t v1 = span0 + BEV_WTILE < p.pt_end ? span0 + BEV_WTILE : p.pt_end;  // valid points [v0, v1)
    int sa = p.seg_lo, sb = p.seg_lo;
    if (v0 < v1 && p.seg_hi - p.seg_lo > 1) {
      sa = bev_find_segment_warp(p.seg_offsets, p.seg_lo, p.seg_hi, v0, lane);
      sb = sa;
      if (__ldg(p.seg_offsets + sa + 1) < v1) sb = bev_find_segment_warp(p.seg_offsets, sa, p.seg_hi, v1 - 1, lane);
    }
    const bool in_smem = staged(t);
    if (in_smem) {
      lv_mbar_wait(&bar[s], (phase >> s) & 1);
      phase ^= 1u << s;
    }
    const float* buf = tiles + (size_t)s * TILE_FLOATS;
    // the thread's points of this tile, all loads in flight before the first one is used
    float px[BEV_ITEMS], py[BEV_ITEMS], pz[BEV_ITEMS];
#pragma unroll
    for (int r = 0; r < BEV_ITEMS; ++r) {
      const int li = warp * BEV_WTILE + r * 32 + lane;
      const int64_t i = t * BEV_TILE + li;
      px[r] = py[r] = pz[r] = 0.f;
      if (i >= v0 && i < v1) {
        if (STRIDE == 4) {
          float4 v;
          if (in_smem) v = reinterpret_cast<const float4*>(buf)[li];
          else v = lv_ld_stream_f4(reinterpret_cast<const float4*>(p.pts) + i);
          px[r] = v.x; py[r] = v.y; pz[r] = v.z;
        } else if (STRIDE != 0 && in_smem) {
          px[r] = buf[li * STRIDE]; py[r] = buf[li * STRIDE + 1]; pz[r] = buf[li * STRIDE + 2];
        } else {
          const float* q = p.pts + i * p.stride;
          px[r] = __ldg(q); py[r] = __ldg(q + 1); pz[r] = __ldg(q + 2);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < BEV_ITEMS; ++r) {
      const int li = warp * BEV_WTILE + r * 32 + lane;
      const int64_t i = t * BEV_TILE + li;
      unsigned key = 0xffffffffu;
      if (i >= v0 && i < v1) {
        float x = px[r], y = py[r], z = pz[r];
        int seg = sa;
        if (sb != sa) seg = bev_find_segment(p.seg_offsets, sa, sb + 1, i);
        if (p.seg_tm) {
          // PointCloud.transform (data_classes.py:195): float64 product, float32 store.
          const double* M = p.seg_tm + (size_t)seg * 16;
          const double X = x, Y = y, Z = z;
          double ax = fma(__ldg(M + 2), Z, fma(__ldg(M + 1), Y, __ldg(M + 0) * X)) + __ldg(M + 3);
          double ay = fma(__ldg(M + 6), Z, fma(__ldg(M + 5), Y, __ldg(M + 4) * X)) + __ldg(M + 7);
          double az = fma(__ldg(M + 10), Z, fma(__ldg(M + 9), Y, __ldg(M + 8) * X)) + __ldg(M + 11);
          x = (float)ax; y = (float)ay; z = (float)az;
        }
        // car_to_voxel_coords (generating_train_bev.py:47-82): un-fused fp64 multiply then add.
        const double u0 = __dadd_rn(__dmul_rn(p.m0, (double)x), p.t0);
        const double u1 = __dadd_rn(__dmul_rn(p.m1, (double)y), p.t1);
        const double u2 = __dadd_rn(__dmul_rn(p.m2, (double)z), p.t2);
        // trunc(u) in [0, S)  <=>  -1 < u < S ; NaN fails (np.intp(nan) is out of bounds).
        const bool in = (u0 > -1.0) && (u0 < (double)p.S0) && (u1 > -1.0) && (u1 < (double)p.S1) &&
                        (u2 > -1.0) && (u2 < (double)p.S2);
        if (in) {
          const int c0 = (int)u0, c1 = (int)u1, c2 = (int)u2;  // C truncation == np.intp
          const int frame = p.seg_frame ? __ldg(p.seg_frame + seg) : seg;
          key = (unsigned)(frame - p.frame_base) * p.cells + (unsigned)((c1 * p.S1 + c0) * p.S2 + c2);
        }
      }
      // warp-aggregated atomics: one ATOM per distinct cell per warp.  The first hit of a cell
      // (old == 0) marks the cell's quad in the dirty bitmap.
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (pend2_old[r] == 0) atomicOr(p.dirty + (pend2_key[r] >> 7), 1u << ((pend2_key[r] >> 2) & 31));
      pend2_key[r] = pend_key[r];
      pend2_old[r] = pend_old[r];
      pend_key[r] = 0xffffffffu;
      pend_old[r] = 1;
      if (key != 0xffffffffu && lane == (__ffs(peers) - 1)) {
        if (p.u16) {
          const unsigned sh = (key & 1u) << 4;
          pend_old[r] = (atomicAdd(p.counts + (key >> 1), (unsigned)__popc(peers) << sh) >> sh) & 0xffffu;
        } else {
          pend_old[r] = atomicAdd(p.counts + key, (unsigned)__popc(peers));
        }
        pend_key[r] = key;
      }
    }
    if (tma) __syncthreads();  // every thread is done with stage s before it is refilled
  }
#pragma unroll
  for (int r = 0; r < BEV_ITEMS; ++r) {
    if (pend2_old[r] == 0) atomicOr(p.dirty + (pend2_key[r] >> 7), 1u << ((pend2_key[r] >> 2) & 31));
    if (pend_old[r] == 0) atomicOr(p.dirty + (pend_key[r] >> 7), 1u << ((pend_key[r] >> 2) & 31));
  }
}

struct BevOut {
  unsigned* dirty;     // see BevParams
  float* raw;
  float* norm;
  uint8_t* u8;
  const uint8_t* map;
  float* chw;
  float max_intensity;
  unsigned cells;
  int S0, S1, S2;
  int n_frames;        // frames in this sub-batch
  int64_t frame_base;  // first frame (output offset)
  int zeros_done;      // the histogram kernel has already streamed the zeros of the dense outputs (pass A)
  int u16;             // 16-bit counts, two per word (see BevParams)
};

// the four counts of quad q / their reset (kept apart: the finalize kernels issue all their loads before the first store)
__device__ __forceinline__ uint4 bev_load_quad(const unsigned* counts, int64_t q, int u16) {
  if (u16) {
    const uint2 w = reinterpret_cast<const uint2*>(counts)[q];
    return make_uint4(w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16);
  }
  return reinterpret_cast<const uint4*>(counts)[q];
}
__device__ __forceinline__ void bev_clear_quad(unsigned* counts, int64_t q, int u16) {
  if (u16) reinterpret_cast<uint2*>(counts)[q] = make_uint2(0u, 0u);
  else reinterpret_cast<uint4*>(counts)[q] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ void bev_cell(unsigned c, float max_intensity, float& raw, float& nrm, uint8_t& q) {
  raw = (float)c;                                                  // exact below 2^24
  nrm = fminf(fmaxf(__fdiv_rn(raw, max_intensity), 0.0f), 1.0f);   // (bev/max).clip(0,1), :103-104
  q = (uint8_t)rintf(__fmul_rn(nrm, 255.0f));                      // np.round(bev*255).astype(uint8), :213
}

// The grid is sparse (2% of the cells of a Lyft sweep are hit): the finalize kernels read the
// dirty bitmap (1 bit per quad of counts) instead of the counts, load - and clear - only the
// dirty quads, and stream the dense outputs.  A quad's bit is cleared by the one thread that
// owns the quad, so counts and bitmap are all-zero again when the kernel ends.
// flat path (cells % 4 == 0): the counts of the frames in flight are one flat array of quads.
// A warp task = 1024 consecutive quads = 32 words of the bitmap (lane l holds word l, one
// coalesced load).  Pass A streams zeros over the whole task - it depends on nothing, so the
// store stream never waits; pass B revisits only the dirty quads (2-7% on a Lyft sweep): loads
// and clears their counts, four independent loads at a time, and overwrites the outputs.
__global__ void __launch_bounds__(256) bev_finalize_flat4_kernel(unsigned* counts, BevOut o) {
  const int lane = threadIdx.x & 31;
  const int64_t total_quads = (int64_t)(o.cells / 4) * o.n_frames;
  const int64_t n_tasks = (total_quads + 1023) >> 10;
  const int64_t out0 = (int64_t)o.frame_base * (o.cells / 4);
  float4* raw = o.raw ? reinterpret_cast<float4*>(o.raw) + out0 : nullptr;
  float4* nrm = o.norm ? reinterpret_cast<float4*>(o.norm) + out0 : nullptr;
  uchar4* u8 = o.u8 ? reinterpret_cast<uchar4*>(o.u8) + out0 : nullptr;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const uchar4 zb = make_uchar4(0, 0, 0, 0);
  for (int64_t task = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); task < n_tasks; task += (int64_t)gridDim.x * 8) {
    const int64_t q0 = task << 10;
    const int64_t w = (q0 >> 5) + lane;
    unsigned word = 0;
    if ((w << 5) < total_quads) {
      word = o.dirty[w];
      if (word) o.dirty[w] = 0;   // the warp owns these 32 words
    }
    // pass A (unless bev_hist_kernel did it beside its atomics)
    if (!o.zeros_done) {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const int64_t q = q0 + k * 32 + lane;
        if (q < total_quads) {
          if (raw) lv_st_stream_f4(raw + q, z4);
          if (nrm) lv_st_stream_f4(nrm + q, z4);
          if (u8) u8[q] = zb;
        }
      }
    }
    if (!__any_sync(0xffffffffu, word != 0)) continue;
    // pass B: lane l revisits the dirty quads of its column (quads q0 + 32 k + l, bit l of
    // word k): same thread as in pass A, coalesced with its neighbours, four loads in flight
    unsigned col = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) col |= ((__shfl_sync(0xffffffffu, word, k) >> lane) & 1u) << k;
    while (__any_sync(0xffffffffu, col != 0)) {
      int64_t q[4];
      uint4 c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[j] = -1;
        if (col) {
          q[j] = q0 + (__ffs(col) - 1) * 32 + lane;
          col &= col - 1;
          c[j] = bev_load_quad(counts, q[j], o.u16);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (q[j] < 0) continue;
        bev_clear_quad(counts, q[j], o.u16);
        float4 r, n;
        uchar4 b;
        bev_cell(c[j].x, o.max_intensity, r.x, n.x, b.x);
        bev_cell(c[j].y, o.max_intensity, r.y, n.y, b.y);
        bev_cell(c[j].z, o.max_intensity, r.z, n.z, b.z);
        bev_cell(c[j].w, o.max_intensity, r.w, n.w, b.w);
        if (raw) raw[q[j]] = r;      // same thread, same address as its pass-A store: ordered
        if (nrm) nrm[q[j]] = n;
        if (u8) u8[q[j]] = b;
      }
    }
  }
}

// any cell count (cells % 4 != 0): one thread per cell; the bitmap is cleared by a memset after
__global__ void __launch_bounds__(256) bev_finalize_scalar_kernel(unsigned* counts, BevOut o) {
  const int64_t total = (int64_t)o.cells * o.n_frames;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    unsigned c = 0;
    if (o.dirty[q >> 7] & (1u << ((q >> 2) & 31))) {
      if (o.u16) {
        unsigned short* c16 = reinterpret_cast<unsigned short*>(counts);
        c = c16[q];
        c16[q] = 0;
      } else {
        c = counts[q];
        counts[q] = 0;
      }
    }
    const int64_t out_q = o.frame_base * (int64_t)o.cells + q;
    float r, n;
    uint8_t b;
    bev_cell(c, o.max_intensity, r, n, b);
    if (o.raw) o.raw[out_q] = r;
    if (o.norm) o.norm[out_q] = n;
    if (o.u8) o.u8[out_q] = b;
  }
}

// HWC(3) -> CHW(6) path: one thread = 4 consecutive x of one row (S2 == 3, S1 % 4 == 0).
__global__ void __launch_bounds__(256) bev_finalize_hwc3_kernel(unsigned* counts, BevOut o) {
  const int64_t groups_per_frame = (int64_t)o.S0 * (o.S1 / 4);
  const int64_t total = groups_per_frame * o.n_frames;
  const int64_t plane = (int64_t)o.S0 * o.S1;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = g / groups_per_frame;
    const int64_t gi = g - f * groups_per_frame;        // (y, x/4) flattened: y*(S1/4) + x4
    unsigned c[12];                                        // 12 consecutive counts = 3 quads
    bool dirty[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // the three bitmap loads, then the three count loads, are independent
      const unsigned gq = (unsigned)(g * 3 + k);
      dirty[k] = (o.dirty[gq >> 5] >> (gq & 31)) & 1u;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      *reinterpret_cast<uint4*>(c + 4 * k) = dirty[k] ? bev_load_quad(counts, g * 3 + k, o.u16) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (!dirty[k]) continue;
      const unsigned gq = (unsigned)(g * 3 + k);
      bev_clear_quad(counts, g * 3 + k, o.u16);
      atomicAnd(o.dirty + (gq >> 5), ~(1u << (gq & 31)));   // this thread owns the quad's bit
    }
    float r[12], n[12];
    uint8_t b[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) bev_cell(c[k], o.max_intensity, r[k], n[k], b[k]);
    const int64_t of = o.frame_base + f;
    const int64_t cell0 = of * (int64_t)o.cells + gi * 12;  // flat index of the first of the 12 cells
    if (o.raw) {
      float4* d = reinterpret_cast<float4*>(o.raw + cell0);
      lv_st_stream_f4(d + 0, make_float4(r[0], r[1], r[2], r[3]));
      lv_st_stream_f4(d + 1, make_float4(r[4], r[5], r[6], r[7]));
      lv_st_stream_f4(d + 2, make_float4(r[8], r[9], r[10], r[11]));
    }
    if (o.norm) {
      float4* d = reinterpret_cast<float4*>(o.norm + cell0);
      lv_st_stream_f4(d + 0, make_float4(n[0], n[1], n[2], n[3]));
      lv_st_stream_f4(d + 1, make_float4(n[4], n[5], n[6], n[7]));
      lv_st_stream_f4(d + 2, make_float4(n[8], n[9], n[10], n[11]));
    }
    if (o.u8) {
      uchar4* d = reinterpret_cast<uchar4*>(o.u8 + cell0);
      d[0] = make_uchar4(b[0], b[1], b[2], b[3]);
      d[1] = make_uchar4(b[4], b[5], b[6], b[7]);
      d[2] = make_uchar4(b[8], b[9], b[10], b[11]);
    }
    if (o.chw) {
      // dataset.py:88,104,106: concat(im, map) -> float32 / 255 -> CHW
      const uchar4* mp = reinterpret_cast<const uchar4*>(o.map + cell0);
      const uchar4 m0 = mp[0], m1 = mp[1], m2 = mp[2];
      const uint8_t mb[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
      float* dst = o.chw + of * 6 * plane + gi * 4;   // (y*S1 + x) = gi*4
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        lv_st_stream_f4(reinterpret_cast<float4*>(dst + ch * plane),
                        make_float4(__fdiv_rn((float)b[ch], 255.0f), __fdiv_rn((float)b[3 + ch], 255.0f),
                                    __fdiv_rn((float)b[6 + ch], 255.0f), __fdiv_rn((float)b[9 + ch], 255.0f)));
        lv_st_stream_f4(reinterpret_cast<float4*>(dst + (3 + ch) * plane),
                        make_float4(__fdiv_rn((float)mb[ch], 255.0f), __fdiv_rn((float)mb[3 + ch], 255.0f),
                                    __fdiv_rn((float)mb[6 + ch], 255.0f), __fdiv_rn((float)mb[9 + ch], 255.0f)));
      }
    }
  }
}

__global__ void __launch_bounds__(256) bev_normalize_kernel(const float* in, int64_t n, float max_intensity,
                                                          float* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fminf(fmaxf(__fdiv_rn(in[i], max_intensity), 0.0f), 1.0f);
}

struct Tm12 { double m[12]; };
__global__ void __launch_bounds__(256) transform_points_kernel(const float* __restrict__ pts, int stride, int64_t n,
                                                             Tm12 tm, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* q = pts + i * stride;
    const double X = __ldg(q), Y = __ldg(q + 1), Z = __ldg(q + 2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double* M = tm.m + 4 * r;
      out[(int64_t)r * n + i] = fma(M[3], 1.0, fma(M[2], Z, fma(M[1], Y, M[0] * X)));
    }
  }
}

static int bev_grid(const lv_handle* h, int64_t work_items, int per_sm) {
  int64_t blocks = lv_div_up(work_items, 256);
  int64_t cap = (int64_t)h->num_sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int lv_bev_rasterize(lv_handle* h, const float* d_points, int32_t point_stride,
                                int32_t n_segments, const int64_t* h_seg_offsets,
                                const int32_t* h_seg_frame, const double* h_seg_tm, int32_t n_frames,
                                const int32_t shape[3], const double voxel_size[3], double z_offset,
                                float max_intensity, float* d_raw, float* d_norm, uint8_t* d_u8,
                                const uint8_t* d_map_u8, float* d_chw, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_bev_rasterize: null handle");
  LV_REQUIRE(shape && voxel_size && h_seg_offsets, "lv_bev_rasterize: null shape/voxel_size/seg_offsets");
  LV_REQUIRE(n_segments >= 0 && n_frames >= 0, "lv_bev_rasterize: negative counts");
  LV_REQUIRE(point_stride >= 3, "lv_bev_rasterize: point_stride must be >= 3, got %d", point_stride);
  LV_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "Voxel volume shape should be 3 positive dimensions (x,y,z)");
  LV_REQUIRE(shape[0] == shape[1],
             "lv_bev_rasterize: shape[0] must equal shape[1] (the reference indexes bev[y,x,z] into an array "
             "of shape `shape`, generating_train_bev.py:90-99), got %d x %d", shape[0], shape[1]);
  LV_REQUIRE(voxel_size[0] > 0 && voxel_size[1] > 0 && voxel_size[2] > 0, "lv_bev_rasterize: voxel_size must be > 0");
  LV_REQUIRE(max_intensity > 0.0f, "lv_bev_rasterize: max_intensity must be > 0");
  const int64_t cells64 = (int64_t)shape[0] * shape[1] * shape[2];
  LV_REQUIRE(cells64 < (1ll << 31), "lv_bev_rasterize: grid too large");
  LV_REQUIRE(!d_chw || (d_map_u8 && shape[2] == 3 && shape[1] % 4 == 0),
             "lv_bev_rasterize: CHW output needs a map raster, shape[2]==3 and shape[1]%%4==0");
  const int64_t n_total = n_segments > 0 ? h_seg_offsets[n_segments] : 0;
  LV_REQUIRE(n_segments == 0 || h_seg_offsets[0] == 0, "lv_bev_rasterize: seg_offsets[0] must be 0");
  for (int s = 0; s < n_segments; ++s) {
    LV_REQUIRE(h_seg_offsets[s + 1] >= h_seg_offsets[s], "lv_bev_rasterize: seg_offsets must be non-decreasing");
    const int f = h_seg_frame ? h_seg_frame[s] : s;
    LV_REQUIRE(f >= 0 && f < n_frames, "lv_bev_rasterize: segment %d maps to frame %d outside [0,%d)", s, f, n_frames);
    if (h_seg_frame && s > 0)
      LV_REQUIRE(h_seg_frame[s] >= h_seg_frame[s - 1], "lv_bev_rasterize: seg_frame must be non-decreasing");
  }
  LV_REQUIRE(n_total == 0 || d_points, "lv_bev_rasterize: null points");
  if (n_frames == 0) return LV_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));

  // frames in flight.  Measured (bench.py, 128 frames of 336x336x3): 23 frames (32 MB of counts,
  // L2-resident) 0.200 ms, 64 frames 0.167 ms, 128 frames (173 MB) 0.153 ms - fewer, larger
  // launches beat L2 residency of the counts, so the default budget is 192 MB, balanced.
  const unsigned cells = (unsigned)cells64;
  int64_t fif = h->bev_frames_in_flight > 0 ? h->bev_frames_in_flight : (192ll << 20) / (cells64 * 4);
  if (fif < 1) fif = 1;
  if (fif > n_frames) fif = n_frames;
  fif = lv_div_up(n_frames, lv_div_up(n_frames, fif));
  while (fif > 1 && fif * cells64 >= 0xffffffffll) --fif;
  // 16-bit counts when no frame can overflow one (a frame's count is at most its number of points)
  bool u16 = h->bev_u16 != 0 && cells % 4 == 0;
  {
    std::vector<int64_t> per_frame((size_t)n_frames, 0);
    for (int s = 0; s < n_segments && u16; ++s) {
      const int f = h_seg_frame ? h_seg_frame[s] : s;
      per_frame[f] += h_seg_offsets[s + 1] - h_seg_offsets[s];
      if (per_frame[f] > 65535) u16 = false;
    }
  }
  LV_CHECK(h->bev_counts.ensure((size_t)fif * cells * sizeof(unsigned), stream, 0));
  const size_t dirty_bytes = (size_t)(lv_div_up(fif * cells64, 128) + 32) * sizeof(unsigned);
  LV_CHECK(h->bev_dirty.ensure(dirty_bytes, stream, 0));

  const void *d_off = nullptr, *d_frame = nullptr, *d_tm = nullptr;
  if (n_segments > 0) {
    LV_CHECK(h->bev_seg_offsets.sync(h_seg_offsets, sizeof(int64_t) * (n_segments + 1), stream, &d_off));
    if (h_seg_frame) LV_CHECK(h->bev_seg_frame.sync(h_seg_frame, sizeof(int32_t) * n_segments, stream, &d_frame));
    if (h_seg_tm) LV_CHECK(h->bev_seg_tm.sync(h_seg_tm, sizeof(double) * 16 * n_segments, stream, &d_tm));
  }

  BevParams p;
  p.pts = d_points;
  p.stride = point_stride;
  p.seg_offsets = (const int64_t*)d_off;
  p.seg_frame = (const int32_t*)d_frame;
  p.seg_tm = (const double*)d_tm;
  // create_transformation_matrix_to_voxel_space (:47-62): float64 throughout.
  // 1/voxel_size first, then shape/2 + offset/voxel_size (a true division).
  p.m0 = 1.0 / voxel_size[0];
  p.m1 = 1.0 / voxel_size[1];
  p.m2 = 1.0 / voxel_size[2];
  p.t0 = (double)shape[0] / 2 + 0.0 / voxel_size[0];
  p.t1 = (double)shape[1] / 2 + 0.0 / voxel_size[1];
  p.t2 = (double)shape[2] / 2 + z_offset / voxel_size[2];
  p.S0 = shape[0]; p.S1 = shape[1]; p.S2 = shape[2];
  p.cells = cells;
  p.counts = h->bev_counts.as<unsigned>();
  p.u16 = u16 ? 1 : 0;
  p.dirty = h->bev_dirty.as<unsigned>();
  // TMA bulk copies need 16-byte aligned tiles: tile t starts at byte t * 1024 * stride * 4
  p.use_tma = (point_stride == 4 || point_stride == 5) && (reinterpret_cast<uintptr_t>(d_points) & 15) == 0 &&
              h->bev_tma && !h->disable_tma;
  if (point_stride == 4)
    LV_CHECK_CUDA(cudaFuncSetAttribute(bev_hist_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BEV_STAGES * BEV_TILE * 16));
  if (point_stride == 5)
    LV_CHECK_CUDA(cudaFuncSetAttribute(bev_hist_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, BEV_STAGES * BEV_TILE * 20));

  BevOut o;
  o.dirty = p.dirty;
  o.raw = d_raw; o.norm = d_norm; o.u8 = d_u8; o.map = d_map_u8; o.chw = d_chw;
  o.max_intensity = max_intensity;
  o.cells = cells; o.S0 = shape[0]; o.S1 = shape[1]; o.S2 = shape[2];
  o.u16 = p.u16;

  int seg = 0;
  for (int64_t f0 = 0; f0 < n_frames; f0 += fif) {
    const int64_t f1 = (f0 + fif < n_frames) ? f0 + fif : n_frames;
    // segments of frames [f0, f1)
    const int s0 = seg;
    while (seg < n_segments && (h_seg_frame ? h_seg_frame[seg] : seg) < f1) ++seg;
    const int s1 = seg;
    p.seg_lo = s0; p.seg_hi = s1;
    p.frame_base = (int)f0;
    o.zeros_done = 0;
    p.z_quads = 0;
    p.z_per_tile = 0;
    p.z_raw = nullptr; p.z_norm = nullptr; p.z_u8 = nullptr;
    if (!d_chw && cells % 4 == 0 && h->bev_fused_zero && s1 > s0 && h_seg_offsets[s1] > h_seg_offsets[s0] &&
        ((reinterpret_cast<uintptr_t>(d_raw) | reinterpret_cast<uintptr_t>(d_norm)) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(d_u8) & 3) == 0) {
      const int64_t q0 = f0 * (int64_t)(cells / 4);
      p.z_quads = (f1 - f0) * (int64_t)(cells / 4);
      p.z_raw = d_raw ? reinterpret_cast<float4*>(d_raw) + q0 : nullptr;
      p.z_norm = d_norm ? reinterpret_cast<float4*>(d_norm) + q0 : nullptr;
      p.z_u8 = d_u8 ? reinterpret_cast<uchar4*>(d_u8) + q0 : nullptr;
      const int64_t n_tiles = lv_div_up(h_seg_offsets[s1], BEV_TILE) - h_seg_offsets[s0] / BEV_TILE;
      p.z_per_tile = (int)lv_div_up(p.z_quads, n_tiles);
      o.zeros_done = 1;
    }
    if (s1 > s0) {
      p.pt_begin = h_seg_offsets[s0];
      p.pt_end = h_seg_offsets[s1];
      const int64_t npts = p.pt_end - p.pt_begin;
      if (npts > 0) {
        const int64_t n_tiles = lv_div_up(p.pt_end, BEV_TILE) - p.pt_begin / BEV_TILE;
        if (point_stride == 4 && (reinterpret_cast<uintptr_t>(d_points) & 15) == 0) {
          const size_t smem = p.use_tma ? BEV_STAGES * BEV_TILE * 16 : 0;   // 64 KB: 3 CTAs per SM
          const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)h->num_sms * (p.use_tma ? 3 : 8));
          bev_hist_kernel<4><<<grid, BEV_THREADS, smem, stream>>>(p);
        } else if (point_stride == 5 && p.use_tma) {
          const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)h->num_sms * 2);   // 80 KB: 2 CTAs per SM
          bev_hist_kernel<5><<<grid, BEV_THREADS, BEV_STAGES * BEV_TILE * 20, stream>>>(p);
        } else {
          const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)h->num_sms * 8);
          bev_hist_kernel<0><<<grid, BEV_THREADS, 0, stream>>>(p);
        }
        LV_LAUNCH_CHECK(h);
      }
    }
    o.n_frames = (int)(f1 - f0);
    o.frame_base = f0;
    if (d_chw) {
      const int64_t items = (int64_t)o.n_frames * shape[0] * (shape[1] / 4);
      bev_finalize_hwc3_kernel<<<bev_grid(h, items, 8), 256, 0, stream>>>(p.counts, o);
    } else if (cells % 4 == 0) {
      const int64_t tasks = lv_div_up((int64_t)(cells / 4) * o.n_frames, 1024);
      bev_finalize_flat4_kernel<<<bev_grid(h, tasks * 32, 8), 256, 0, stream>>>(p.counts, o);
    } else {
      const int64_t items = (int64_t)o.n_frames * cells;
      bev_finalize_scalar_kernel<<<bev_grid(h, items, 8), 256, 0, stream>>>(p.counts, o);
      LV_CHECK_CUDA(cudaMemsetAsync(p.dirty, 0, dirty_bytes, stream));
    }
    LV_LAUNCH_CHECK(h);
  }
  return LV_OK;
}

extern "C" int lv_transform_points(lv_handle* h, const float* d_points, int32_t point_stride, int64_t n,
                                   const double* h_tm16, double* d_out, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_transform_points: null handle");
  LV_REQUIRE(n >= 0 && point_stride >= 3 && h_tm16, "lv_transform_points: bad arguments");
  if (n == 0) return LV_OK;
  LV_REQUIRE(d_points && d_out, "lv_transform_points: null pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  Tm12 tm;
  for (int i = 0; i < 12; ++i) tm.m[i] = h_tm16[i];
  transform_points_kernel<<<bev_grid(h, n, 8), 256, 0, (cudaStream_t)stream_>>>(d_points, point_stride, n, tm, d_out);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_bev_normalize(lv_handle* h, const float* d_in, int64_t n, float max_intensity, float* d_out,
                                lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_bev_normalize: null handle");
  LV_REQUIRE(n >= 0 && (n == 0 || (d_in && d_out)), "lv_bev_normalize: bad arguments");
  LV_REQUIRE(max_intensity != 0.0f, "lv_bev_normalize: max_intensity must be non-zero");
  if (n == 0) return LV_OK;
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  bev_normalize_kernel<<<bev_grid(h, n, 8), 256, 0, (cudaStream_t)stream_>>>(d_in, n, max_intensity, d_out);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_bev_rasterize_host(lv_handle* h, const float* h_points, int32_t point_stride,
                                     int32_t n_segments, const int64_t* h_seg_offsets,
                                     const int32_t* h_seg_frame, const double* h_seg_tm, int32_t n_frames,
                                     const int32_t shape[3], const double voxel_size[3], double z_offset,
                                     float max_intensity, float* h_raw, float* h_norm, uint8_t* h_u8,
                                     const uint8_t* h_map_u8, float* h_chw) {
  LV_REQUIRE(h != nullptr, "lv_bev_rasterize_host: null handle");
  LV_REQUIRE(shape && voxel_size && h_seg_offsets, "lv_bev_rasterize_host: null shape/voxel_size/seg_offsets");
  LV_REQUIRE(n_segments >= 0 && n_frames >= 0 && point_stride >= 3, "lv_bev_rasterize_host: bad counts");
  LV_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "Voxel volume shape should be 3 positive dimensions (x,y,z)");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const int64_t n_total = n_segments > 0 ? h_seg_offsets[n_segments] : 0;
  const size_t cells = (size_t)shape[0] * shape[1] * shape[2];
  const size_t pt_bytes = (size_t)n_total * point_stride * sizeof(float);
  LV_CHECK(h->bev_stage_points.ensure(pt_bytes, st));
  if (pt_bytes) LV_CHECK_CUDA(cudaMemcpyAsync(h->bev_stage_points.ptr, h_points, pt_bytes, cudaMemcpyHostToDevice, st));
  float *d_raw = nullptr, *d_norm = nullptr, *d_chw = nullptr;
  uint8_t *d_u8 = nullptr, *d_map = nullptr;
  const size_t fr = (size_t)n_frames;
  if (h_raw) { LV_CHECK(h->bev_stage_out[0].ensure(fr * cells * 4, st)); d_raw = h->bev_stage_out[0].as<float>(); }
  if (h_norm) { LV_CHECK(h->bev_stage_out[1].ensure(fr * cells * 4, st)); d_norm = h->bev_stage_out[1].as<float>(); }
  if (h_u8) { LV_CHECK(h->bev_stage_out[2].ensure(fr * cells, st)); d_u8 = h->bev_stage_out[2].as<uint8_t>(); }
  if (h_chw) {
    LV_REQUIRE(h_map_u8 != nullptr, "lv_bev_rasterize_host: CHW output needs a map raster");
    const size_t plane = (size_t)shape[0] * shape[1];
    LV_CHECK(h->bev_stage_out[3].ensure(fr * plane * 6 * 4, st));
    d_chw = h->bev_stage_out[3].as<float>();
    LV_CHECK(h->bev_stage_map.ensure(fr * plane * 3, st));
    d_map = h->bev_stage_map.as<uint8_t>();
    LV_CHECK_CUDA(cudaMemcpyAsync(d_map, h_map_u8, fr * plane * 3, cudaMemcpyHostToDevice, st));
  }
  LV_CHECK(lv_bev_rasterize(h, h->bev_stage_points.as<float>(), point_stride, n_segments, h_seg_offsets, h_seg_frame,
                            h_seg_tm, n_frames, shape, voxel_size, z_offset, max_intensity, d_raw, d_norm, d_u8,
                            d_map, d_chw, st));
  if (h_raw) LV_CHECK_CUDA(cudaMemcpyAsync(h_raw, d_raw, fr * cells * 4, cudaMemcpyDeviceToHost, st));
  if (h_norm) LV_CHECK_CUDA(cudaMemcpyAsync(h_norm, d_norm, fr * cells * 4, cudaMemcpyDeviceToHost, st));
  if (h_u8) LV_CHECK_CUDA(cudaMemcpyAsync(h_u8, d_u8, fr * cells, cudaMemcpyDeviceToHost, st));
  if (h_chw)
    LV_CHECK_CUDA(cudaMemcpyAsync(h_chw, d_chw, fr * (size_t)shape[0] * shape[1] * 6 * 4, cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  return LV_OK;
}
