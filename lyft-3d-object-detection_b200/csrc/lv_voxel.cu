// Hard voxelizer (sm_100a): exact first-come semantics on a parallel machine.
//
// Replaces spconv.utils.VoxelGeneratorV2.generate as called from
// second/second/data/preprocess.py:299-317; algorithm = the in-tree sibling
// second/second/utils/simplevis.py:9-61.
//
// The reference is a serial loop: voxel ids are handed out in order of first
// appearance, slots inside a voxel in point order.  Atomics give arbitrary order,
// so order is RECONSTRUCTED (all per frame, many frames per launch):
//
//   K1 cells      cell(i) = ((cz*gy)+cy)*gx+cx from floorf((p-lo)/vs) (fp32 IEEE sub/div);
//                 first[cell] = atomicMin(point index)   (warp-aggregated)
//   K2 count      creator(i) := first[cell(i)] == i ; per-chunk creator counts
//   K3 scan       per-frame exclusive scan of the chunk counts -> voxel_num = min(total, V)
//   K4 assign     rank(i) = #creators before i = voxel id; coords[rank] = (z,y,x);
//                 first[cell] := ~rank; `break` cut-off = index of the creator of rank V
//   K5 keys       vid(i) = ~first[cell(i)]; kept(i) = vid < V and i < cut;
//                 digit histogram of pass 1
//   K6..K10       stable LSD radix sort of the kept points by voxel id (2 passes of
//                 <= 9 bits, 3 beyond 2^18 voxels): chunk histogram -> per-frame scan
//                 -> stable in-chunk rank (warp match + per-warp digit counters) + scatter.
//                 Stability makes the order inside a voxel the point order for free.
//   K11 heads     segment [start,end) of every voxel in the sorted list
//   K12 gather    one (sub)warp per voxel: the first min(count,T) points of its segment
//                 are copied into voxels[v, slot, :], the rest of the row is zero-filled,
//                 num_points[v] = min(count, T).  Every output byte is written once.
//   K13 reset     creators put EMPTY back into first[] (touched-cell reset) so the
//                 dense map never needs a memset.
//
// HBM layout: points (N,C) f32 rows; workspace per sub-batch of frames: first[] dense
// int32 map [frames_in_flight][gz*gy*gx] (all EMPTY between calls), cell[] int32,
// key/val ping-pong buffers, chunk histograms.  Outputs padded per frame:
// voxels (F,V,T,C) f32, coords (F,V,3) i32 zyx, num_points (F,V) i32, voxel_num (F) i32.
// Algorithmic bytes: 4*C per point read + V*(T*C*4 + 16) written.
#include <math.h>

#include "lv_common.cuh"
#include "lv_decorate.cuh"

#define VX_THREADS 256
#define VX_ITEMS 8
#define VX_CHUNK (VX_THREADS * VX_ITEMS)  // 2048 points per chunk; chunks never straddle frames
#define VX_EMPTY 0x7f7f7f7f               // memset-able "no point yet"
#define VX_CREATOR_BIT 0x40000000         // flag kept in cell[] (grid cells < 2^28)
#define VX_MAX_DIGIT_BITS 9
#define VX_MAX_DIGITS (1 << VX_MAX_DIGIT_BITS)
#define VX_DROPPED 0xffffffffu

struct VoxParams {
  const float* pts;            // (N_total, C)
  int C;
  const int64_t* frame_off;    // device [F+1] global point offsets
  const int32_t* frame_chunk;  // device [F+1] chunk prefix
  int f0, f1;                  // frames of this sub-batch
  int chunk_lo;                // first chunk of the sub-batch (global chunk id)
  int64_t pt_lo;               // first point of the sub-batch (global point index)
  float lo[3], vs[3];
  int grid[3];                 // gx, gy, gz
  int64_t G;                   // cells per frame
  int T, V, overflow, zero_tail;
  // workspace
  int32_t* map;                // [f1-f0][G]
  int32_t* cell;               // [points in sub-batch]
  uint32_t* key0;              // pass-0 keys (vid or VX_DROPPED), indexed like cell
  uint32_t* keyA; int32_t* valA;
  uint32_t* keyB; int32_t* valB;
  int32_t* chunk_cnt;          // [chunks in sub-batch] creator counts -> exclusive bases
  int32_t* hist;               // [chunks in sub-batch * digits]
  int32_t* frame_total;        // [frames in sub-batch] creators
  int32_t* frame_cut;          // [frames in sub-batch] break cut-off (local index)
  int32_t* frame_kept;         // [frames in sub-batch] kept points
  int32_t* seg_start;          // [frames in sub-batch][V]
  int32_t* seg_end;
  // outputs
  float* voxels; int32_t* coords; int32_t* num_points; int32_t* voxel_num;
  int64_t* row_base;           // [F+1] first output row of every frame (padded: f*V; concat: prefix of voxel_num)
  int concat;                  // 1: frames back to back, coords carry the batch index (coord_cols == 4)
  int coord_cols;              // 3 (z,y,x) or 4 (b,z,y,x)
  int64_t capacity;            // output rows available
};

struct ChunkLoc {
  int f;           // global frame id
  int fl;          // frame index inside the sub-batch
  int c;           // chunk index inside the frame
  int nchunks;     // chunks of the frame
  int64_t start;   // global point index of the frame start
  int n;           // points of the frame
};

// One thread resolves the frame of the CTA's chunk by binary search, then broadcasts.
__device__ __forceinline__ ChunkLoc vx_locate(const VoxParams& p, int* smem4) {
  if (threadIdx.x == 0) {
    const int g = p.chunk_lo + blockIdx.x;
    int lo = p.f0, hi = p.f1;  // frame_chunk[lo] <= g < frame_chunk[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(p.frame_chunk + mid) <= g) lo = mid; else hi = mid;
    }
    smem4[0] = lo;
  }
  __syncthreads();
  ChunkLoc L;
  L.f = smem4[0];
  L.fl = L.f - p.f0;
  const int cb = __ldg(p.frame_chunk + L.f);
  L.c = p.chunk_lo + blockIdx.x - cb;
  L.nchunks = __ldg(p.frame_chunk + L.f + 1) - cb;
  L.start = __ldg(p.frame_off + L.f);
  L.n = (int)(__ldg(p.frame_off + L.f + 1) - L.start);
  return L;
}

// ---------------------------------------------------------------- K1: cells + first index
template <bool C4>
__global__ void __launch_bounds__(VX_THREADS) vx_cells_kernel(VoxParams p) {
  __shared__ int sm[4];
  const ChunkLoc L = vx_locate(p, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int base = L.c * VX_CHUNK + warp * (32 * VX_ITEMS);
#pragma unroll 2
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int li = base + r * 32 + lane;  // local point index inside the frame
    int cell = -1;
    if (li < L.n) {
      const int64_t gi = L.start + li;
      float x, y, z;
      if (C4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.pts) + gi);
        x = v.x; y = v.y; z = v.z;
      } else {
        const float* q = p.pts + gi * p.C;
        x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2);
      }
      // simplevis.py:38-42: c = floor((p - lo) / vs) in float32, bounds on the float
      const float cx = floorf(__fdiv_rn(__fsub_rn(x, p.lo[0]), p.vs[0]));
      const float cy = floorf(__fdiv_rn(__fsub_rn(y, p.lo[1]), p.vs[1]));
      const float cz = floorf(__fdiv_rn(__fsub_rn(z, p.lo[2]), p.vs[2]));
      if (cx >= 0.f && cx < (float)p.grid[0] && cy >= 0.f && cy < (float)p.grid[1] && cz >= 0.f &&
          cz < (float)p.grid[2])
        cell = ((int)cz * p.grid[1] + (int)cy) * p.grid[0] + (int)cx;
      p.cell[gi - p.pt_lo] = cell;
    }
    // lanes of one cell: only the lowest lane (= lowest index) needs to bid
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    if (cell >= 0 && lane == __ffs(peers) - 1) atomicMin(map + cell, li);
  }
}

// ---------------------------------------------------------------- K2: creators per chunk
__device__ __forceinline__ int vx_block_sum(int v, int* sm_warp /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < VX_THREADS / 32; ++w) t += sm_warp[w];
  return t;
}

__global__ void __launch_bounds__(VX_THREADS) vx_count_kernel(VoxParams p) {
  __shared__ int sm[4];
  __shared__ int sw[VX_THREADS / 32];
  const ChunkLoc L = vx_locate(p, sm);
  const int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int base = L.c * VX_CHUNK;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k * VX_THREADS + threadIdx.x;
    if (li < L.n) {
      const int cell = p.cell[L.start + li - p.pt_lo];
      if (cell >= 0 && map[cell] == li) ++cnt;
    }
  }
  const int total = vx_block_sum(cnt, sw);
  if (threadIdx.x == 0) p.chunk_cnt[blockIdx.x] = total;
}

// ---------------------------------------------------------------- block exclusive scan helper
// in-place exclusive scan of `n` ints by one CTA (VX_THREADS threads); returns the total.
__device__ int vx_cta_exclusive_scan(int32_t* data, int n, int* sw /*[8]*/, int* carry_sm) {
  if (threadIdx.x == 0) *carry_sm = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += VX_THREADS * 4) {
    int v[4];
    int s = 0;
    const int i0 = base + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (i0 + k < n) ? data[i0 + k] : 0;
      s += v[k];
    }
    int inc = s;  // inclusive warp scan of the thread sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    int woff = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < VX_THREADS / 32; ++w) {
      const int t = sw[w];
      if (w < warp) woff += t;
      tile_total += t;
    }
    int run = *carry_sm + woff + inc - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) data[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) *carry_sm += tile_total;
    __syncthreads();
  }
  return *carry_sm;
}

// ---------------------------------------------------------------- K3: per-frame scan of chunk counts
__global__ void __launch_bounds__(VX_THREADS) vx_scan_chunks_kernel(VoxParams p) {
  __shared__ int sw[VX_THREADS / 32];
  __shared__ int carry;
  const int fl = blockIdx.x, f = p.f0 + fl;
  const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  const int total = vx_cta_exclusive_scan(p.chunk_cnt + cb, nch, sw, &carry);
  if (threadIdx.x == 0) {
    p.frame_total[fl] = total;
    p.frame_cut[fl] = 0x7fffffff;
    p.voxel_num[f] = total < p.V ? total : p.V;
    if (!p.concat) {
      p.row_base[f] = (int64_t)f * p.V;
      if (f + 1 == p.f1) p.row_base[f + 1] = (int64_t)(f + 1) * p.V;
    }
  }
}

// concat layout (merge_second_batch, second/second/data/preprocess.py:28-50): row_base = prefix of voxel_num
__global__ void vx_row_base_kernel(VoxParams p) {
  const int lane = threadIdx.x;
  int64_t carry = p.row_base[p.f0];
  for (int b = p.f0; b < p.f1; b += 32) {
    const int f = b + lane;
    int v = f < p.f1 ? p.voxel_num[f] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (f < p.f1) p.row_base[f + 1] = carry + inc;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// ---------------------------------------------------------------- K4: voxel ids for creators
__global__ void __launch_bounds__(VX_THREADS) vx_assign_kernel(VoxParams p) {
  __shared__ int sm[4];
  __shared__ int sw[VX_THREADS / 32];
  const ChunkLoc L = vx_locate(p, sm);
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread-contiguous items: thread t owns local points base + t*8 .. +7 (index order)
  const int base = L.c * VX_CHUNK + threadIdx.x * VX_ITEMS;
  int cell[VX_ITEMS];
  unsigned flags = 0;
  int s = 0;
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k;
    cell[k] = -1;
    if (li < L.n) {
      cell[k] = p.cell[L.start + li - p.pt_lo];
      if (cell[k] >= 0 && map[cell[k]] == li) {
        flags |= 1u << k;
        ++s;
      }
    }
  }
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sw[warp] = inc;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int w = 0; w < VX_THREADS / 32; ++w)
    if (w < warp) woff += sw[w];
  int rank = p.chunk_cnt[blockIdx.x] + woff + inc - s;
  if (flags == 0) return;
  const int gx = p.grid[0], gy = p.grid[1];
  const int64_t row0 = p.row_base[L.f];
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    if (!(flags & (1u << k))) continue;
    const int li = base + k;
    const int c = cell[k];
    map[c] = ~rank;  // voxel id, stored negative so it never equals a point index
    p.cell[L.start + li - p.pt_lo] = c | VX_CREATOR_BIT;
    if (rank < p.V) {
      const int cx = c % gx, cy = (c / gx) % gy, cz = c / (gx * gy);
      const int64_t row = row0 + rank;
      if (row < p.capacity) {
        if (p.coord_cols == 4) {  // batch index first (preprocess.py:44-50)
          *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(L.f, cz, cy, cx);
        } else {
          int32_t* co = p.coords + row * 3;
          co[0] = cz; co[1] = cy; co[2] = cx;  // reversed (z,y,x), simplevis.py:42
        }
      }
    } else if (rank == p.V && p.overflow == LV_OVERFLOW_BREAK) {
      p.frame_cut[L.fl] = li;  // simplevis.py:48-49: the loop stops here
    }
    ++rank;
  }
}

// ---------------------------------------------------------------- K5: keys + pass-1 histogram
struct SortPass {
  int shift, bits;   // digit = (key >> shift) & ((1<<bits)-1)
  int pass;          // 0: input key0 (indexed by point), 1..: input compacted key/val
};

__global__ void __launch_bounds__(VX_THREADS) vx_keys_kernel(VoxParams p, SortPass sp) {
  __shared__ int sm[4];
  __shared__ int hist[VX_MAX_DIGITS];
  const ChunkLoc L = vx_locate(p, sm);
  const int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int D = 1 << sp.bits;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) hist[d] = 0;
  __syncthreads();
  const int cut = p.frame_cut[L.fl];
  const int base = L.c * VX_CHUNK;
  const int lane = threadIdx.x & 31;
#pragma unroll 2
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k * VX_THREADS + threadIdx.x;
    unsigned key = VX_DROPPED;
    if (li < L.n) {
      const int64_t wi = L.start + li - p.pt_lo;
      const int cell = p.cell[wi] & ~VX_CREATOR_BIT;
      if (p.cell[wi] >= 0) {
        const int vid = ~map[cell];
        if (vid < p.V && li < cut) key = (unsigned)vid;
      }
      p.key0[wi] = key;
    }
    const unsigned digit = key == VX_DROPPED ? 0xffffffffu : ((key >> sp.shift) & (D - 1));
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (key != VX_DROPPED && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
  }
  __syncthreads();
  // table layout per frame: [digit][chunk]
  int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) tab[(int64_t)d * L.nchunks + L.c] = hist[d];
}

// histogram of a later pass over the compacted (key,val) list of the frame
__global__ void __launch_bounds__(VX_THREADS) vx_hist_kernel(VoxParams p, SortPass sp, const uint32_t* __restrict__ keys) {
  __shared__ int sm[4];
  __shared__ int hist[VX_MAX_DIGITS];
  const ChunkLoc L = vx_locate(p, sm);
  const int D = 1 << sp.bits;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) hist[d] = 0;
  __syncthreads();
  const int kept = p.frame_kept[L.fl];
  const int base = L.c * VX_CHUNK;
  const int lane = threadIdx.x & 31;
  if (base < kept) {
#pragma unroll 2
    for (int k = 0; k < VX_ITEMS; ++k) {
      const int li = base + k * VX_THREADS + threadIdx.x;
      unsigned digit = 0xffffffffu;
      if (li < kept) digit = (keys[L.start + li - p.pt_lo] >> sp.shift) & (D - 1);
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (digit != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
    }
  }
  __syncthreads();
  int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) tab[(int64_t)d * L.nchunks + L.c] = hist[d];
}

// ---------------------------------------------------------------- K6/K9: per-frame scan of [digit][chunk]
__global__ void __launch_bounds__(VX_THREADS) vx_scan_hist_kernel(VoxParams p, SortPass sp) {
  __shared__ int sw[VX_THREADS / 32];
  __shared__ int carry;
  const int fl = blockIdx.x, f = p.f0 + fl;
  const int D = 1 << sp.bits;
  const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  const int total = vx_cta_exclusive_scan(p.hist + (int64_t)cb * D, nch * D, sw, &carry);
  if (threadIdx.x == 0) p.frame_kept[fl] = total;
}

// ---------------------------------------------------------------- K7/K10: stable rank + scatter
// Warp w owns items [w*256, (w+1)*256) of the chunk in 8 rounds of 32 lanes, so
// (warp, round, lane) order is point order.  Per-warp digit counters live in shared
// memory; after the rounds an exclusive scan over the warps (plus the scanned global
// table) turns them into the warp's base for each digit.
__global__ void __launch_bounds__(VX_THREADS) vx_scatter_kernel(VoxParams p, SortPass sp,
                                                               const uint32_t* __restrict__ keys_in,
                                                               const int32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out,
                                                               int32_t* __restrict__ vals_out) {
  __shared__ int sm[4];
  extern __shared__ int cnt[];  // [8][D]
  const ChunkLoc L = vx_locate(p, sm);
  const int D = 1 << sp.bits;
  const int limit = sp.pass == 0 ? L.n : p.frame_kept[L.fl];
  const int chunk_base = L.c * VX_CHUNK;
  if (chunk_base >= limit) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (VX_THREADS / 32) * D; i += VX_THREADS) cnt[i] = 0;
  __syncthreads();
  int* mycnt = cnt + warp * D;
  unsigned key[VX_ITEMS];
  int val[VX_ITEMS], rnk[VX_ITEMS];
  const int base = chunk_base + warp * (32 * VX_ITEMS);
  const unsigned lt = lv_lanemask_lt();
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int li = base + r * 32 + lane;
    key[r] = VX_DROPPED;
    val[r] = 0;
    if (li < limit) {
      const int64_t wi = L.start + li - p.pt_lo;
      key[r] = keys_in[wi];
      val[r] = sp.pass == 0 ? li : vals_in[wi];
    }
    const bool live = key[r] != VX_DROPPED;
    const unsigned digit = live ? ((key[r] >> sp.shift) & (D - 1)) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    int old = 0;
    const int leader = __ffs(peers) - 1;
    if (live && lane == leader) {
      old = mycnt[digit];
      mycnt[digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  // warp bases: global table entry (already exclusive-scanned) + counts of the lower warps
  const int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) {
    int off = tab[(int64_t)d * L.nchunks + L.c];
#pragma unroll
    for (int w = 0; w < VX_THREADS / 32; ++w) {
      const int t = cnt[w * D + d];
      cnt[w * D + d] = off;
      off += t;
    }
  }
  __syncthreads();
  const int64_t out0 = L.start - p.pt_lo;  // the frame's sorted list starts at its first point slot
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    if (key[r] == VX_DROPPED) continue;
    const unsigned digit = (key[r] >> sp.shift) & (D - 1);
    const int64_t pos = out0 + mycnt[digit] + rnk[r];
    keys_out[pos] = key[r];
    vals_out[pos] = val[r];
  }
}

// ---------------------------------------------------------------- K11: voxel segments
__global__ void __launch_bounds__(VX_THREADS) vx_heads_kernel(VoxParams p, const uint32_t* __restrict__ keys) {
  __shared__ int sm[4];
  const ChunkLoc L = vx_locate(p, sm);
  const int kept = p.frame_kept[L.fl];
  const int base = L.c * VX_CHUNK;
  if (base >= kept) return;
  const uint32_t* k = keys + (L.start - p.pt_lo);
  int32_t* ss = p.seg_start + (int64_t)L.fl * p.V;
  int32_t* se = p.seg_end + (int64_t)L.fl * p.V;
#pragma unroll 2
  for (int j = 0; j < VX_ITEMS; ++j) {
    const int li = base + j * VX_THREADS + threadIdx.x;
    if (li >= kept) continue;
    const unsigned me = k[li];
    if (li == 0 || k[li - 1] != me) ss[me] = li;
    if (li == kept - 1 || k[li + 1] != me) se[me] = li + 1;
  }
}

// ---------------------------------------------------------------- K12: gather into voxels
// LPV lanes cooperate on one voxel (32 for pillars, 8 for T=5).  grid = (GX, frames); the
// voxel groups of a CTA loop over the frame's voxels, so no CTA is launched for rows that
// do not exist (voxel_num is only known on the device).
template <int LPV, bool C4>
__global__ void __launch_bounds__(VX_THREADS) vx_gather_kernel(VoxParams p, const int32_t* __restrict__ vals) {
  const int fl = blockIdx.y, f = p.f0 + fl;
  const int sub = threadIdx.x % LPV;
  const int vnum = p.voxel_num[f];
  const int limit = (p.zero_tail && !p.concat) ? p.V : vnum;
  const int64_t row0 = p.row_base[f];
  const int64_t fstart = __ldg(p.frame_off + f);
  const int groups = VX_THREADS / LPV;
  for (int v = blockIdx.x * groups + threadIdx.x / LPV; v < limit; v += gridDim.x * groups) {
    const int64_t row = row0 + v;
    if (row >= p.capacity) break;
    float* out = p.voxels + row * p.T * p.C;
    int n = 0;
    int s = 0;
    if (v < vnum) {
      s = p.seg_start[(int64_t)fl * p.V + v];
      n = p.seg_end[(int64_t)fl * p.V + v] - s;
      if (n > p.T) n = p.T;
    }
    if (sub == 0) {
      p.num_points[row] = n;
      if (v >= vnum) {
        int32_t* co = p.coords + row * p.coord_cols;
        for (int j = 0; j < p.coord_cols; ++j) co[j] = 0;
      }
    }
    const int32_t* vv = vals + (fstart - p.pt_lo) + s;
    if (C4) {
      float4* o4 = reinterpret_cast<float4*>(out);
      for (int t = sub; t < p.T; t += LPV) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < n) q = __ldg(reinterpret_cast<const float4*>(p.pts) + fstart + vv[t]);
        lv_st_stream_f4(o4 + t, q);
      }
    } else {
      for (int t = sub; t < p.T; t += LPV) {
        const float* src = (t < n) ? p.pts + (fstart + vv[t]) * p.C : nullptr;
        for (int c = 0; c < p.C; ++c) out[t * p.C + c] = src ? __ldg(src + c) : 0.f;
      }
    }
  }
}

// K12': gather fused with the PillarFeatureNet decoration (pointpillars.py:203-231): the
// (P,T,4) voxel tensor is never written; one warp per pillar pulls its <= T points through
// the sorted index list straight into registers and emits the decorated (T,C_out) block.
__global__ void __launch_bounds__(VX_THREADS, 6) vx_gather_decorate_kernel(VoxParams p, const int32_t* __restrict__ vals,
                                                                          DecoCfg d, float* __restrict__ decorated) {
  extern __shared__ float stage[];  // [8 warps][T*C_out]
  const int fl = blockIdx.y, f = p.f0 + fl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vnum = p.voxel_num[f];
  const int64_t row0 = p.row_base[f];
  const int64_t fstart = __ldg(p.frame_off + f);
  const int per = d.T * d.C_out;
  float* st = stage + warp * per;
  const float4* pts4 = reinterpret_cast<const float4*>(p.pts) + fstart;
  for (int v = blockIdx.x * (VX_THREADS / 32) + warp; v < vnum; v += gridDim.x * (VX_THREADS / 32)) {
    const int64_t row = row0 + v;
    if (row >= p.capacity) break;
    const int s = p.seg_start[(int64_t)fl * p.V + v];
    int n = p.seg_end[(int64_t)fl * p.V + v] - s;
    if (n > d.T) n = d.T;
    const int32_t* vv = vals + (fstart - p.pt_lo) + s;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (lane < n) a = __ldg(pts4 + vv[lane]);
    if (lane + 32 < n) b = __ldg(pts4 + vv[lane + 32]);
    const int4 co = *reinterpret_cast<const int4*>(p.coords + row * 4);  // b, z, y, x (written by K4)
    if (lane == 0) p.num_points[row] = n;
    lv_decorate_warp(a, b, n, co.z, co.w, d, st, decorated + row * per, lane);
  }
}

// ---------------------------------------------------------------- K13: touched-cell reset
__global__ void __launch_bounds__(VX_THREADS) vx_reset_kernel(VoxParams p) {
  __shared__ int sm[4];
  const ChunkLoc L = vx_locate(p, sm);
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int base = L.c * VX_CHUNK;
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k * VX_THREADS + threadIdx.x;
    if (li < L.n) {
      const int c = p.cell[L.start + li - p.pt_lo];
      if (c >= 0 && (c & VX_CREATOR_BIT)) map[c & ~VX_CREATOR_BIT] = VX_EMPTY;
    }
  }
}

// ================================================================= host side
extern "C" int lv_voxel_grid_size(const lv_voxel_config* cfg, int32_t grid_xyz[3]) {
  LV_REQUIRE(cfg && grid_xyz, "lv_voxel_grid_size: null argument");
  for (int j = 0; j < 3; ++j) {
    // simplevis.py:27-30: np.round((hi - lo) / vs) in float32, half to even
    volatile float span = cfg->coors_range[3 + j] - cfg->coors_range[j];
    volatile float g = span / cfg->voxel_size[j];
    grid_xyz[j] = (int32_t)rintf(g);
  }
  return LV_OK;
}

static int vx_bits_for(int v) {  // bits needed for ids in [0, v)
  int b = 1;
  while ((1ll << b) < v) ++b;
  return b;
}

static int vx_run(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                  const int64_t* h_frame_offsets, float* d_voxels, int32_t* d_coords, int32_t* d_num_points,
                  int32_t* d_voxel_num, int concat, int64_t capacity, int64_t* d_row_base, const DecoCfg* deco,
                  float* d_decorated, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_voxelize: null handle");
  LV_REQUIRE(cfg && h_frame_offsets, "lv_voxelize: null config / frame offsets");
  LV_REQUIRE(n_frames >= 0, "lv_voxelize: negative frame count");
  LV_REQUIRE(cfg->num_features >= 3, "lv_voxelize: points need >= 3 features, got %d", cfg->num_features);
  LV_REQUIRE(cfg->max_points > 0 && cfg->max_voxels > 0, "lv_voxelize: max_points and max_voxels must be > 0");
  LV_REQUIRE(cfg->overflow_mode == LV_OVERFLOW_CONTINUE || cfg->overflow_mode == LV_OVERFLOW_BREAK,
             "lv_voxelize: bad overflow_mode %d", cfg->overflow_mode);
  for (int j = 0; j < 3; ++j) LV_REQUIRE(cfg->voxel_size[j] > 0.f, "lv_voxelize: voxel_size must be > 0");
  int32_t grid[3];
  LV_CHECK(lv_voxel_grid_size(cfg, grid));
  LV_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0, "lv_voxelize: empty grid %d x %d x %d", grid[0], grid[1], grid[2]);
  const int64_t G = (int64_t)grid[0] * grid[1] * grid[2];
  LV_REQUIRE(G < (1ll << 28), "lv_voxelize: grid of %lld cells exceeds the dense-map limit (2^28)", (long long)G);
  LV_REQUIRE(n_frames == 0 || h_frame_offsets[0] == 0, "lv_voxelize: frame_offsets[0] must be 0");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE((d_voxels || deco) && d_coords && d_num_points && d_voxel_num, "lv_voxelize: null output");
  LV_REQUIRE(!deco || (d_decorated && concat && cfg->num_features == 4 && cfg->max_points <= 64 &&
                       (reinterpret_cast<uintptr_t>(d_points) & 15) == 0),
             "lv_pillarize_concat: needs 4 features per point, max_points <= 64 and 16-byte aligned points");
  const int V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features;

  // chunk prefix per frame (a frame with no points still owns zero chunks)
  std::vector<int32_t> frame_chunk(n_frames + 1, 0);
  for (int f = 0; f < n_frames; ++f) {
    const int64_t n = h_frame_offsets[f + 1] - h_frame_offsets[f];
    LV_REQUIRE(n >= 0, "lv_voxelize: frame_offsets must be non-decreasing");
    LV_REQUIRE(n < VX_EMPTY, "lv_voxelize: frame %d has too many points", f);
    const int64_t ch = lv_div_up(n, VX_CHUNK);
    LV_REQUIRE((int64_t)frame_chunk[f] + ch < (1ll << 31), "lv_voxelize: too many chunks");
    frame_chunk[f + 1] = frame_chunk[f] + (int32_t)ch;
  }
  const int64_t n_total = h_frame_offsets[n_frames];
  LV_REQUIRE(n_total == 0 || d_points, "lv_voxelize: null points");
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));

  const void *d_off = nullptr, *d_chunk = nullptr;
  LV_CHECK(h->vox_frame_offsets.sync(h_frame_offsets, sizeof(int64_t) * (n_frames + 1), stream, &d_off));
  LV_CHECK(h->vox_chunk_table.sync(frame_chunk.data(), sizeof(int32_t) * (n_frames + 1), stream, &d_chunk));

  // sub-batches: bounded by the dense map budget and by the point workspace
  const int64_t map_budget = h->vox_dense_map_limit_bytes > 0 ? h->vox_dense_map_limit_bytes : (64ll << 20);
  int64_t fif = map_budget / (G * 4);
  if (fif < 1) fif = 1;
  if (fif > n_frames) fif = n_frames;
  const int64_t max_pts = 8ll << 20;

  const int bits = vx_bits_for(V);
  const int npass = (bits + VX_MAX_DIGIT_BITS - 1) / VX_MAX_DIGIT_BITS;
  const int dbits = (bits + npass - 1) / npass;
  const int D = 1 << dbits;

  VoxParams p;
  memset(&p, 0, sizeof(p));
  p.pts = d_points; p.C = C;
  p.frame_off = (const int64_t*)d_off;
  p.frame_chunk = (const int32_t*)d_chunk;
  for (int j = 0; j < 3; ++j) {
    p.lo[j] = cfg->coors_range[j];
    p.vs[j] = cfg->voxel_size[j];
    p.grid[j] = grid[j];
  }
  p.G = G; p.T = T; p.V = V; p.overflow = cfg->overflow_mode; p.zero_tail = cfg->zero_tail;
  p.voxels = d_voxels; p.coords = d_coords; p.num_points = d_num_points; p.voxel_num = d_voxel_num;
  p.concat = concat;
  p.coord_cols = concat ? 4 : 3;
  p.capacity = concat ? capacity : (int64_t)n_frames * V;
  if (!d_row_base) {
    LV_CHECK(h->vox_row_base.ensure((size_t)(n_frames + 1) * sizeof(int64_t), stream));
    d_row_base = h->vox_row_base.as<int64_t>();
  }
  p.row_base = d_row_base;
  if (concat) {
    LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coords) & 15) == 0, "lv_voxelize_concat: coords must be 16-byte aligned");
    LV_CHECK_CUDA(cudaMemsetAsync(d_row_base, 0, sizeof(int64_t), stream));
  }
  const bool c4 = (C == 4) && (reinterpret_cast<uintptr_t>(d_points) & 15) == 0;
  const bool out4 = c4 && (reinterpret_cast<uintptr_t>(d_voxels) & 15) == 0;

  int f0 = 0;
  while (f0 < n_frames) {
    int f1 = f0 + 1;
    while (f1 < n_frames && f1 - f0 < fif && h_frame_offsets[f1 + 1] - h_frame_offsets[f0] <= max_pts) ++f1;
    const int nf = f1 - f0;
    const int64_t pt_lo = h_frame_offsets[f0], npts = h_frame_offsets[f1] - pt_lo;
    const int chunk_lo = frame_chunk[f0], nchunks = frame_chunk[f1] - chunk_lo;
    p.f0 = f0; p.f1 = f1; p.chunk_lo = chunk_lo; p.pt_lo = pt_lo;

    LV_CHECK(h->vox_map.ensure((size_t)nf * G * 4, stream, 0x7f));
    LV_CHECK(h->vox_cell.ensure((size_t)(npts + 1) * 4 * 2, stream));            // cell + key0
    LV_CHECK(h->vox_keys[0].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_keys[1].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_vals[0].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_vals[1].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_chunk.ensure((size_t)(nchunks + 1) * 4, stream));
    LV_CHECK(h->vox_hist.ensure((size_t)(nchunks + 1) * D * 4, stream));
    LV_CHECK(h->vox_frame_state.ensure((size_t)nf * (3 + 2 * (size_t)V) * 4, stream));
    p.map = h->vox_map.as<int32_t>();
    p.cell = h->vox_cell.as<int32_t>();
    p.key0 = h->vox_cell.as<uint32_t>() + (npts + 1);
    p.keyA = h->vox_keys[0].as<uint32_t>(); p.valA = h->vox_vals[0].as<int32_t>();
    p.keyB = h->vox_keys[1].as<uint32_t>(); p.valB = h->vox_vals[1].as<int32_t>();
    p.chunk_cnt = h->vox_chunk.as<int32_t>();
    p.hist = h->vox_hist.as<int32_t>();
    int32_t* fs = h->vox_frame_state.as<int32_t>();
    p.frame_total = fs; p.frame_cut = fs + nf; p.frame_kept = fs + 2 * nf;
    p.seg_start = fs + 3 * nf; p.seg_end = p.seg_start + (size_t)nf * V;

    if (nchunks > 0) {
      if (c4) vx_cells_kernel<true><<<nchunks, VX_THREADS, 0, stream>>>(p);
      else vx_cells_kernel<false><<<nchunks, VX_THREADS, 0, stream>>>(p);
      LV_LAUNCH_CHECK(h);
      vx_count_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p);
      LV_LAUNCH_CHECK(h);
    }
    vx_scan_chunks_kernel<<<nf, VX_THREADS, 0, stream>>>(p);
    LV_LAUNCH_CHECK(h);
    if (concat) {
      vx_row_base_kernel<<<1, 32, 0, stream>>>(p);
      LV_LAUNCH_CHECK(h);
    }
    const uint32_t* sorted_keys = p.keyA;
    const int32_t* sorted_vals = p.valA;
    if (nchunks > 0) {
      vx_assign_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p);
      LV_LAUNCH_CHECK(h);
      const size_t smem = (size_t)(VX_THREADS / 32) * D * sizeof(int);
      for (int pass = 0; pass < npass; ++pass) {
        SortPass sp{pass * dbits, dbits, pass};
        const uint32_t* kin = pass == 0 ? p.key0 : ((pass & 1) ? p.keyA : p.keyB);
        const int32_t* vin = (pass & 1) ? p.valA : p.valB;
        uint32_t* kout = (pass & 1) ? p.keyB : p.keyA;
        int32_t* vout = (pass & 1) ? p.valB : p.valA;
        if (pass == 0) vx_keys_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p, sp);
        else vx_hist_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p, sp, kin);
        LV_LAUNCH_CHECK(h);
        vx_scan_hist_kernel<<<nf, VX_THREADS, 0, stream>>>(p, sp);
        LV_LAUNCH_CHECK(h);
        vx_scatter_kernel<<<nchunks, VX_THREADS, smem, stream>>>(p, sp, kin, vin, kout, vout);
        LV_LAUNCH_CHECK(h);
        sorted_keys = kout;
        sorted_vals = vout;
      }
      vx_heads_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p, sorted_keys);
      LV_LAUNCH_CHECK(h);
    }
    if (deco) {
      const size_t smem = (size_t)(VX_THREADS / 32) * T * deco->C_out * sizeof(float);
      if (smem > 48 * 1024)
        LV_CHECK_CUDA(cudaFuncSetAttribute(vx_gather_decorate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int gx = (int)lv_div_up((int64_t)h->num_sms * 12, nf);
      if (gx < 8) gx = 8;
      dim3 grid_g((unsigned)gx, (unsigned)nf);
      vx_gather_decorate_kernel<<<grid_g, VX_THREADS, smem, stream>>>(p, sorted_vals, *deco, d_decorated);
      LV_LAUNCH_CHECK(h);
    } else {
      const int lpv = T >= 24 ? 32 : (T >= 12 ? 16 : 8);
      const int64_t full = lv_div_up((int64_t)V * lpv, VX_THREADS);
      int64_t gx = lv_div_up((int64_t)h->num_sms * 16, nf);
      if (gx < 8) gx = 8;
      if (gx > full) gx = full;
      dim3 grid_g((unsigned)gx, (unsigned)nf);
      if (out4) {
        if (lpv == 32) vx_gather_kernel<32, true><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
        else if (lpv == 16) vx_gather_kernel<16, true><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
        else vx_gather_kernel<8, true><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
      } else {
        if (lpv == 32) vx_gather_kernel<32, false><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
        else if (lpv == 16) vx_gather_kernel<16, false><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
        else vx_gather_kernel<8, false><<<grid_g, VX_THREADS, 0, stream>>>(p, sorted_vals);
      }
      LV_LAUNCH_CHECK(h);
    }
    if (nchunks > 0) {
      vx_reset_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p);
      LV_LAUNCH_CHECK(h);
    }
    f0 = f1;
  }
  return LV_OK;
}

extern "C" int lv_voxelize(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                           const int64_t* h_frame_offsets, float* d_voxels, int32_t* d_coords, int32_t* d_num_points,
                           int32_t* d_voxel_num, lv_stream stream) {
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, d_voxels, d_coords, d_num_points, d_voxel_num, 0, 0,
                nullptr, nullptr, nullptr, stream);
}

extern "C" int lv_voxelize_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                                  const int64_t* h_frame_offsets, int64_t capacity_rows, float* d_voxels,
                                  int32_t* d_coords4, int32_t* d_num_points, int32_t* d_voxel_num,
                                  int64_t* d_voxel_offsets, lv_stream stream) {
  LV_REQUIRE(capacity_rows >= 0, "lv_voxelize_concat: negative capacity");
  LV_REQUIRE(n_frames == 0 || d_voxel_offsets, "lv_voxelize_concat: null voxel_offsets");
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, d_voxels, d_coords4, d_num_points, d_voxel_num, 1,
                capacity_rows, d_voxel_offsets, nullptr, nullptr, stream);
}

extern "C" int lv_pillarize_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                                   const int64_t* h_frame_offsets, int64_t capacity_rows, float vx, float vy,
                                   float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                   float* d_decorated, int32_t* d_coords4, int32_t* d_num_points, int32_t* d_voxel_num,
                                   int64_t* d_voxel_offsets, lv_stream stream) {
  LV_REQUIRE(cfg != nullptr, "lv_pillarize_concat: null config");
  LV_REQUIRE(capacity_rows >= 0, "lv_pillarize_concat: negative capacity");
  LV_REQUIRE(n_frames == 0 || d_voxel_offsets, "lv_pillarize_concat: null voxel_offsets");
  const int c_out = lv_pillar_out_channels(cfg->num_features, variant, with_distance);
  LV_REQUIRE(c_out > 0, "lv_pillarize_concat: bad variant %d", variant);
  DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, cfg->max_points, c_out};
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, nullptr, d_coords4, d_num_points, d_voxel_num, 1,
                capacity_rows, d_voxel_offsets, &d, d_decorated, stream);
}

extern "C" int lv_voxelize_host(lv_handle* h, const lv_voxel_config* cfg, const float* h_points, int32_t n_frames,
                                const int64_t* h_frame_offsets, float* h_voxels, int32_t* h_coords,
                                int32_t* h_num_points, int32_t* h_voxel_num) {
  LV_REQUIRE(h != nullptr, "lv_voxelize_host: null handle");
  LV_REQUIRE(cfg && h_frame_offsets && n_frames >= 0, "lv_voxelize_host: bad arguments");
  LV_REQUIRE(cfg->num_features >= 3 && cfg->max_points > 0 && cfg->max_voxels > 0, "lv_voxelize_host: bad config");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_voxels && h_coords && h_num_points && h_voxel_num, "lv_voxelize_host: null output");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const int64_t n_total = h_frame_offsets[n_frames];
  const size_t V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features, F = n_frames;
  const size_t pt_bytes = (size_t)n_total * C * 4;
  LV_CHECK(h->vox_stage_points.ensure(pt_bytes, st));
  LV_CHECK(h->vox_stage_out[0].ensure(F * V * T * C * 4, st));
  LV_CHECK(h->vox_stage_out[1].ensure(F * V * 3 * 4, st));
  LV_CHECK(h->vox_stage_out[2].ensure(F * V * 4, st));
  LV_CHECK(h->vox_stage_out[3].ensure(F * 4, st));
  if (pt_bytes) LV_CHECK_CUDA(cudaMemcpyAsync(h->vox_stage_points.ptr, h_points, pt_bytes, cudaMemcpyHostToDevice, st));
  LV_CHECK(lv_voxelize(h, cfg, h->vox_stage_points.as<float>(), n_frames, h_frame_offsets, h->vox_stage_out[0].as<float>(),
                       h->vox_stage_out[1].as<int32_t>(), h->vox_stage_out[2].as<int32_t>(),
                       h->vox_stage_out[3].as<int32_t>(), st));
  // voxel_num first: with zero_tail == 0 only the live rows are brought back
  LV_CHECK_CUDA(cudaMemcpyAsync(h_voxel_num, h->vox_stage_out[3].ptr, F * 4, cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  for (size_t f = 0; f < F; ++f) {
    const size_t rows = cfg->zero_tail ? V : (size_t)h_voxel_num[f];
    if (rows == 0) continue;
    LV_CHECK_CUDA(cudaMemcpyAsync(h_voxels + f * V * T * C, h->vox_stage_out[0].as<float>() + f * V * T * C,
                                  rows * T * C * 4, cudaMemcpyDeviceToHost, st));
    LV_CHECK_CUDA(cudaMemcpyAsync(h_coords + f * V * 3, h->vox_stage_out[1].as<int32_t>() + f * V * 3, rows * 3 * 4,
                                  cudaMemcpyDeviceToHost, st));
    LV_CHECK_CUDA(cudaMemcpyAsync(h_num_points + f * V, h->vox_stage_out[2].as<int32_t>() + f * V, rows * 4,
                                  cudaMemcpyDeviceToHost, st));
  }
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  return LV_OK;
}
