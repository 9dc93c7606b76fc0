// Hard voxelizer (sm_100a): exact first-come semantics on a parallel machine.
//
// Replaces spconv.utils.VoxelGeneratorV2.generate as called from
// second/second/data/preprocess.py:299-317; algorithm = the in-tree sibling
// second/second/utils/simplevis.py:9-61.
//
// The reference is a serial loop: voxel ids are handed out in order of first
// appearance, slots inside a voxel in point order.  Atomics give arbitrary order,
// so order is RECONSTRUCTED, per frame, many frames per launch, in six kernels:
//
//   K1 cells     cell(i) = ((cz*gy)+cy)*gx+cx from floorf((p-lo)/vs) (fp32 IEEE sub/div);
//                first[cell] = atomicMin(point index)                 (warp-aggregated)
//   K2 assign    creator(i) := first[cell(i)] == i.  Single-pass scan of the creator flags
//                (decoupled look-back over the chunks of the frame): rank(i) = voxel id.
//                first[cell] := ~rank, creator_cell[rank] = cell, voxel_num = min(total, V),
//                `break` cut-off = index of the creator of rank V.
//   K3 keys      vid(i) = ~first[cell(i)]; kept(i) = vid < V and i < cut; per-chunk histogram
//                of the HIGH digit vid >> L (a "bin" = 2^L consecutive voxel ids)
//   K4 scan      per-frame exclusive scan of the [bin][chunk] histogram
//   K5 scatter   stable multi-split of the kept points into their bins (warp match +
//                per-warp digit counters); creators put EMPTY back into first[]
//                (touched-cell reset, so the dense map never needs a memset)
//   K6 bins      one CTA per (frame, bin): stable rank by the LOW digit inside shared memory
//                -> slot = number of earlier points of the same voxel; the first T point
//                indices of each of the 2^L voxels land in a shared-memory slot table; then
//                one (sub)warp per voxel gathers its points and writes the output row -
//                either the raw (T,C) voxel, or, fused, the PillarFeatureNet decoration
//                (T,C_out) (pointpillars.py:203-231) - plus num_points and coordinates.
//                Every output byte is written exactly once.
//
// Stability of K5 and K6 makes the order inside a voxel the point order for free.
//
// HBM layout: points (N,C) f32 rows; workspace per sub-batch of frames: first[] dense
// int32 map [frames_in_flight][gz*gy*gx] (all EMPTY between calls), cell[], key0[],
// one (key,val) list, creator_cell[], [bin][chunk] histograms, look-back descriptors.
// Outputs padded per frame (lv_voxelize) or concatenated with a batch column
// (lv_voxelize_concat / lv_pillarize_concat = merge_second_batch on the device).
// Algorithmic bytes: 4*C per point read + V*(T*C*4 + 16) written (T*C_out*4 when fused).
#include <math.h>

#include <cooperative_groups.h>

#include "lv_common.cuh"
#include "lv_decorate.cuh"

namespace cg = cooperative_groups;

#define VX_THREADS 256
#define VX_WARPS (VX_THREADS / 32)
#define VX_ITEMS 8
#define VX_CHUNK (VX_THREADS * VX_ITEMS)  // 2048 points per chunk; chunks never straddle frames
#define VX_EMPTY 0x7f7f7f7f               // memset-able "no point yet"
#define VX_CREATOR_BIT 0x40000000         // flag kept in cell[] (grid cells < 2^28)
#define VX_DROPPED 0xffffffffu
#define VX_MAX_BINS 2048
#define VX_MAX_LOW_BITS 10
#ifndef VX_BINS_MINB
#define VX_BINS_MINB 4   // CTAs per SM of the voxel / decorate / mean modes of vx_bins_kernel (64 registers)
#endif

struct VoxParams {
  const float* pts;            // (N_total, C)
  int C;
  const int64_t* frame_off;    // device [F+1] global point offsets
  const int32_t* frame_chunk;  // device [F+1] chunk prefix
  const int32_t* chunk_frame;  // device [chunks] frame of every chunk
  int f0, f1;                  // frames of this sub-batch
  int chunk_lo;                // first chunk of the sub-batch (global chunk id)
  int64_t pt_lo;               // first point of the sub-batch (global point index)
  float lo[3], vs[3];
  int grid[3];                 // gx, gy, gz
  unsigned row_div_m;          // ceil(2^32 / T): voxel of a flat slot index < 2^20
  unsigned div_m[2];           // exact division of a cell id (< 2^28) by gx*gy [0] and by gx [1]:
  int div_s[2];                //   q = (n * m) >> s   (vx_make_div)
  int64_t G;                   // int32 words of first[] per frame: cells (dense map) or 2 x slots (hash table)
  int hash_bits;               // 0: dense map indexed by cell id.  b: open-addressing table of 2^b (key, value) slots per frame
  int map_shift;               // 0 dense, 1 hash: entry e has its value word at e << map_shift (its key word behind it)
  int T, V, overflow, zero_tail;
  unsigned tma_bytes;          // bytes of one full chunk of rows when K1 stages them by TMA, else 0
  int low_bits, n_bins;        // bin = vid >> low_bits
  // workspace (indexed relative to the sub-batch)
  int32_t* map;                // [f1-f0][G]
  int32_t* cell;               // [points]
  uint32_t* key0;              // [points] vid or VX_DROPPED
  uint32_t* keys; int32_t* vals;  // [points] kept points grouped by bin (per frame, at the frame's start)
  int32_t* creator_cell;       // [points] cell of voxel `rank` of a frame at [frame start + rank]
  unsigned long long* chunk_state;  // [chunks] look-back descriptors (flag << 62 | value)
  int32_t* hist;               // [chunks * n_bins], per frame laid out [bin][chunk]
  int32_t* frame_cut;          // [frames] break cut-off (local index)
  int32_t* frame_kept;         // [frames] kept points
  int32_t* bin_start;          // [frames][n_bins + 1] first list position of every bin of a frame (+ the kept total)
  int two_level_scan;          // 1: hist rows hold positions RELATIVE to their bin's start (large frames: one CTA per (frame, bin)
                               //    scans its row, one per frame scans the bin totals); 0: absolute positions (one CTA per frame)
  unsigned* frame_sync;        // [frames][2] arrival counters of the fused prologue (all zero between calls)
  int* err;                    // device flag: a bounded spin of the fused prologue gave up
  // outputs
  float* voxels; int32_t* coords; int32_t* num_points; int32_t* voxel_num;
  int64_t* row_base;           // [F+1] first output row of every frame (concat: prefix of voxel_num)
  int concat;                  // 1: frames back to back, coords carry the batch index (coord_cols == 4)
  int coord_cols;              // 3 (z,y,x) or 4 (b,z,y,x)
  int64_t capacity;            // output rows available
  unsigned long long* prof;    // VX_PROFILE builds: [4] accumulated clocks of the bins kernel phases + [4] CTA count
};

struct ChunkLoc {
  int f;           // global frame id
  int fl;          // frame index inside the sub-batch
  int c;           // chunk index inside the frame
  int nchunks;     // chunks of the frame
  int64_t start;   // global point index of the frame start
  int n;           // points of the frame
};

// cell id -> (cz, cy, cx) without integer division: m = ceil(2^s / d), s = 28 + ceil(log2 d)
// is exact for every n < 2^28 (n * (m*d - 2^s) < n * d <= 2^s).
__device__ __forceinline__ void vx_cell_coords(const VoxParams& p, int c, int& cz, int& cy, int& cx) {
  cz = (int)(((unsigned long long)(unsigned)c * p.div_m[0]) >> p.div_s[0]);
  const int rem = c - cz * (p.grid[0] * p.grid[1]);
  cy = (int)(((unsigned long long)(unsigned)rem * p.div_m[1]) >> p.div_s[1]);
  cx = rem - cy * p.grid[0];
}

// Frame of the CTA's chunk: one broadcast read of the host-built chunk -> frame table (no search,
// no barrier).
__device__ __forceinline__ ChunkLoc vx_locate(const VoxParams& p) {
  ChunkLoc L;
  L.f = __ldg(p.chunk_frame + p.chunk_lo + blockIdx.x);
  L.fl = L.f - p.f0;
  const int cb = __ldg(p.frame_chunk + L.f);
  L.c = p.chunk_lo + blockIdx.x - cb;
  L.nchunks = __ldg(p.frame_chunk + L.f + 1) - cb;
  L.start = __ldg(p.frame_off + L.f);
  L.n = (int)(__ldg(p.frame_off + L.f + 1) - L.start);
  return L;
}

// ---------------------------------------------------------------- K1: cells + first index
// A full chunk (2048 rows) whose first row is 16-byte aligned is staged in shared memory by
// one TMA bulk copy (cp.async.bulk + mbarrier); partial or unaligned chunks use plain loads.
// Open-addressing table for grids whose dense map would not stay in L2 (1056 x 1280 x 40 cells = 216 MB per cloud
// for the 0.05 m SECOND config): 2^b slots of (value word, key word) per frame, a slot is claimed by a 64-bit CAS
// of (cell id << 32 | point index) and lowered by a 64-bit atomicMin - the high word is equal, so the minimum is
// the lowest point index.  Empty = every byte 0x7f, like the dense map.  Linear probing from slot `h` on (the home
// slot is (cell * 2654435761) >> (32 - bits)); the host sizes the table at >= 1.25 x the points of the largest
// frame.  Returns the slot, which replaces the cell id in cell[].
#define VX_EMPTY64 0x7f7f7f7f7f7f7f7full
__device__ __forceinline__ int vx_hash_insert(int32_t* tab, int bits, int cell, int li, unsigned h) {
  unsigned long long* t = reinterpret_cast<unsigned long long*>(tab);
  const unsigned mask = (1u << bits) - 1u;
  const unsigned long long mine = ((unsigned long long)(unsigned)cell << 32) | (unsigned)li;
  while (true) {
    // one round trip per probe: the CAS claims an empty slot and otherwise returns what is there; a slot that already
    // belongs to the cell only needs the atomicMin when it holds a LARGER index (points arrive roughly in index order,
    // so mostly it does not)
    const unsigned long long cur = atomicCAS(t + h, VX_EMPTY64, mine);
    if (cur == VX_EMPTY64) return (int)h;
    if ((unsigned)(cur >> 32) == (unsigned)cell) {
      if (mine < cur) atomicMin(t + h, mine);
      return (int)h;
    }
    h = (h + 1) & mask;
  }
}

template <bool C4, bool HASH>
__global__ void __launch_bounds__(VX_THREADS, HASH ? 5 : 1) vx_cells_kernel(VoxParams p) {
  extern __shared__ __align__(128) float tile[];  // [VX_CHUNK][C] when p.tma_bytes != 0
  __shared__ __align__(8) uint64_t bar;
  const ChunkLoc L = vx_locate(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = p.pts + (L.start + (int64_t)L.c * VX_CHUNK) * p.C;
  const bool staged = p.tma_bytes != 0 && (L.c + 1) * VX_CHUNK <= L.n && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (threadIdx.x == 0) {
    if (staged) {
      lv_mbar_init(&bar, 1);
      lv_mbar_init_fence();
      lv_mbar_expect_tx(&bar, p.tma_bytes);
      lv_tma_load_1d(tile, src, p.tma_bytes, &bar);
    }
    p.chunk_state[blockIdx.x] = 0ull;                // look-back descriptor of this chunk: invalid
    if (L.c == 0) p.frame_cut[L.fl] = 0x7fffffff;
  }
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  if (staged) {
    __syncthreads();  // the barrier is initialised
    lv_mbar_wait(&bar, 0);
  }
  const int base = L.c * VX_CHUNK + warp * (32 * VX_ITEMS);
  // table form: the first probes of all eight rounds are issued before the first answer is looked at (eight CAS round
  // trips in flight per thread instead of one after the other - the table of a batch of clouds does not stay in L2)
  int h_cell[HASH ? VX_ITEMS : 1], h_lead[HASH ? VX_ITEMS : 1];
  unsigned h_slot[HASH ? VX_ITEMS : 1];
  unsigned long long h_cur[HASH ? VX_ITEMS : 1];
  constexpr int UNROLL = HASH ? VX_ITEMS : 2;   // (the table form keeps its eight answers in registers)
#pragma unroll UNROLL
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int ti = warp * (32 * VX_ITEMS) + r * 32 + lane;  // row inside the chunk
    const int li = base + r * 32 + lane;                    // local point index inside the frame
    int cell = -1;
    if (li < L.n) {
      const int64_t gi = L.start + li;
      float x, y, z;
      if (C4) {
        float4 v;
        if (staged) v = reinterpret_cast<const float4*>(tile)[ti];
        else v = __ldg(reinterpret_cast<const float4*>(p.pts) + gi);
        x = v.x; y = v.y; z = v.z;
      } else if (staged) {
        const float* q = tile + ti * p.C;
        x = q[0]; y = q[1]; z = q[2];
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
      if (!HASH) p.cell[gi - p.pt_lo] = cell;
    }
    // lanes of one cell: only the lowest lane (= lowest index) needs to bid
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int leader = __ffs(peers) - 1;
    if (HASH) {
      h_cell[r] = cell;
      h_lead[r] = leader;
      h_slot[r] = ((unsigned)cell * 2654435761u) >> (32 - p.hash_bits);
      h_cur[r] = 0ull;
      if (cell >= 0 && lane == leader)
        h_cur[r] = atomicCAS(reinterpret_cast<unsigned long long*>(map) + h_slot[r], VX_EMPTY64,
                             ((unsigned long long)(unsigned)cell << 32) | (unsigned)li);
    } else if (cell >= 0 && lane == leader) {
      atomicMin(map + cell, li);
    }
  }
  if (HASH) {
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      const int li = base + r * 32 + lane, cell = h_cell[r];
      int slot = -1;
      if (cell >= 0 && lane == h_lead[r]) {
        const unsigned long long mine = ((unsigned long long)(unsigned)cell << 32) | (unsigned)li, cur = h_cur[r];
        if (cur == VX_EMPTY64) {
          slot = (int)h_slot[r];                       // claimed by the first probe
        } else if ((unsigned)(cur >> 32) == (unsigned)cell) {
          if (mine < cur) atomicMin(reinterpret_cast<unsigned long long*>(map) + h_slot[r], mine);
          slot = (int)h_slot[r];
        } else {                                       // collision: probe on from the next slot
          slot = vx_hash_insert(map, p.hash_bits, cell, li, (h_slot[r] + 1) & ((1u << p.hash_bits) - 1u));
        }
      }
      slot = __shfl_sync(0xffffffffu, slot, h_lead[r]);
      if (li < L.n) p.cell[L.start + li - p.pt_lo] = cell >= 0 ? slot : -1;
    }
  }
}

// ---------------------------------------------------------------- K2: voxel ids (single-pass scan)
#define VX_FLAG_AGG (1ull << 62)
#define VX_FLAG_PREFIX (2ull << 62)

__device__ __forceinline__ unsigned long long vx_ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void vx_st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 8 CTAs per SM (32 registers, no spills): the kernel waits for its gathers and its look-back, so residency is what it needs -
// pillarize stage 0.660 -> 0.654 ms against the 40-register build at 6 CTAs per SM
__global__ void __launch_bounds__(VX_THREADS, 8) vx_assign_kernel(VoxParams p) {
  __shared__ int sw[VX_WARPS];
  __shared__ int s_excl;
  const ChunkLoc L = vx_locate(p);
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread-contiguous items: thread t owns local points base + t*8 .. +7 (index order)
  const int base = L.c * VX_CHUNK + threadIdx.x * VX_ITEMS;
  int cell[VX_ITEMS];
  unsigned flags = 0;
  int s = 0;
  // two batches of independent loads: the eight cells, then the eight first[] entries
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k;
    cell[k] = li < L.n ? p.cell[L.start + li - p.pt_lo] : -1;
  }
  int first[VX_ITEMS];
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) first[k] = cell[k] >= 0 ? map[(int64_t)cell[k] << p.map_shift] : -1;
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    if (cell[k] >= 0 && first[k] == base + k) {
      flags |= 1u << k;
      ++s;
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
  int woff = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < VX_WARPS; ++w) {
    if (w < warp) woff += sw[w];
    agg += sw[w];
  }
  // decoupled look-back over the previous chunks of this frame (warp 0)
  if (warp == 0) {
    unsigned long long* st = p.chunk_state + blockIdx.x;
    int excl = 0;
    if (L.c == 0) {
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | (unsigned)agg);
    } else {
      if (lane == 0) vx_st_state(st, VX_FLAG_AGG | (unsigned)agg);
      int look = L.c - 1;  // chunk index (inside the frame) the window ends at
      while (true) {
        const int idx = look - lane;
        unsigned long long v = VX_FLAG_PREFIX;  // lanes before the frame start act as a zero prefix
        if (idx >= 0) {
          do { v = vx_ld_state(st - (L.c - idx)); } while ((v >> 62) == 0);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_p = __ffs(is_prefix) - 1;  // nearest lane holding an inclusive prefix
        // aggregates of the lanes nearer than the prefix + the prefix itself; all 32 if none yet
        int contrib = (first_p < 0 || lane <= first_p) ? (int)(unsigned)(v & 0xffffffffull) : 0;
        if (idx < 0) contrib = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (first_p >= 0) break;  // a prefix (real, or the virtual one before chunk 0) was in the window
        look -= 32;
      }
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | (unsigned)(excl + agg));
    }
    if (lane == 0) {
      s_excl = excl;
      if (L.c == L.nchunks - 1) {
        const int total = excl + agg;
        p.voxel_num[L.f] = total < p.V ? total : p.V;
      }
    }
  }
  __syncthreads();
  if (flags == 0) return;
  int rank = s_excl + woff + inc - s;
  int32_t* ccell = p.creator_cell + (L.start - p.pt_lo);
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    if (!(flags & (1u << k))) continue;
    const int li = base + k;
    const int c = cell[k];
    map[(int64_t)c << p.map_shift] = ~rank;  // voxel id, stored negative so it never equals a point index
    p.cell[L.start + li - p.pt_lo] = c | VX_CREATOR_BIT;
    ccell[rank] = p.map_shift ? map[((int64_t)c << 1) + 1] : c;  // the voxel's cell id (hash: the slot's key word); rank < points of the frame
    if (rank == p.V && p.overflow == LV_OVERFLOW_BREAK) p.frame_cut[L.fl] = li;  // simplevis.py:48-49
    ++rank;
  }
}

// ---------------------------------------------------------------- K3: keys + bin histogram
__global__ void __launch_bounds__(VX_THREADS) vx_keys_kernel(VoxParams p) {
  extern __shared__ int hist[];  // [n_bins]
  const ChunkLoc L = vx_locate(p);
  const int32_t* map = p.map + (int64_t)L.fl * p.G;
  const int D = p.n_bins;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) hist[d] = 0;
  __syncthreads();
  const int cut = p.frame_cut[L.fl];
  const int base = L.c * VX_CHUNK;
  const int lane = threadIdx.x & 31;
  // two batches of independent loads: the eight cells, then the eight voxel ids
  int craw[VX_ITEMS], vid[VX_ITEMS];
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k * VX_THREADS + threadIdx.x;
    craw[k] = li < L.n ? p.cell[L.start + li - p.pt_lo] : -1;
  }
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) vid[k] = craw[k] >= 0 ? ~map[(int64_t)(craw[k] & ~VX_CREATOR_BIT) << p.map_shift] : 0x7fffffff;
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k * VX_THREADS + threadIdx.x;
    unsigned key = VX_DROPPED;
    if (li < L.n) {
      if (vid[k] < p.V && li < cut) key = (unsigned)vid[k];
      p.key0[L.start + li - p.pt_lo] = key;
    }
    const unsigned digit = key == VX_DROPPED ? 0xffffffffu : (key >> p.low_bits);
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    if (key != VX_DROPPED && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
  }
  __syncthreads();
  int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) tab[(int64_t)d * L.nchunks + L.c] = hist[d];
}

// ---------------------------------------------------------------- K4: per-frame scan of [bin][chunk]
// in-place exclusive scan of `n` ints by one CTA; returns the total.
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
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    int woff = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < VX_WARPS; ++w) {
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

__global__ void __launch_bounds__(VX_THREADS) vx_scan_hist_kernel(VoxParams p) {
  __shared__ int sw[VX_WARPS];
  __shared__ int carry;
  const int fl = blockIdx.x, f = p.f0 + fl;
  const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  const int total = vx_cta_exclusive_scan(p.hist + (int64_t)cb * p.n_bins, nch * p.n_bins, sw, &carry);
  if (threadIdx.x == 0) {
    p.frame_kept[fl] = total;
    if (nch == 0) p.voxel_num[f] = 0;  // a frame without points never ran K2
  }
  // first list position of every bin (what K6 reads), the kept total behind them
  __syncthreads();
  int32_t* bstart = p.bin_start + (int64_t)fl * (p.n_bins + 1);
  for (int d = threadIdx.x; d < p.n_bins; d += VX_THREADS) bstart[d] = nch > 0 ? p.hist[(int64_t)cb * p.n_bins + (int64_t)d * nch] : 0;
  if (threadIdx.x == 0) bstart[p.n_bins] = total;
}

// K4 in two levels for frames with many chunks (a 1.06 M-point cloud has 519: one CTA would scan 30-120 k counters
// alone): grid (n_bins, frames) - every CTA scans the chunk row of ITS bin in place and leaves the bin's total in
// bin_start[bin]; then one CTA per frame turns the totals into bin starts.  K5 adds the two.
__global__ void __launch_bounds__(VX_THREADS) vx_scan_rows_kernel(VoxParams p) {
  __shared__ int sw[VX_WARPS];
  __shared__ int carry;
  const int d = blockIdx.x, fl = blockIdx.y, f = p.f0 + fl;
  const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  const int total = nch > 0 ? vx_cta_exclusive_scan(p.hist + (int64_t)cb * p.n_bins + (int64_t)d * nch, nch, sw, &carry) : 0;
  if (threadIdx.x == 0) p.bin_start[(int64_t)fl * (p.n_bins + 1) + d] = total;
}
__global__ void __launch_bounds__(VX_THREADS) vx_scan_bins_kernel(VoxParams p) {
  __shared__ int sw[VX_WARPS];
  __shared__ int carry;
  const int fl = blockIdx.x, f = p.f0 + fl;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  int32_t* bstart = p.bin_start + (int64_t)fl * (p.n_bins + 1);
  const int total = vx_cta_exclusive_scan(bstart, p.n_bins, sw, &carry);
  if (threadIdx.x == 0) {
    bstart[p.n_bins] = total;
    p.frame_kept[fl] = total;
    if (nch == 0) p.voxel_num[f] = 0;  // a frame without points never ran K2
  }
}

// ---------------------------------------------------------------- K5: stable multi-split into bins
// Warp w owns items [w*256, (w+1)*256) of the chunk in 8 rounds of 32 lanes, so
// (warp, round, lane) order is point order.  Per-warp bin counters live in shared memory;
// after the rounds an exclusive scan over the warps (plus the scanned global table) turns
// them into the warp's base for each bin.
__global__ void __launch_bounds__(VX_THREADS) vx_scatter_kernel(VoxParams p) {
  extern __shared__ int cnt[];  // [8][n_bins]
  const ChunkLoc L = vx_locate(p);
  const int D = p.n_bins;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < VX_WARPS * D; i += VX_THREADS) cnt[i] = 0;
  __syncthreads();
  int* mycnt = cnt + warp * D;
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  unsigned key[VX_ITEMS];
  int rnk[VX_ITEMS];
  const int base = L.c * VX_CHUNK + warp * (32 * VX_ITEMS);
  const unsigned lt = lv_lanemask_lt();
  // all sixteen loads first (keys and cells), then the touched-cell reset: K3 was the last
  // reader of first[]
  {
    int craw[VX_ITEMS];
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      const int li = base + r * 32 + lane;
      const int64_t wi = L.start + li - p.pt_lo;
      key[r] = li < L.n ? p.key0[wi] : VX_DROPPED;
      craw[r] = li < L.n ? p.cell[wi] : -1;
    }
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r)
      if (craw[r] >= 0 && (craw[r] & VX_CREATOR_BIT)) {
        const int64_t e = (int64_t)(craw[r] & ~VX_CREATOR_BIT) << p.map_shift;
        map[e] = VX_EMPTY;
        if (p.map_shift) map[e + 1] = VX_EMPTY;
      }
  }
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const bool live = key[r] != VX_DROPPED;
    const unsigned digit = live ? (key[r] >> p.low_bits) : 0xffffffffu;
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
  const int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  const int32_t* bstart = p.bin_start + (int64_t)L.fl * (D + 1);
  for (int d = threadIdx.x; d < D; d += VX_THREADS) {
    int off = tab[(int64_t)d * L.nchunks + L.c];
    if (p.two_level_scan) off += bstart[d];
#pragma unroll
    for (int w = 0; w < VX_WARPS; ++w) {
      const int t = cnt[w * D + d];
      cnt[w * D + d] = off;
      off += t;
    }
  }
  __syncthreads();
  const int64_t out0 = L.start - p.pt_lo;  // the frame's list starts at its first point slot
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    if (key[r] == VX_DROPPED) continue;
    const int64_t pos = out0 + mycnt[key[r] >> p.low_bits] + rnk[r];
    p.keys[pos] = key[r];
    p.vals[pos] = base + r * 32 + lane;
  }
}

// ---------------------------------------------------------------- KX: K1-K5 in ONE launch (frames of <= 64 chunks)
// One CTA keeps its 2,048 points in registers from the coordinate rule to the final multi-split position; what the
// kernel boundaries of K1-K5 ordered is ordered by device-side mechanisms instead:
//   * frame barrier 1 (arrival counter per frame) between the atomicMin bids and the creator test,
//   * the decoupled look-back of K2 for the voxel ids,
//   * polling: a non-creator spins on first[cell] until its creator - always an EARLIER point, i.e. the same or an
//     earlier chunk - has replaced the point index by ~rank,
//   * frame barrier 2 between the publication of the per-chunk bin counts and their prefix; the [bin][chunk]
//     table of a frame (<= 64 chunks) is read by every CTA of the frame instead of being scanned by a kernel.
// Needs every CTA of a frame resident at the same time (the host checks chunks-per-frame against the occupancy of
// the kernel); CTAs are dispatched in blockIdx order, so the earliest unfinished frame always has all of its CTAs
// on the machine - the same forward-progress assumption as the look-back.  cell[] and key0[] are never written.
// Spins are bounded (~1 s) and raise p.err instead of hanging the GPU.
// OPT-IN (lv_set_option "vox_fused_prologue" 1): bit-identical to K1-K5 (tests/test_gpu_voxel_fused.py), but measured
// SLOWER on 128 C5 frames - pillarize stage 0.724 vs 0.671 ms; ncu: 228 us against 168 us for the five kernels.  43 %
// of its stall samples sit in the frame barriers: every phase keeps the latency it had as a kernel of its own (a
// chunk's loads and bids take ~15 us under load whether or not a kernel boundary follows), while the CTA now holds
// 48 registers x 256 threads through all of them (5 CTAs/SM instead of 8) and waits for the slowest chunk of its
// frame twice.  Re-reading cell[] / first[] / key0[] between kernels costs little: they are L2 hits.
#ifndef VX_FUSED_MINB
#define VX_FUSED_MINB 5
#endif
#define VX_FUSED_MAX_CHUNKS 64
#define VX_FUSED_MAX_BINS 1024
#define VX_FUSED_BINS_PER_THREAD (VX_FUSED_MAX_BINS / VX_THREADS)

__device__ __forceinline__ int vx_ld_relaxed(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void vx_st_relaxed(int32_t* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned vx_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// every CTA of the frame arrives (release: barrier, fence, atomic) and waits for `target` arrivals (acquire)
__device__ __forceinline__ void vx_frame_barrier(unsigned* ctr, unsigned target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (vx_ld_acquire(ctr) < target) {
      __nanosleep(64);
      if (clock64() - t0 > (1ll << 31)) {
        *err = 1;
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <bool C4>
__global__ void __launch_bounds__(VX_THREADS, VX_FUSED_MINB) vx_fused_kernel(VoxParams p) {
  extern __shared__ __align__(128) int xsm[];    // the TMA tile [VX_CHUNK][C], then cnt[8][D] | binbase[D] | table [D][chunks]
  __shared__ __align__(8) uint64_t bar;
  __shared__ int sw[VX_WARPS];
  __shared__ int s_excl, s_cut;
  const ChunkLoc L = vx_locate(p);
  const int D = p.n_bins;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = lv_lanemask_lt();
  int32_t* map = p.map + (int64_t)L.fl * p.G;
  unsigned* sync = p.frame_sync + 2 * L.fl;
  const float* tile = reinterpret_cast<const float*>(xsm);
  const float* src = p.pts + (L.start + (int64_t)L.c * VX_CHUNK) * p.C;
  const bool staged = p.tma_bytes != 0 && (L.c + 1) * VX_CHUNK <= L.n && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (threadIdx.x == 0 && staged) {
    lv_mbar_init(&bar, 1);
    lv_mbar_init_fence();
    lv_mbar_expect_tx(&bar, p.tma_bytes);
    lv_tma_load_1d(xsm, src, p.tma_bytes, &bar);
  }
  if (staged) {
    __syncthreads();  // the barrier is initialised
    lv_mbar_wait(&bar, 0);
  }
  // ---- K1: cells (kept in registers) and the bids for first[]
  const int base = L.c * VX_CHUNK + warp * (32 * VX_ITEMS);   // (warp, round, lane) order is point order
  int cell[VX_ITEMS];
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int ti = warp * (32 * VX_ITEMS) + r * 32 + lane;
    const int li = base + r * 32 + lane;
    cell[r] = -1;
    if (li < L.n) {
      const int64_t gi = L.start + li;
      float x, y, z;
      if (C4) {
        float4 v;
        if (staged) v = reinterpret_cast<const float4*>(tile)[ti];
        else v = __ldg(reinterpret_cast<const float4*>(p.pts) + gi);
        x = v.x; y = v.y; z = v.z;
      } else if (staged) {
        const float* q = tile + ti * p.C;
        x = q[0]; y = q[1]; z = q[2];
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
        cell[r] = ((int)cz * p.grid[1] + (int)cy) * p.grid[0] + (int)cx;
    }
  }
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const unsigned peers = __match_any_sync(0xffffffffu, cell[r]);
    if (cell[r] >= 0 && lane == __ffs(peers) - 1) atomicMin(map + cell[r], base + r * 32 + lane);
  }
  vx_frame_barrier(sync, (unsigned)L.nchunks, p.err);      // every bid of the frame has landed
  // the tile is consumed: its shared memory becomes the per-warp bin counters
  int* cnt = xsm;                     // [8][D]
  int* binbase = xsm + VX_WARPS * D;  // [D]
  for (int i = threadIdx.x; i < VX_WARPS * D; i += VX_THREADS) cnt[i] = 0;

  // ---- K2: creators, voxel ids in first-come order
  unsigned bal[VX_ITEMS];
  int vid[VX_ITEMS];              // first[cell] as read (a point index, or ~id once the creator has ranked), then the id
  int wc = 0;
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) vid[r] = cell[r] >= 0 ? vx_ld_relaxed(map + cell[r]) : -1;
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    bal[r] = __ballot_sync(0xffffffffu, cell[r] >= 0 && vid[r] == base + r * 32 + lane);
    wc += __popc(bal[r]);
  }
  if (lane == 0) sw[warp] = wc;
  __syncthreads();
  int woff = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < VX_WARPS; ++w) {
    if (w < warp) woff += sw[w];
    agg += sw[w];
  }
  if (warp == 0) {
    unsigned long long* st = p.chunk_state + blockIdx.x;
    int excl = 0;
    if (L.c == 0) {
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | (unsigned)agg);
    } else {
      if (lane == 0) vx_st_state(st, VX_FLAG_AGG | (unsigned)agg);
      int look = L.c - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long v = VX_FLAG_PREFIX;
        if (idx >= 0) {
          do { v = vx_ld_state(st - (L.c - idx)); } while ((v >> 62) == 0);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_p = __ffs(is_prefix) - 1;
        int contrib = (first_p < 0 || lane <= first_p) ? (int)(unsigned)(v & 0xffffffffull) : 0;
        if (idx < 0) contrib = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (first_p >= 0) break;
        look -= 32;
      }
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | (unsigned)(excl + agg));
    }
    if (lane == 0) {
      s_excl = excl;
      // `break` (simplevis.py:48-49): the creator of id V cuts the frame.  It lies in an earlier chunk (every
      // point here is past it), in this chunk (it writes s_cut below), or further on / nowhere.
      s_cut = (p.overflow == LV_OVERFLOW_BREAK && excl > p.V) ? 0 : 0x7fffffff;
      if (L.c == L.nchunks - 1) {
        const int total = excl + agg;
        p.voxel_num[L.f] = total < p.V ? total : p.V;
      }
    }
  }
  __syncthreads();
  {
    int run = s_excl + woff;
    int32_t* ccell = p.creator_cell + (L.start - p.pt_lo);
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      if ((bal[r] >> lane) & 1u) {
        const int rank = run + __popc(bal[r] & lt);
        vid[r] = ~rank;
        vx_st_relaxed(map + cell[r], ~rank);  // voxel id, stored negative so it never equals a point index
        ccell[rank] = cell[r];                // rank < number of points of the frame
        if (rank == p.V && p.overflow == LV_OVERFLOW_BREAK) s_cut = base + r * 32 + lane;
      }
      run += __popc(bal[r]);
    }
  }
  __syncthreads();
  // ---- K3: voxel id of every other point (its creator is an earlier point: poll until the id is there)
  const int cut = s_cut;
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    if (cell[r] >= 0 && vid[r] >= 0) {
      const long long t0 = clock64();
      int v;
      while ((v = vx_ld_relaxed(map + cell[r])) >= 0) {
        __nanosleep(32);
        if (clock64() - t0 > (1ll << 31)) {
          *p.err = 2;
          v = -1;
          break;
        }
      }
      vid[r] = v;
    }
    vid[r] = ~vid[r];   // (cell < 0: ~(-1) = 0, dropped below)
    if (cell[r] < 0 || vid[r] >= p.V || base + r * 32 + lane >= cut) vid[r] = -1;   // dropped
  }
  // ---- K5 (first half): stable rank inside (warp, bin)
  int* mycnt = cnt + warp * D;
  int rnk[VX_ITEMS];
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const unsigned digit = vid[r] < 0 ? 0xffffffffu : ((unsigned)vid[r] >> p.low_bits);
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (vid[r] >= 0 && lane == leader) {
      old = mycnt[digit];
      mycnt[digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  // per-bin counts of the chunk -> the frame's [bin][chunk] table; counters become offsets inside the CTA
  int32_t* tab = p.hist + (int64_t)(__ldg(p.frame_chunk + L.f) - p.chunk_lo) * D;
  for (int d = threadIdx.x; d < D; d += VX_THREADS) {
    int off = 0;
#pragma unroll
    for (int w = 0; w < VX_WARPS; ++w) {
      const int t = cnt[w * D + d];
      cnt[w * D + d] = off;
      off += t;
    }
    tab[(int64_t)d * L.nchunks + L.c] = off;
  }
  vx_frame_barrier(sync + 1, (unsigned)L.nchunks, p.err);  // the table of the frame is complete; first[] is read out
  // touched-cell reset and the look-back descriptor back to "invalid": nobody of the frame needs them any more
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r)
    if ((bal[r] >> lane) & 1u) map[cell[r]] = VX_EMPTY;
  if (threadIdx.x == 0) p.chunk_state[blockIdx.x] = 0ull;
  // ---- K4: this chunk's base inside every bin = bin start + counts of the earlier chunks.  The frame's table
  //      comes to shared memory in one batch of independent coalesced loads; thread t owns bins [t*PER, (t+1)*PER)
  {
    int* stab = binbase + D;        // [D][nchunks]
    const int n_tab = D * L.nchunks;
    for (int i = threadIdx.x; i < n_tab; i += VX_THREADS) stab[i] = __ldcg(tab + i);
    __syncthreads();
    const int PER = (D + VX_THREADS - 1) / VX_THREADS;   // <= VX_FUSED_BINS_PER_THREAD
    int tot[VX_FUSED_BINS_PER_THREAD], below[VX_FUSED_BINS_PER_THREAD];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < VX_FUSED_BINS_PER_THREAD; ++j) {
      tot[j] = 0; below[j] = 0;
      const int d = threadIdx.x * PER + j;
      if (j < PER && d < D) {
        const int* row = stab + d * L.nchunks;
        for (int c = 0; c < L.nchunks; ++c) {
          const int t = row[c];
          tot[j] += t;
          if (c < L.c) below[j] += t;
        }
      }
      mine += tot[j];
    }
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sw[warp] = inc;
    __syncthreads();
    int wbase = 0, kept = 0;
#pragma unroll
    for (int w = 0; w < VX_WARPS; ++w) {
      if (w < warp) wbase += sw[w];
      kept += sw[w];
    }
    int run = wbase + inc - mine;
    int32_t* bstart = p.bin_start + (int64_t)L.fl * (D + 1);
#pragma unroll
    for (int j = 0; j < VX_FUSED_BINS_PER_THREAD; ++j) {
      const int d = threadIdx.x * PER + j;
      if (j < PER && d < D) {
        binbase[d] = run + below[j];
        if (L.c == 0) bstart[d] = run;
        run += tot[j];
      }
    }
    if (L.c == 0 && threadIdx.x == 0) bstart[D] = kept;
  }
  __syncthreads();
  // ---- K5 (second half): the kept points grouped by bin, in point order
  const int64_t out0 = L.start - p.pt_lo;
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    if (vid[r] < 0) continue;
    const unsigned digit = (unsigned)vid[r] >> p.low_bits;
    const int64_t pos = out0 + binbase[digit] + mycnt[digit] + rnk[r];
    p.keys[pos] = (unsigned)vid[r];
    p.vals[pos] = base + r * 32 + lane;
  }
}

// ---------------------------------------------------------------- KF: K1-K5 of one frame inside a thread-block cluster
// Small grids (the pillar configs: 400 x 400 x 1 cells) fit the dense first[] map of a frame into the shared memory
// of a cluster of CS CTAs (distributed shared memory): cell c lives in CTA c / part at offset c % part.  One cluster
// of CS x 1024 threads then does K1-K5 for its frame without a single L2 atomic or map gather:
//   A  map := EMPTY                                                                  (cluster barrier)
//   B  cell(i); atomicMin(map[cell], i) on the owner CTA's shared memory             (cluster barrier)
//   C  creator(i) = map[cell(i)] == i; creators per warp -> wtot[] of every CTA      (cluster barrier)
//   D  rank = exclusive count of earlier creators; map[cell] := ~rank, creator_cell, voxel_num, cut      (barrier)
//   E  vid = ~map[cell]; key; stable rank inside the warp per bin (match_any + per-warp counters)
//   F  counters -> bases: exclusive over the warps of the CTA, CTA totals to every CTA (barrier), over CTAs, bins
//   G  keys / vals scattered to their bin positions (what K5 writes), bin table and frame_kept for K6
// Warp w of the cluster owns the contiguous points [w * span, (w + 1) * span), round by round, so (warp, round,
// lane) order is point order.  The map never exists in global memory, so there is nothing to reset.
// OPT-IN (lv_set_option "vox_frame_kernel" 1): bit-identical to K1-K5 (tests/test_gpu_voxel.py runs both), but
// measured SLOWER on 128 C5 frames - 0.71 vs 0.67 ms for the pillarize stage.  clock64 per phase (-DVX_PROFILE),
// one cluster of 4 x 1024 threads per 53,146-point frame: A 2.8k, B 37k, C 14k, D 10k, E 15k, F 9.5k, G 7k clocks
// = 49 us per frame, 33 clusters at a time.  The scattered 4-byte remote shared-memory operations (three of four
// cells live in another CTA) run at roughly one per two cycles and SM; a variant that partitions the POINTS by
// owner CTA so that every map access is local is the next thing to try.
#define VF_THREADS 1024
#define VF_WARPS (VF_THREADS / 32)
#define VF_MAX_CS 8
#define VF_ROUNDS 16   // rounds of 32 points a warp keeps in registers: frames of up to CS x 16,384 points

__device__ __forceinline__ int* vf_remote(cg::cluster_group& cluster, int* map, int cell, int part) {
  const int owner = cell / part;
  return cluster.map_shared_rank(map + (cell - owner * part), owner);
}

template <bool C4>
__global__ void __launch_bounds__(VF_THREADS, 1) vx_frame_kernel(VoxParams p, int part) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int fl = blockIdx.x / CS, f = p.f0 + fl;
  const int D = p.n_bins;
  extern __shared__ int fsm[];
  int* map = fsm;                         // [part]
  int* cnt = map + part;                  // [32][D]
  int* ctot = cnt + VF_WARPS * D;         // [VF_MAX_CS][D]  per-CTA bin totals, replicated in every CTA
  int* dbase = ctot + VF_MAX_CS * D;      // [D]
  int* wtot = dbase + D;                  // [VF_MAX_CS * 32] creators per warp, replicated in every CTA
  int* misc = wtot + VF_MAX_CS * VF_WARPS;  // [0] break cut-off
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = crank * VF_WARPS + warp, NW = CS * VF_WARPS;
  const int64_t start = __ldg(p.frame_off + f);
  const int n = (int)(__ldg(p.frame_off + f + 1) - start);
  const int span = (((n + NW - 1) / NW) + 31) & ~31;
  const int lo = gw * span;
  const int hi = min(lo + span, n);       // rounds: base = lo, lo + 32, ... < hi
  const int64_t wi0 = start - p.pt_lo;    // workspace index of the frame's first point
  const unsigned lt = lv_lanemask_lt();

#ifdef VX_PROFILE
  long long tq[8]; int nq = 0;
#define VF_MARK() tq[nq++] = clock64()
#else
#define VF_MARK()
#endif
  VF_MARK();
  // ---- A
  for (int i = threadIdx.x; i < part; i += VF_THREADS) map[i] = VX_EMPTY;
  for (int i = threadIdx.x; i < VF_WARPS * D; i += VF_THREADS) cnt[i] = 0;
  if (threadIdx.x == 0) misc[0] = 0x7fffffff;
  cluster.sync();
  VF_MARK();

  // ---- B: the warp's points stay in registers from here on (VF_ROUNDS rounds of 32: the host checks the span)
  int cell[VF_ROUNDS];
#pragma unroll
  for (int r = 0; r < VF_ROUNDS; ++r) {
    const int li = lo + r * 32 + lane;
    cell[r] = -1;
    if (li < hi) {
      const int64_t gi = start + li;
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
        cell[r] = ((int)cz * p.grid[1] + (int)cy) * p.grid[0] + (int)cx;
    }
  }
#pragma unroll
  for (int r = 0; r < VF_ROUNDS; ++r) {
    if (lo + r * 32 >= hi) break;   // warp-uniform
    const unsigned peers = __match_any_sync(0xffffffffu, cell[r]);
    if (cell[r] >= 0 && lane == __ffs(peers) - 1) atomicMin(vf_remote(cluster, map, cell[r], part), lo + r * 32 + lane);
  }
  cluster.sync();
  VF_MARK();

  // ---- C: creator flags, one bit per round
  unsigned creator = 0;
  {
    int first[VF_ROUNDS];
#pragma unroll
    for (int r = 0; r < VF_ROUNDS; ++r) first[r] = cell[r] >= 0 ? *vf_remote(cluster, map, cell[r], part) : -1;
#pragma unroll
    for (int r = 0; r < VF_ROUNDS; ++r)
      if (cell[r] >= 0 && first[r] == lo + r * 32 + lane) creator |= 1u << r;
  }
  int mine = __popc(creator);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane < CS) cluster.map_shared_rank(wtot, lane)[gw] = mine;
  cluster.sync();
  VF_MARK();

  // ---- D: voxel ids in first-come order
  int excl = 0;
  for (int w = lane; w < gw; w += 32) excl += wtot[w];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) excl += __shfl_xor_sync(0xffffffffu, excl, o);
  if (gw == NW - 1 && lane == 0) {
    const int total = excl + mine;
    p.voxel_num[f] = total < p.V ? total : p.V;
  }
  if (mine > 0) {   // warp-uniform
    int32_t* ccell = p.creator_cell + wi0;
    int run = excl;
#pragma unroll
    for (int r = 0; r < VF_ROUNDS; ++r) {
      const bool cr = (creator >> r) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, cr);
      if (cr) {
        const int rank = run + __popc(bal & lt);
        *vf_remote(cluster, map, cell[r], part) = ~rank;   // voxel id, stored negative so it never equals a point index
        ccell[rank] = cell[r];                             // rank < number of points of the frame
        if (rank == p.V && p.overflow == LV_OVERFLOW_BREAK)   // simplevis.py:48-49
          for (int c = 0; c < CS; ++c) cluster.map_shared_rank(misc, c)[0] = lo + r * 32 + lane;
      }
      run += __popc(bal);
    }
  }
  cluster.sync();
  VF_MARK();

  // ---- E: keys and the stable rank inside (warp, bin); cell[] now holds the key
  const int cut = misc[0];
  int* mycnt = cnt + warp * D;
  int rnk[VF_ROUNDS];
#pragma unroll
  for (int r = 0; r < VF_ROUNDS; ++r) {
    int vid = 0x7fffffff;
    if (cell[r] >= 0) vid = ~*vf_remote(cluster, map, cell[r], part);
    cell[r] = (vid < p.V && lo + r * 32 + lane < cut) ? vid : -1;
  }
#pragma unroll
  for (int r = 0; r < VF_ROUNDS; ++r) {
    rnk[r] = 0;
    if (lo + r * 32 >= hi) break;   // warp-uniform
    const unsigned digit = cell[r] < 0 ? 0xffffffffu : ((unsigned)cell[r] >> p.low_bits);
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (cell[r] >= 0 && lane == leader) {
      old = mycnt[digit];
      mycnt[digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  VF_MARK();

  // ---- F
  for (int d = threadIdx.x; d < D; d += VF_THREADS) {
    int acc = 0;
    for (int w = 0; w < VF_WARPS; ++w) {
      const int t = cnt[w * D + d];
      cnt[w * D + d] = acc;
      acc += t;
    }
    for (int c = 0; c < CS; ++c) cluster.map_shared_rank(ctot, c)[crank * D + d] = acc;
  }
  cluster.sync();
  for (int d = threadIdx.x; d < D; d += VF_THREADS) {
    int tot = 0;
    for (int c = 0; c < CS; ++c) tot += ctot[c * D + d];
    dbase[d] = tot;
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the bin totals (D <= a few hundred)
    int carry = 0;
    for (int d0 = 0; d0 < D; d0 += 32) {
      const int d = d0 + lane;
      const int v = d < D ? dbase[d] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (d < D) dbase[d] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (crank == 0 && lane == 0) {
      p.frame_kept[fl] = carry;
      p.bin_start[(int64_t)fl * (D + 1) + D] = carry;
    }
  }
  __syncthreads();
  {
    const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
    const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
    int32_t* tab = p.hist + (int64_t)cb * D;
    for (int d = threadIdx.x; d < D; d += VF_THREADS) {
      int below = dbase[d];
      for (int c = 0; c < crank; ++c) below += ctot[c * D + d];
      for (int w = 0; w < VF_WARPS; ++w) cnt[w * D + d] += below;
      if (crank == 0) p.bin_start[(int64_t)fl * (D + 1) + d] = dbase[d];   // what K6 reads: first position of bin d
      if (crank == 0 && nch > 0) tab[(int64_t)d * nch] = dbase[d];
    }
  }
  __syncthreads();
  VF_MARK();

  // ---- G: what K5 writes - the kept points grouped by bin, in point order
#pragma unroll
  for (int r = 0; r < VF_ROUNDS; ++r) {
    if (cell[r] < 0) continue;
    const int64_t pos = wi0 + mycnt[(unsigned)cell[r] >> p.low_bits] + rnk[r];
    p.keys[pos] = (unsigned)cell[r];
    p.vals[pos] = lo + r * 32 + lane;
  }
#ifdef VX_PROFILE
  VF_MARK();
  if (p.prof && threadIdx.x == 0 && crank == 0) {
    for (int i = 0; i + 1 < nq; ++i) atomicAdd(p.prof + 8 + i, (unsigned long long)(tq[i + 1] - tq[i]));
    atomicAdd(p.prof + 15, 1ull);
  }
#endif
}

// ---------------------------------------------------------------- K6: bins -> output rows
// grid = (n_bins, frames).  Shared memory: slot table [2^L][T] of point indices, running
// per-voxel counts [2^L], per-warp tile counters [8][2^L], (fused) decoration stage.
#define VX_OUT_VOXELS 0     // raw (T,C) voxels
#define VX_OUT_DECORATE 1   // PillarFeatureNet decoration fused into the gather, one warp per pillar
#define VX_OUT_PFN 2        // decoration + PFNLayer (inference) fused into the gather: (rows, units) features
#define VX_OUT_MEAN 3       // SimpleVoxel mean VFE (voxel_encoder.py:219-225) fused into the gather
// FIXED: the Lyft PointPillars shape (T = 60, PillarFeatureNet decoration, 9 channels: rows of 540 floats) as
// compile-time constants - the row loops unroll and the layout tests fold away.
template <int MODE, bool C4, bool FIXED = false>
__global__ void __launch_bounds__(VX_THREADS, MODE == VX_OUT_PFN ? 3 : VX_BINS_MINB) vx_bins_kernel(VoxParams p, DecoCfg d, float* __restrict__ decorated,
                                                                PfnCfg pfn) {
  constexpr bool DECO = MODE == VX_OUT_DECORATE || MODE == VX_OUT_PFN;
  if (FIXED) {
    p.T = 60; d.T = 60; d.C_out = 9; d.variant = LV_PILLAR_PFN; d.with_distance = 0;
  }
  extern __shared__ __align__(16) int smem[];
  const int NV = 1 << p.low_bits;
  int* slot_tab = smem;                    // [NV][T]
  int* total = slot_tab + NV * p.T;        // [NV]
  int* s_cell = total + NV;                // [NV] creator cell of every voxel of the bin
  int* wcnt = s_cell + NV;                 // [8][NV]
  float* stage = reinterpret_cast<float*>(wcnt + VX_WARPS * NV);  // [8][T*C_out] (DECO only)
  __shared__ long long s_row0;

  const int fl = blockIdx.y, f = p.f0 + fl, b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef VX_PROFILE
  const long long tp0 = clock64();
#endif
  const int vnum = p.voxel_num[f];
  if (b == 0 && threadIdx.x < 2 && p.frame_sync) p.frame_sync[2 * fl + threadIdx.x] = 0u;   // fused prologue: counters back to zero
  if ((b << p.low_bits) >= vnum && b != 0) return;  // bin beyond the last voxel: nothing to do
  // first output row of the frame: prefix of voxel_num over the earlier frames of the batch
  if (warp == 0) {
    long long rb;
    if (p.concat) {
      int acc = 0;
      for (int g = p.f0 + lane; g < f; g += 32) acc += p.voxel_num[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      rb = p.row_base[p.f0] + acc;
    } else {
      rb = (long long)f * p.V;
    }
    if (lane == 0) {
      s_row0 = rb;
      if (b == 0 && p.concat) {
        if (f > p.f0) p.row_base[f] = rb;
        if (f + 1 == p.f1) p.row_base[f + 1] = rb + vnum;
      }
    }
  }
  const int v0 = b << p.low_bits;
  int nv = vnum - v0;
  if (nv > NV) nv = NV;
  const int cb = __ldg(p.frame_chunk + f) - p.chunk_lo;
  const int nch = __ldg(p.frame_chunk + f + 1) - __ldg(p.frame_chunk + f);
  if (nv <= 0 || nch == 0) return;  // (uniform across the CTA)
  const int seg_lo = p.bin_start[(int64_t)fl * (p.n_bins + 1) + b], seg_hi = p.bin_start[(int64_t)fl * (p.n_bins + 1) + b + 1];
  const int64_t fstart = __ldg(p.frame_off + f);
  const uint32_t* keys = p.keys + (fstart - p.pt_lo);
  const int32_t* vals = p.vals + (fstart - p.pt_lo);

  for (int i = threadIdx.x; i < NV; i += VX_THREADS) {
    total[i] = 0;
    s_cell[i] = i < nv ? p.creator_cell[(fstart - p.pt_lo) + v0 + i] : 0;   // in flight during phase 1
  }
#ifdef VX_PROFILE
  const long long tp1 = clock64();
#endif
  const unsigned lt = lv_lanemask_lt();
  const unsigned lowmask = NV - 1;
  // ---- phase 1: stable rank by voxel inside the bin, tile by tile (2048 points) ----
  for (int tile0 = seg_lo; tile0 < seg_hi; tile0 += VX_CHUNK) {
    unsigned dig[VX_ITEMS];
    int val[VX_ITEMS], rnk[VX_ITEMS];
    // every warp owns a contiguous span of the tile (point order = warp, round, lane); the
    // span shrinks with the tile so that small bins still keep all 8 warps busy
    int n_tile = seg_hi - tile0;
    if (n_tile > VX_CHUNK) n_tile = VX_CHUNK;
    const int span = (((n_tile + VX_WARPS - 1) / VX_WARPS) + 31) & ~31;
    const int rounds = span >> 5;
    const int base = tile0 + warp * span;
    const int tile_hi = tile0 + n_tile;
    // the tile's keys and values: all loads in flight while the counters are reset
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      const int pos = base + r * 32 + lane;
      const bool in = r < rounds && pos < tile_hi;
      dig[r] = in ? (keys[pos] & lowmask) : 0xffffffffu;
      val[r] = in ? vals[pos] : 0;
    }
    for (int i = threadIdx.x; i < VX_WARPS * NV; i += VX_THREADS) wcnt[i] = 0;
    __syncthreads();
    int* mycnt = wcnt + warp * NV;
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      rnk[r] = 0;
      if (r >= rounds) continue;  // warp-uniform
      const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
      const int leader = __ffs(peers) - 1;
      int old = 0;
      if (dig[r] != 0xffffffffu && lane == leader) {
        old = mycnt[dig[r]];
        mycnt[dig[r]] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rnk[r] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    // warp bases = running per-voxel total + counts of the lower warps; totals advance
    for (int v = threadIdx.x; v < NV; v += VX_THREADS) {
      int off = total[v];
#pragma unroll
      for (int w = 0; w < VX_WARPS; ++w) {
        const int t = wcnt[w * NV + v];
        wcnt[w * NV + v] = off;
        off += t;
      }
      total[v] = off;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < VX_ITEMS; ++r) {
      if (dig[r] == 0xffffffffu) continue;
      const int slot = mycnt[dig[r]] + rnk[r];
      if (slot < p.T) slot_tab[dig[r] * p.T + slot] = val[r];  // first T points in point order
    }
    __syncthreads();
  }
  __syncthreads();
#ifdef VX_PROFILE
  const long long tp2 = clock64();
  struct ProfEnd {
    unsigned long long* prof; long long t0, t1, t2;
    __device__ ~ProfEnd() {
      if (prof && threadIdx.x == 0) {
        const long long t3 = clock64();
        atomicAdd(prof + 0, (unsigned long long)(t1 - t0));
        atomicAdd(prof + 1, (unsigned long long)(t2 - t1));
        atomicAdd(prof + 2, (unsigned long long)(t3 - t2));
        atomicAdd(prof + 3, 1ull);
      }
    }
  } prof_end{p.prof, tp0, tp1, tp2};
#endif
  // ---- phase 2: the output rows of the bin (contiguous in memory) ----
  const long long row0 = s_row0;
  if (row0 + v0 + nv > p.capacity) nv = (int)(p.capacity - (row0 + v0) > 0 ? p.capacity - (row0 + v0) : 0);
  if (MODE == VX_OUT_MEAN) {
    // one thread per voxel: mean of its stored points (the padded slots add zeros), count, coordinates.
    // `decorated` is the (rows, d.C_out) mean matrix, d.C_out <= 4 leading channels.
    const float4* pts4 = reinterpret_cast<const float4*>(p.pts) + fstart;
    for (int v = threadIdx.x; v < nv; v += VX_THREADS) {
      int n = total[v];
      if (n > p.T) n = p.T;
      const long long row = row0 + v0 + v;
      int cx, cy, cz;
      vx_cell_coords(p, s_cell[v], cz, cy, cx);
      p.num_points[row] = n;
      if (p.coord_cols == 4) {
        *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(f, cz, cy, cx);  // preprocess.py:44-50
      } else {
        int32_t* co = p.coords + row * 3;
        co[0] = cz; co[1] = cy; co[2] = cx;
      }
      const int* sl = slot_tab + v * p.T;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = 0; t < n; ++t) {
        const float4 q = __ldg(pts4 + sl[t]);
        s4.x += q.x; s4.y += q.y; s4.z += q.z; s4.w += q.w;
      }
      const float fn = (float)n;
      const float m4[4] = {__fdiv_rn(s4.x, fn), __fdiv_rn(s4.y, fn), __fdiv_rn(s4.z, fn), __fdiv_rn(s4.w, fn)};
      float* o = decorated + row * d.C_out;
      for (int c = 0; c < d.C_out; ++c) o[c] = m4[c];
    }
    return;
  }
  if (MODE == VX_OUT_VOXELS && C4) {
    // 2a: one thread per voxel: count and coordinates
    const float4* pts4 = reinterpret_cast<const float4*>(p.pts) + fstart;
    for (int v = threadIdx.x; v < nv; v += VX_THREADS) {
      int n = total[v];
      if (n > p.T) n = p.T;
      total[v] = n;
      const long long row = row0 + v0 + v;
      int cx, cy, cz;
      vx_cell_coords(p, s_cell[v], cz, cy, cx);
      p.num_points[row] = n;
      if (p.coord_cols == 4) {
        *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(f, cz, cy, cx);  // preprocess.py:44-50
      } else {
        int32_t* co = p.coords + row * 3;
        co[0] = cz; co[1] = cy; co[2] = cx;  // reversed (z,y,x), simplevis.py:42
      }
    }
    __syncthreads();
    // 2b: the whole CTA streams the rows of the bin as one flat array of float4 (one per slot):
    //     live slots gather their point, padding slots are zeros from registers
    float4* out4 = reinterpret_cast<float4*>(p.voxels) + (row0 + v0) * p.T;
    const int total4 = nv * p.T;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = threadIdx.x; j < total4; j += VX_THREADS) {
      const int v = p.T == 1 ? j : (int)(((unsigned long long)(unsigned)j * p.row_div_m) >> 32);   // j / T (j < 2^20; ceil(2^32 / 1) does not fit)
      const int r = j - v * p.T;
      float4 o = z4;
      if (r < total[v]) o = __ldg(pts4 + slot_tab[v * p.T + r]);
      lv_st_stream_f4(out4 + j, o);
    }
    return;
  }
  if (DECO) {
    const int per = d.T * d.C_out;
    // per-warp stage: the dense (T, C_out) row (+ pad), or T slots at stride LV_PFN_STRIDE
    float* st = stage + warp * (MODE == VX_OUT_PFN ? d.T * LV_PFN_STRIDE : per + 4);
    const float4* pts4 = reinterpret_cast<const float4*>(p.pts) + fstart;
    PfnRegs<9, 2> pfn_regs;   // the fused path serves the PointPillars PFN: 9 inputs, 64 units
    if (MODE == VX_OUT_PFN) pfn_regs.load(pfn, lane);
    // two pillars per iteration: the point gathers of both are in flight, and the zero part
    // of both rows is already streaming out, before either gather is consumed
    // a warp takes GROUP consecutive pillars at a time, pair by pair.  Measured (pillarize stage / fused-PFN pillar
    // path, ms): GROUP 2: 0.659 / 1.622, GROUP 4: 0.663 / 1.605 - four consecutive 256-byte feature rows make a
    // full 1 KB run, the 2,160-byte decorated rows gain nothing.  (Four pillars of <= 8 points decorated by one warp
    // at once - 8 lanes each - measured slower still: 0.668 ms.)
    constexpr int GROUP = MODE == VX_OUT_PFN ? 4 : 2;
    for (int vq = warp * GROUP; vq < nv; vq += VX_WARPS * GROUP) {
     for (int v = vq; v < vq + GROUP && v < nv; v += 2) {
      const long long row = row0 + v0 + v;
      if (row >= p.capacity) break;
      const bool two = (v + 1 < nv) && (row + 1 < p.capacity);
      int n0 = total[v], n1 = two ? total[v + 1] : 0;
      if (n0 > p.T) n0 = p.T;
      if (n1 > p.T) n1 = p.T;
      const int* sl = slot_tab + v * p.T;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (two && n0 <= 16 && n1 <= 16 && p.T >= 32) {
        // two small pillars (the common case: 5.4 points per pillar on a Lyft sweep) share the warp.
        // T >= 32: the second half's stage starts 16 slots into the warp's T-slot stage and holds up to 16 more
        const int hi = lane >> 4, sub = lane & 15;
        const int nm = hi ? n1 : n0;
        float4 a = z4;
        if (sub < nm) a = __ldg(pts4 + sl[hi * p.T + sub]);
        const int cm = s_cell[v + hi];
        float* dst0 = decorated + row * per;
        if (MODE == VX_OUT_DECORATE) {
          lv_decorate_zero_tail(n0, d, dst0, lane);
          lv_decorate_zero_tail(n1, d, dst0 + per, lane);
        }
        int cx, cy, cz;
        vx_cell_coords(p, cm, cz, cy, cx);
        if (sub == 0) {
          p.num_points[row + hi] = nm;
          *reinterpret_cast<int4*>(p.coords + (row + hi) * 4) = make_int4(f, cz, cy, cx);  // preprocess.py:44-50
        }
        if (MODE == VX_OUT_DECORATE) {
          lv_decorate_half(a, nm, cy, cx, d, st + hi * (16 * d.C_out + 4), dst0 + hi * per, lane);
        } else {  // VX_OUT_PFN: both pillars staged at once, then the PFN over each
          lv_decorate_half_stage(a, nm, cy, cx, d, st + hi * (16 * LV_PFN_STRIDE), lane, LV_PFN_STRIDE);
          float* f0 = decorated + row * pfn.units;
          lv_pfn_warp<9, 2>(st, n0, d.T, pfn_regs, f0, lane);
          lv_pfn_warp<9, 2>(st + 16 * LV_PFN_STRIDE, n1, d.T, pfn_regs, f0 + pfn.units, lane);
        }
        continue;
      }
      float4 a0 = z4, b0 = z4, a1 = z4, b1 = z4;
      if (lane < n0) a0 = __ldg(pts4 + sl[lane]);
      if (lane + 32 < n0) b0 = __ldg(pts4 + sl[lane + 32]);
      if (lane < n1) a1 = __ldg(pts4 + sl[p.T + lane]);
      if (lane + 32 < n1) b1 = __ldg(pts4 + sl[p.T + lane + 32]);
      const int c0 = s_cell[v], c1 = two ? s_cell[v + 1] : 0;
      float* dst0 = decorated + row * per;
      if (MODE == VX_OUT_DECORATE) {
        lv_decorate_zero_tail(n0, d, dst0, lane);
        if (two) lv_decorate_zero_tail(n1, d, dst0 + per, lane);
      }
      int cx0, cy0, cz0, cx1, cy1, cz1;
      vx_cell_coords(p, c0, cz0, cy0, cx0);
      vx_cell_coords(p, c1, cz1, cy1, cx1);
      if (lane == 0) {
        p.num_points[row] = n0;
        *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(f, cz0, cy0, cx0);  // preprocess.py:44-50
        if (two) {
          p.num_points[row + 1] = n1;
          *reinterpret_cast<int4*>(p.coords + (row + 1) * 4) = make_int4(f, cz1, cy1, cx1);
        }
      }
      if (MODE == VX_OUT_DECORATE) {
        lv_decorate_warp(a0, b0, n0, cy0, cx0, d, st, dst0, lane);
        if (two) lv_decorate_warp(a1, b1, n1, cy1, cx1, d, st, dst0 + per, lane);
      } else {  // VX_OUT_PFN: `decorated` is the (rows, units) feature matrix
        float* f0 = decorated + row * pfn.units;
        int live = lv_decorate_stage(a0, b0, n0, cy0, cx0, d, st, lane, LV_PFN_STRIDE);
        lv_pfn_warp<9, 2>(st, live, d.T, pfn_regs, f0, lane);
        if (two) {
          live = lv_decorate_stage(a1, b1, n1, cy1, cx1, d, st, lane, LV_PFN_STRIDE);
          lv_pfn_warp<9, 2>(st, live, d.T, pfn_regs, f0 + pfn.units, lane);
        }
      }
     }
    }
  } else {
    // LPV lanes per voxel: 32 for pillars, 8 for T <= 8
    const int LPV = p.T > 16 ? 32 : (p.T > 8 ? 16 : 8);
    const int groups = VX_THREADS / LPV;
    const int sub = threadIdx.x % LPV;
    for (int v = threadIdx.x / LPV; v < nv; v += groups) {
      const long long row = row0 + v0 + v;
      if (row >= p.capacity) break;
      int n = total[v];
      if (n > p.T) n = p.T;
      const int* sl = slot_tab + v * p.T;
      if (sub == 0) {
        p.num_points[row] = n;
        int cx, cy, cz;
        vx_cell_coords(p, s_cell[v], cz, cy, cx);
        if (p.coord_cols == 4) {
          *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(f, cz, cy, cx);
        } else {
          int32_t* co = p.coords + row * 3;
          co[0] = cz; co[1] = cy; co[2] = cx;  // reversed (z,y,x), simplevis.py:42
        }
      }
      float* out = p.voxels + row * p.T * p.C;
      if (C4) {
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int t = sub; t < p.T; t += LPV) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t < n) q = __ldg(reinterpret_cast<const float4*>(p.pts) + fstart + sl[t]);
          lv_st_stream_f4(o4 + t, q);
        }
      } else {
        for (int t = sub; t < p.T; t += LPV) {
          const float* src = (t < n) ? p.pts + (fstart + sl[t]) * p.C : nullptr;
          for (int cc = 0; cc < p.C; ++cc) out[t * p.C + cc] = src ? __ldg(src + cc) : 0.f;
        }
      }
    }
  }
}

#include "lv_voxel_list.cuh"

// rows [voxel_num, V) of the padded layout (generate_multi_gpu, preprocess.py:311-317)
__global__ void __launch_bounds__(VX_THREADS) vx_zero_tail_kernel(VoxParams p) {
  const int f = p.f0 + blockIdx.y;
  const int vnum = p.voxel_num[f];
  const int per = p.T * p.C;
  for (int v = vnum + blockIdx.x; v < p.V; v += gridDim.x) {
    const int64_t row = (int64_t)f * p.V + v;
    float* out = p.voxels + row * per;
    for (int i = threadIdx.x; i < per; i += VX_THREADS) out[i] = 0.f;
    if (threadIdx.x < 3) p.coords[row * 3 + threadIdx.x] = 0;
    if (threadIdx.x == 3) p.num_points[row] = 0;
  }
}

// ---------------------------------------------------------------- multi-GPU batch un-padding
// VoxelNet.forward (second/second/pytorch/models/voxelnet.py:346-358): a padded batch
// (B,V,T,C) / (B,V) / (B,V,4) + num_voxels (B) becomes the concatenated lists the network
// consumes.  grid = (row blocks, B); every CTA derives its sample's first output row from the
// counts of the samples before it.  Rows are copied as float4 when T*C % 4 == 0.
__global__ void __launch_bounds__(256) vx_unpad_kernel(const float* __restrict__ voxels, const int32_t* __restrict__ num_points,
                                                      const int32_t* __restrict__ coors, const int32_t* __restrict__ num_voxels,
                                                      int B, int V, int per, int coor_cols, float* __restrict__ out_voxels,
                                                      int32_t* __restrict__ out_num, int32_t* __restrict__ out_coors,
                                                      int64_t* __restrict__ out_total) {
  __shared__ long long s_base;
  const int b = blockIdx.y;
  if (threadIdx.x < 32) {
    long long acc = 0;
    for (int g = threadIdx.x; g < b; g += 32) {
      int nv = num_voxels[g];
      acc += nv < 0 ? 0 : (nv > V ? V : nv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) s_base = acc;
  }
  __syncthreads();
  int nv = num_voxels[b];
  nv = nv < 0 ? 0 : (nv > V ? V : nv);
  const long long base = s_base;
  if (b == B - 1 && blockIdx.x == 0 && threadIdx.x == 0) *out_total = base + nv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (per & 3) == 0 && ((reinterpret_cast<uintptr_t>(voxels) | reinterpret_cast<uintptr_t>(out_voxels)) & 15) == 0;
  for (int r = blockIdx.x * 8 + warp; r < nv; r += gridDim.x * 8) {
    const long long src = (long long)b * V + r, dst = base + r;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(voxels + src * per);
      float4* d4 = reinterpret_cast<float4*>(out_voxels + dst * per);
      for (int i = lane; i < per / 4; i += 32) lv_st_stream_f4(d4 + i, lv_ld_stream_f4(s4 + i));
    } else {
      for (int i = lane; i < per; i += 32) out_voxels[dst * per + i] = voxels[src * per + i];
    }
    if (lane == 0) out_num[dst] = num_points[src];
    if (lane < coor_cols) out_coors[dst * coor_cols + lane] = coors[src * coor_cols + lane];
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

static void vx_make_div(uint32_t d, unsigned* m, int* s) {
  int lg = 0;
  while ((1ull << lg) < d) ++lg;
  *s = 28 + lg;
  *m = (unsigned)(((1ull << *s) + d - 1) / d);
}

template <typename K>
static int vx_set_smem(K kernel, size_t bytes) {
  if (bytes > 40 * 1024)  // dynamic + the kernels' few static bytes must stay under the 48 KB default
    LV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return LV_OK;
}

static int vx_run(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                  const int64_t* h_frame_offsets, float* d_voxels, int32_t* d_coords, int32_t* d_num_points,
                  int32_t* d_voxel_num, int concat, int64_t capacity, int64_t* d_row_base, const DecoCfg* deco,
                  float* d_decorated, const PfnCfg* pfn, lv_stream stream_, int mean_channels = 0) {
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
  const int64_t n_cells = (int64_t)grid[0] * grid[1] * grid[2];
  LV_REQUIRE(n_cells < (1ll << 28), "lv_voxelize: grid of %lld cells exceeds the cell-id limit (2^28)", (long long)n_cells);
  int64_t G = n_cells;   // int32 words of first[] per frame (replaced by the hash-table size below)
  LV_REQUIRE(n_frames == 0 || h_frame_offsets[0] == 0, "lv_voxelize: frame_offsets[0] must be 0");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE((d_voxels || deco || mean_channels) && d_coords && d_num_points && d_voxel_num, "lv_voxelize: null output");
  LV_REQUIRE(mean_channels == 0 || (d_decorated && cfg->num_features == 4 && mean_channels <= 4 &&
                                    (reinterpret_cast<uintptr_t>(d_points) & 15) == 0),
             "lv_voxelize_mean_concat: needs 4 features per point, 1..4 output channels and 16-byte aligned points");
  LV_REQUIRE(!deco || (d_decorated && concat && cfg->num_features == 4 && cfg->max_points <= 64 &&
                       (reinterpret_cast<uintptr_t>(d_points) & 15) == 0),
             "lv_pillarize_concat: needs 4 features per point, max_points <= 64 and 16-byte aligned points");
  const int V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features;

  // bins: 2^L voxel ids each, L as large as a 32 KB slot table allows
  int L = VX_MAX_LOW_BITS;
  while (L > 0 && ((size_t)1 << L) * T * 4 > 32 * 1024) --L;
  while ((((int64_t)V + (1 << L) - 1) >> L) > VX_MAX_BINS && L < 20) ++L;
  // few frames per call (a single ten-sweep cloud): one CTA per (frame, bin) leaves most SMs idle in K6, so the bins shrink
  // until the call has ~4 CTAs per SM or 512 bins per frame (every bin costs K3-K5 a counter per chunk)
  if (h->vox_small_bins >= 0)
    while (L > 5 && (int64_t)n_frames * (((int64_t)V + (1 << L) - 1) >> L) < 4ll * h->num_sms &&
           (((int64_t)V + (1 << (L - 1)) - 1) >> (L - 1)) <= 512)
      --L;
  const int n_bins = (int)(((int64_t)V + (1 << L) - 1) >> L);
  const int NV = 1 << L;
  const size_t deco_stage = pfn ? (size_t)VX_WARPS * T * LV_PFN_STRIDE * 4
                                 : (deco ? (size_t)VX_WARPS * (T * deco->C_out + 4) * 4 : 0);
  const size_t smem_bins = ((size_t)NV * T + 2 * NV + (size_t)VX_WARPS * NV) * 4 + deco_stage;
  LV_REQUIRE(smem_bins <= 220 * 1024, "lv_voxelize: max_points %d x max_voxels %d needs %zu bytes of shared memory "
             "per bin (limit 220 KB)", T, V, smem_bins);
  LV_REQUIRE((int64_t)NV * T < (1ll << 20), "lv_voxelize: bin of %d voxels x %d points is too large", NV, T);
  const size_t smem_scatter = (size_t)VX_WARPS * n_bins * 4;
  const size_t smem_keys = (size_t)n_bins * 4;

  // chunk prefix per frame (a frame with no points owns zero chunks)
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
  std::vector<int32_t> chunk_frame((size_t)frame_chunk[n_frames] + 1, 0);
  for (int f = 0; f < n_frames; ++f)
    for (int32_t c = frame_chunk[f]; c < frame_chunk[f + 1]; ++c) chunk_frame[c] = f;
  const void* d_chunk_frame = nullptr;
  LV_CHECK(h->vox_chunk_frame.sync(chunk_frame.data(), sizeof(int32_t) * chunk_frame.size(), stream, &d_chunk_frame));

  // Grids whose dense map cannot stay in L2 (> 48 MB per frame; the 0.05 m SECOND grid needs 216 MB) get an
  // open-addressing table sized by the POINTS instead: 2^b >= 1.25 x the largest frame, 8 bytes per slot - 16 MB
  // for a 1.06 M-point cloud, so that six clouds share one launch with every atomic in L2.
  // vox_hash_map: 0 = automatic (such grids, three or more frames per call), 1 = always, -1 = never.
  int64_t max_frame_pts0 = 0;
  for (int f = 0; f < n_frames; ++f)
    if (h_frame_offsets[f + 1] - h_frame_offsets[f] > max_frame_pts0) max_frame_pts0 = h_frame_offsets[f + 1] - h_frame_offsets[f];
  int hash_bits = 0;
  if (h->vox_hash_map >= 0) {
    int b = 10;
    while ((1ll << b) < max_frame_pts0 + max_frame_pts0 / 4 + 1) ++b;
    // measured on 1.06 M-point clouds at the 0.05 m SECOND grid (dense / table, ms per cloud): one cloud 0.120 / 0.152, two
    // 0.101 / 0.104, six 0.090 / 0.075, twelve 0.088 / 0.073 - the table pays from three clouds per call on
    if (b <= 28 && (h->vox_hash_map == 1 || (n_frames >= 3 && n_cells * 4 > (48ll << 20) && (8ll << b) < n_cells * 4))) hash_bits = b;
  }
  if (hash_bits) G = 2ll << hash_bits;
  // sub-batches: bounded by the dense map budget (L2 residency) and by the point workspace,
  // then balanced so that no launch is a small remainder
  // 96 MB of the 126 MB L2 measured best (bench.py: 64 MB 0.94 ms, 96 MB 0.90 ms, 128 MB 0.88 ms per
  // 128 pillar frames); beyond the L2 the random atomicMin traffic would go to HBM
  const int64_t map_budget = h->vox_dense_map_limit_bytes > 0 ? h->vox_dense_map_limit_bytes : (96ll << 20);
  int64_t fif = map_budget / (G * 4);
  if (fif < 1) fif = 1;
  if (fif > n_frames) fif = n_frames;
  fif = lv_div_up(n_frames, lv_div_up(n_frames, fif));
  const int64_t max_pts = 8ll << 20;

  VoxParams p;
  memset(&p, 0, sizeof(p));
  p.pts = d_points; p.C = C;
  p.frame_off = (const int64_t*)d_off;
  p.frame_chunk = (const int32_t*)d_chunk;
  p.chunk_frame = (const int32_t*)d_chunk_frame;
  for (int j = 0; j < 3; ++j) {
    p.lo[j] = cfg->coors_range[j];
    p.vs[j] = cfg->voxel_size[j];
    p.grid[j] = grid[j];
  }
  vx_make_div((uint32_t)grid[0] * (uint32_t)grid[1], &p.div_m[0], &p.div_s[0]);
  vx_make_div((uint32_t)grid[0], &p.div_m[1], &p.div_s[1]);
  p.hash_bits = hash_bits; p.map_shift = hash_bits ? 1 : 0;
  p.G = G; p.T = T; p.V = V; p.overflow = cfg->overflow_mode; p.zero_tail = cfg->zero_tail;
  p.low_bits = L; p.n_bins = n_bins;
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
  DecoCfg dcfg;
  memset(&dcfg, 0, sizeof(dcfg));
  if (deco) dcfg = *deco;
  PfnCfg pcfg;
  memset(&pcfg, 0, sizeof(pcfg));
  if (pfn) pcfg = *pfn;

  // K1 stages full chunks by TMA when a chunk of rows is a multiple of 16 bytes and fits 64 KB
  p.tma_bytes = 0;
  if (!h->disable_tma && C <= 8 && ((size_t)VX_CHUNK * C * 4) % 16 == 0) p.tma_bytes = (unsigned)(VX_CHUNK * C * 4);
  if (c4 && hash_bits) LV_CHECK(vx_set_smem(vx_cells_kernel<true, true>, p.tma_bytes));
  else if (c4) LV_CHECK(vx_set_smem(vx_cells_kernel<true, false>, p.tma_bytes));
  else if (hash_bits) LV_CHECK(vx_set_smem(vx_cells_kernel<false, true>, p.tma_bytes));
  else LV_CHECK(vx_set_smem(vx_cells_kernel<false, false>, p.tma_bytes));
  LV_CHECK(vx_set_smem(vx_keys_kernel, smem_keys));
  LV_CHECK(vx_set_smem(vx_scatter_kernel, smem_scatter));
  if (mean_channels) {
    dcfg.C_out = mean_channels;
    dcfg.T = T;
    LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_MEAN, true>, smem_bins));
  } else if (pfn) {
    LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_PFN, true>, smem_bins));
    LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_PFN, true, true>, smem_bins));
  } else if (deco) {
    LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_DECORATE, true>, smem_bins));
    LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_DECORATE, true, true>, smem_bins));
  } else if (out4) LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_VOXELS, true>, smem_bins));
  else LV_CHECK(vx_set_smem(vx_bins_kernel<VX_OUT_VOXELS, false>, smem_bins));
  p.row_div_m = (unsigned)(((1ull << 32) + (uint32_t)T - 1) / (uint32_t)T);
#ifdef VX_PROFILE
  static unsigned long long* d_prof = nullptr;
  if (!d_prof) { cudaMalloc(&d_prof, 128); cudaMemset(d_prof, 0, 128); }
  p.prof = d_prof;
  {
    unsigned long long hp[16];
    cudaMemcpy(hp, d_prof, 128, cudaMemcpyDeviceToHost);
    if (hp[15]) fprintf(stderr, "[vx_frame profile] clusters %llu  A %.0f  B %.0f  C %.0f  D %.0f  E %.0f  F %.0f  G %.0f clocks\n", hp[15],
                        (double)hp[8] / hp[15], (double)hp[9] / hp[15], (double)hp[10] / hp[15], (double)hp[11] / hp[15],
                        (double)hp[12] / hp[15], (double)hp[13] / hp[15], (double)hp[14] / hp[15]);
    if (hp[3]) fprintf(stderr, "[vx_bins profile] CTAs %llu  prologue %.0f  phase1 %.0f  phase2 %.0f clocks per CTA (thread 0)\n", hp[3],
                       (double)hp[0] / hp[3], (double)hp[1] / hp[3], (double)hp[2] / hp[3]);
    cudaMemset(d_prof, 0, 128);
  }
#endif

  // cluster-per-frame prologue (vx_frame_kernel): the smallest cluster whose shared memory holds the frame's map
  int frame_cs = 0, frame_part = 0;
  size_t frame_smem = 0;
  int64_t max_frame_pts = 0;
  for (int f = 0; f < n_frames; ++f)
    if (h_frame_offsets[f + 1] - h_frame_offsets[f] > max_frame_pts) max_frame_pts = h_frame_offsets[f + 1] - h_frame_offsets[f];
  if (h->vox_frame_kernel != 0 && n_bins <= 512 && !hash_bits) {
    for (int cs = 1; cs <= VF_MAX_CS; cs *= 2) {
      const int64_t part = lv_div_up(G, cs);
      const size_t bytes = ((size_t)part + (size_t)(VF_WARPS + VF_MAX_CS + 1) * n_bins + VF_MAX_CS * VF_WARPS + 8) * 4;
      if (bytes <= 226 * 1024 && max_frame_pts <= (int64_t)cs * VF_WARPS * VF_ROUNDS * 32) {
        frame_cs = cs; frame_part = (int)part; frame_smem = bytes;
        break;
      }
    }
    if (frame_cs > 0) {
      if (c4) LV_CHECK_CUDA(cudaFuncSetAttribute(vx_frame_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frame_smem));
      else LV_CHECK_CUDA(cudaFuncSetAttribute(vx_frame_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frame_smem));
    }
  }

  // list path (lv_voxel_list.cuh): three kernels, order rebuilt inside the warp that writes the row.  Small dense
  // maps only (8 bytes per cell), 4-float points, rows of at most 64 slots.  OPT-IN (lv_set_option "vox_list_path" 1):
  // bit-identical, but measured slower on 128 C5 frames (see the header of lv_voxel_list.cuh).
  const bool list_path = h->vox_list_path == 1 && frame_cs == 0 && !hash_bits && !mean_channels && c4 && T <= 64 && (deco || out4) &&
                         G * 8 <= (64ll << 20) && max_frame_pts < (1ll << VL_RANK_BITS);
  size_t smem_rows = 0;
  if (list_path) {
    const int64_t budget2 = h->vox_dense_map_limit_bytes > 0 ? h->vox_dense_map_limit_bytes : (192ll << 20);
    fif = budget2 / (G * 8);
    if (fif < 1) fif = 1;
    if (fif > 2048) fif = 2048;
    if (fif > n_frames) fif = n_frames;
    fif = lv_div_up(n_frames, lv_div_up(n_frames, fif));
    smem_rows = ((size_t)((2 * (fif + 1) + 3) & ~3) + VX_WARPS * 64) * 4 + deco_stage;
    if (pfn) LV_CHECK(vx_set_smem(vl_rows_kernel<VX_OUT_PFN>, smem_rows));
    else if (deco) LV_CHECK(vx_set_smem(vl_rows_kernel<VX_OUT_DECORATE>, smem_rows));
    else LV_CHECK(vx_set_smem(vl_rows_kernel<VX_OUT_VOXELS>, smem_rows));
    LV_CHECK(vx_set_smem(vl_cells_kernel<true>, p.tma_bytes));
  }

  const bool fused_ok = h->vox_fused_prologue != 0 && !list_path && !hash_bits && n_bins <= VX_FUSED_MAX_BINS;
  int f0 = 0;
  while (f0 < n_frames) {
    int f1 = f0 + 1;
    while (f1 < n_frames && f1 - f0 < fif && h_frame_offsets[f1 + 1] - h_frame_offsets[f0] <= max_pts) ++f1;
    const int nf = f1 - f0;
    const int64_t pt_lo = h_frame_offsets[f0], npts = h_frame_offsets[f1] - pt_lo;
    const int chunk_lo = frame_chunk[f0], nchunks = frame_chunk[f1] - chunk_lo;
    p.f0 = f0; p.f1 = f1; p.chunk_lo = chunk_lo; p.pt_lo = pt_lo;
    {   // frames with many chunks: the [bin][chunk] table of one frame is too long for one CTA to scan
      int max_ch = 0;
      for (int f = f0; f < f1; ++f)
        if (frame_chunk[f + 1] - frame_chunk[f] > max_ch) max_ch = frame_chunk[f + 1] - frame_chunk[f];
      p.two_level_scan = (int64_t)max_ch * n_bins > 16384 && h->vox_two_level_scan >= 0 ? 1 : 0;
      if (h->vox_two_level_scan == 1) p.two_level_scan = 1;
    }

    if (list_path) {
      LV_CHECK(h->vox_map.ensure((size_t)nf * G * 8, stream, 0x7f));
      LV_CHECK(h->vox_cell.ensure((size_t)(npts + 1) * 4 * 2, stream));            // cell + arrival rank
      LV_CHECK(h->vox_vals[0].ensure((size_t)(npts + 1) * 4, stream));             // lists
      LV_CHECK(h->vox_vrec.ensure((size_t)(npts + 1) * 16, stream));               // per-voxel records
      LV_CHECK(h->vox_chunk.ensure((size_t)(nchunks + 1) * 8, stream));            // look-back descriptors
      LV_CHECK(h->vox_frame_state.ensure((size_t)nf * 2 * 4, stream));
      p.map = h->vox_map.as<int32_t>();
      p.cell = h->vox_cell.as<int32_t>();
      p.chunk_state = h->vox_chunk.as<unsigned long long>();
      VlParams q;
      q.map2 = h->vox_map.as<unsigned long long>();
      q.arr = h->vox_cell.as<uint32_t>() + (npts + 1);
      q.list = h->vox_vals[0].as<int32_t>();
      q.vrec = h->vox_vrec.as<int4>();
      q.frame_total = h->vox_frame_state.as<int32_t>();
      if (nchunks > 0) {
        vl_cells_kernel<true><<<nchunks, VX_THREADS, p.tma_bytes, stream>>>(p, q);
        LV_LAUNCH_CHECK(h);
        vl_assign_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p, q);
        LV_LAUNCH_CHECK(h);
      }
      // frames without points never ran L2: their totals are set here
      bool any_empty = false;
      for (int f = f0; f < f1; ++f) any_empty |= frame_chunk[f + 1] == frame_chunk[f];
      if (any_empty || nchunks == 0) {
        vl_empty_frames_kernel<<<(unsigned)lv_div_up(nf, 256), 256, 0, stream>>>(p, q);
        LV_LAUNCH_CHECK(h);
      }
      const int minb = pfn ? 3 : VX_BINS_MINB;
      int64_t grid_r = (int64_t)h->num_sms * minb * (h->vox_rows_waves > 0 ? h->vox_rows_waves : 4);
      const int64_t max_blocks = lv_div_up(npts < (int64_t)nf * G ? npts : (int64_t)nf * G, VL_ITEMS_PER_CTA);
      if (grid_r > max_blocks) grid_r = max_blocks;
      if (grid_r < 1) grid_r = 1;
      if (pfn) vl_rows_kernel<VX_OUT_PFN><<<(unsigned)grid_r, VX_THREADS, smem_rows, stream>>>(p, q, dcfg, d_decorated, pcfg);
      else if (deco) vl_rows_kernel<VX_OUT_DECORATE><<<(unsigned)grid_r, VX_THREADS, smem_rows, stream>>>(p, q, dcfg, d_decorated, pcfg);
      else vl_rows_kernel<VX_OUT_VOXELS><<<(unsigned)grid_r, VX_THREADS, smem_rows, stream>>>(p, q, dcfg, nullptr, pcfg);
      LV_LAUNCH_CHECK(h);
      if (!concat && cfg->zero_tail) {
        int gx = (int)lv_div_up((int64_t)h->num_sms * 8, nf);
        if (gx > V) gx = V;
        vx_zero_tail_kernel<<<dim3((unsigned)gx, (unsigned)nf), VX_THREADS, 0, stream>>>(p);
        LV_LAUNCH_CHECK(h);
      }
      f0 = f1;
      continue;
    }
    if (frame_cs == 0) LV_CHECK(h->vox_map.ensure((size_t)nf * G * 4, stream, 0x7f));
    LV_CHECK(h->vox_cell.ensure((size_t)(npts + 1) * 4 * 2, stream));            // cell + key0
    LV_CHECK(h->vox_keys[0].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_vals[0].ensure((size_t)(npts + 1) * 4, stream));
    LV_CHECK(h->vox_aux.ensure((size_t)(npts + 1) * 4, stream));                 // creator_cell
    LV_CHECK(h->vox_chunk.ensure((size_t)(nchunks + 1) * 8, stream));            // look-back descriptors
    LV_CHECK(h->vox_hist.ensure((size_t)(nchunks + 1) * n_bins * 4, stream));
    LV_CHECK(h->vox_frame_state.ensure((size_t)nf * 2 * 4, stream));
    LV_CHECK(h->vox_bin_start.ensure((size_t)nf * (n_bins + 1) * 4, stream));
    p.bin_start = h->vox_bin_start.as<int32_t>();
    p.frame_sync = nullptr;
    // fused prologue: every CTA of a frame must be resident at once, its table must fit the CTA's shared memory
    bool fused = false;
    if (fused_ok && frame_cs == 0 && nchunks > 0) {
      int max_ch = 0, min_ch = 1 << 30;
      for (int f = f0; f < f1; ++f) {
        const int ch = frame_chunk[f + 1] - frame_chunk[f];
        if (ch > max_ch) max_ch = ch;
        if (ch < min_ch) min_ch = ch;
      }
      size_t smem_fused = ((size_t)(VX_WARPS + 1 + max_ch) * n_bins) * 4;
      if (smem_fused < p.tma_bytes) smem_fused = p.tma_bytes;
      if (min_ch > 0 && max_ch <= VX_FUSED_MAX_CHUNKS && smem_fused <= 64 * 1024) {
        int per_sm = 0;
        if (c4) {
          LV_CHECK(vx_set_smem(vx_fused_kernel<true>, smem_fused));
          LV_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vx_fused_kernel<true>, VX_THREADS, smem_fused));
        } else {
          LV_CHECK(vx_set_smem(vx_fused_kernel<false>, smem_fused));
          LV_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vx_fused_kernel<false>, VX_THREADS, smem_fused));
        }
        if ((int64_t)per_sm * h->num_sms >= 2 * (int64_t)max_ch) {   // a factor two of slack for co-running kernels
          fused = true;
          // [chunks] look-back descriptors | [frames][2] arrival counters | error flag: all zero between calls
          const size_t st_bytes = lv_align_up((size_t)(nchunks + 1) * 8, 16);
          LV_CHECK(h->vox_fused.ensure(st_bytes + (size_t)nf * 8 + 16, stream, 0));
          p.map = h->vox_map.as<int32_t>();
          p.keys = h->vox_keys[0].as<uint32_t>();
          p.vals = h->vox_vals[0].as<int32_t>();
          p.creator_cell = h->vox_aux.as<int32_t>();
          p.hist = h->vox_hist.as<int32_t>();
          p.chunk_state = h->vox_fused.as<unsigned long long>();
          p.frame_sync = reinterpret_cast<unsigned*>(h->vox_fused.as<unsigned char>() + st_bytes);
          p.err = reinterpret_cast<int*>(p.frame_sync + (size_t)nf * 2);
          p.frame_cut = h->vox_frame_state.as<int32_t>();
          p.frame_kept = p.frame_cut + nf;
          if (c4) vx_fused_kernel<true><<<nchunks, VX_THREADS, smem_fused, stream>>>(p);
          else vx_fused_kernel<false><<<nchunks, VX_THREADS, smem_fused, stream>>>(p);
          LV_LAUNCH_CHECK(h);
        }
      }
    }
    p.map = h->vox_map.as<int32_t>();
    p.cell = h->vox_cell.as<int32_t>();
    p.key0 = h->vox_cell.as<uint32_t>() + (npts + 1);
    p.keys = h->vox_keys[0].as<uint32_t>();
    p.vals = h->vox_vals[0].as<int32_t>();
    p.creator_cell = h->vox_aux.as<int32_t>();
    p.chunk_state = h->vox_chunk.as<unsigned long long>();
    p.hist = h->vox_hist.as<int32_t>();
    p.frame_cut = h->vox_frame_state.as<int32_t>();
    p.frame_kept = p.frame_cut + nf;

    if (fused) {
      // K1-K5 ran as one kernel above
    } else if (frame_cs > 0) {
      // small grid: K1-K5 of every frame inside one thread-block cluster, the dense map in distributed shared memory
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3((unsigned)(nf * frame_cs));
      lc.blockDim = dim3(VF_THREADS);
      lc.dynamicSmemBytes = frame_smem;
      lc.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)frame_cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      if (c4) LV_CHECK_CUDA(cudaLaunchKernelEx(&lc, vx_frame_kernel<true>, p, frame_part));
      else LV_CHECK_CUDA(cudaLaunchKernelEx(&lc, vx_frame_kernel<false>, p, frame_part));
      LV_LAUNCH_CHECK(h);
    } else {
      if (nchunks > 0) {
        if (c4 && hash_bits) vx_cells_kernel<true, true><<<nchunks, VX_THREADS, p.tma_bytes, stream>>>(p);
        else if (c4) vx_cells_kernel<true, false><<<nchunks, VX_THREADS, p.tma_bytes, stream>>>(p);
        else if (hash_bits) vx_cells_kernel<false, true><<<nchunks, VX_THREADS, p.tma_bytes, stream>>>(p);
        else vx_cells_kernel<false, false><<<nchunks, VX_THREADS, p.tma_bytes, stream>>>(p);
        LV_LAUNCH_CHECK(h);
        vx_assign_kernel<<<nchunks, VX_THREADS, 0, stream>>>(p);
        LV_LAUNCH_CHECK(h);
        vx_keys_kernel<<<nchunks, VX_THREADS, smem_keys, stream>>>(p);
        LV_LAUNCH_CHECK(h);
      }
      if (p.two_level_scan) {
        vx_scan_rows_kernel<<<dim3((unsigned)n_bins, (unsigned)nf), VX_THREADS, 0, stream>>>(p);
        LV_LAUNCH_CHECK(h);
        vx_scan_bins_kernel<<<nf, VX_THREADS, 0, stream>>>(p);
      } else {
        vx_scan_hist_kernel<<<nf, VX_THREADS, 0, stream>>>(p);
      }
      LV_LAUNCH_CHECK(h);
      if (nchunks > 0) {
        vx_scatter_kernel<<<nchunks, VX_THREADS, smem_scatter, stream>>>(p);
        LV_LAUNCH_CHECK(h);
      }
    }
    {
      dim3 grid_b((unsigned)n_bins, (unsigned)nf);
      if (mean_channels) vx_bins_kernel<VX_OUT_MEAN, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, d_decorated, pcfg);
      else if (pfn && T == 60 && dcfg.C_out == 9 && dcfg.variant == LV_PILLAR_PFN && !dcfg.with_distance && !h->vox_generic_rows)
        vx_bins_kernel<VX_OUT_PFN, true, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, d_decorated, pcfg);
      else if (pfn) vx_bins_kernel<VX_OUT_PFN, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, d_decorated, pcfg);
      else if (deco && T == 60 && dcfg.C_out == 9 && dcfg.variant == LV_PILLAR_PFN && !dcfg.with_distance && !h->vox_generic_rows)
        vx_bins_kernel<VX_OUT_DECORATE, true, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, d_decorated, pcfg);
      else if (deco) vx_bins_kernel<VX_OUT_DECORATE, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, d_decorated, pcfg);
      else if (out4) vx_bins_kernel<VX_OUT_VOXELS, true><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, nullptr, pcfg);
      else vx_bins_kernel<VX_OUT_VOXELS, false><<<grid_b, VX_THREADS, smem_bins, stream>>>(p, dcfg, nullptr, pcfg);
      LV_LAUNCH_CHECK(h);
    }
    if (!concat && cfg->zero_tail) {
      int gx = (int)lv_div_up((int64_t)h->num_sms * 8, nf);
      if (gx > V) gx = V;
      vx_zero_tail_kernel<<<dim3((unsigned)gx, (unsigned)nf), VX_THREADS, 0, stream>>>(p);
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
                nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int lv_voxelize_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                                  const int64_t* h_frame_offsets, int64_t capacity_rows, float* d_voxels,
                                  int32_t* d_coords4, int32_t* d_num_points, int32_t* d_voxel_num,
                                  int64_t* d_voxel_offsets, lv_stream stream) {
  LV_REQUIRE(capacity_rows >= 0, "lv_voxelize_concat: negative capacity");
  LV_REQUIRE(n_frames == 0 || d_voxel_offsets, "lv_voxelize_concat: null voxel_offsets");
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, d_voxels, d_coords4, d_num_points, d_voxel_num, 1,
                capacity_rows, d_voxel_offsets, nullptr, nullptr, nullptr, stream);
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
                capacity_rows, d_voxel_offsets, &d, d_decorated, nullptr, stream);
}

extern "C" int lv_pillarize_pfn_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                                       const int64_t* h_frame_offsets, int64_t capacity_rows, float vx, float vy,
                                       float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                                       const float* d_weight, const float* d_scale, const float* d_shift, int32_t units,
                                       float* d_features, int32_t* d_coords4, int32_t* d_num_points,
                                       int32_t* d_voxel_num, int64_t* d_voxel_offsets, lv_stream stream) {
  LV_REQUIRE(cfg != nullptr, "lv_pillarize_pfn_concat: null config");
  LV_REQUIRE(capacity_rows >= 0, "lv_pillarize_pfn_concat: negative capacity");
  LV_REQUIRE(n_frames == 0 || d_voxel_offsets, "lv_pillarize_pfn_concat: null voxel_offsets");
  const int c_out = lv_pillar_out_channels(cfg->num_features, variant, with_distance);
  LV_REQUIRE(c_out == 9, "lv_pillarize_pfn_concat: the fused path needs a 9-channel decoration (variant %d, "
             "with_distance %d gives %d)", variant, with_distance, c_out);
  LV_REQUIRE(units == 64, "lv_pillarize_pfn_concat: units must be 64 (the PointPillars PFN width), got %d", units);
  LV_REQUIRE(d_weight && d_scale && d_shift, "lv_pillarize_pfn_concat: null PFN parameters");
  DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, cfg->max_points, c_out};
  PfnCfg c{d_weight, d_scale, d_shift, units};
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, nullptr, d_coords4, d_num_points, d_voxel_num, 1,
                capacity_rows, d_voxel_offsets, &d, d_features, &c, stream);
}

// Host-buffer entry points in two phases: _begin copies the points in, runs the kernels and returns the voxel counts
// (results stay in the handle's staging buffers); _fetch copies `rows` rows of one frame into caller arrays of exactly
// that size.  VoxelGeneratorV2.generate uses them so that it never allocates, touches or transfers the padded
// (max_voxels, T, C) array (preprocess.py:305-310 slices to voxel_num anyway).
extern "C" int lv_voxelize_host_begin(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                                      const float* h_points, int32_t n_frames, const int64_t* h_frame_offsets,
                                      int32_t* h_voxel_num) {
  LV_REQUIRE(h != nullptr, "lv_voxelize_host: null handle");
  LV_REQUIRE(cfg && h_frame_offsets && n_frames >= 0, "lv_voxelize_host: bad arguments");
  LV_REQUIRE(cfg->num_features >= 3 && cfg->max_points > 0 && cfg->max_voxels > 0, "lv_voxelize_host: bad config");
  h->vox_host_frames = 0;
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_voxel_num != nullptr, "lv_voxelize_host: null output");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const int64_t n_total = h_frame_offsets[n_frames];
  const size_t V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features, F = n_frames;
  const size_t pt_bytes = (size_t)n_total * C * 4;
  LV_REQUIRE(pt_bytes == 0 || h_points, "lv_voxelize_host: null points");
  LV_CHECK(h->vox_stage_points.ensure(pt_bytes, st));
  LV_CHECK(h->vox_stage_out[0].ensure(F * V * T * C * 4, st));
  LV_CHECK(h->vox_stage_out[1].ensure(F * V * 3 * 4, st));
  LV_CHECK(h->vox_stage_out[2].ensure(F * V * 4, st));
  LV_CHECK(h->vox_stage_out[3].ensure(F * 4, st));
  if (pt_bytes) LV_CHECK_CUDA(cudaMemcpyAsync(h->vox_stage_points.ptr, h_points, pt_bytes, cudaMemcpyHostToDevice, st));
  if (flt)
    LV_CHECK(lv_voxelize_filtered(h, cfg, flt, h->vox_stage_points.as<float>(), n_frames, h_frame_offsets,
                                  h->vox_stage_out[0].as<float>(), h->vox_stage_out[1].as<int32_t>(),
                                  h->vox_stage_out[2].as<int32_t>(), h->vox_stage_out[3].as<int32_t>(), nullptr, st));
  else
    LV_CHECK(lv_voxelize(h, cfg, h->vox_stage_points.as<float>(), n_frames, h_frame_offsets, h->vox_stage_out[0].as<float>(),
                         h->vox_stage_out[1].as<int32_t>(), h->vox_stage_out[2].as<int32_t>(),
                         h->vox_stage_out[3].as<int32_t>(), st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h_voxel_num, h->vox_stage_out[3].ptr, F * 4, cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  h->vox_host_frames = n_frames;
  h->vox_host_cfg = *cfg;
  return LV_OK;
}

extern "C" int lv_voxelize_host_fetch(lv_handle* h, int32_t frame, int32_t rows, float* h_voxels, int32_t* h_coords,
                                      int32_t* h_num_points) {
  LV_REQUIRE(h != nullptr, "lv_voxelize_host_fetch: null handle");
  LV_REQUIRE(frame >= 0 && frame < h->vox_host_frames, "lv_voxelize_host_fetch: frame %d is not part of the last "
             "lv_voxelize_host_begin (%d frames)", frame, h->vox_host_frames);
  const size_t V = h->vox_host_cfg.max_voxels, T = h->vox_host_cfg.max_points, C = h->vox_host_cfg.num_features, f = frame;
  LV_REQUIRE(rows >= 0 && (size_t)rows <= V, "lv_voxelize_host_fetch: rows %d outside [0, max_voxels]", rows);
  if (rows == 0) return LV_OK;
  LV_REQUIRE(h_voxels && h_coords && h_num_points, "lv_voxelize_host_fetch: null output");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  LV_CHECK_CUDA(cudaMemcpyAsync(h_voxels, h->vox_stage_out[0].as<float>() + f * V * T * C, (size_t)rows * T * C * 4,
                                cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h_coords, h->vox_stage_out[1].as<int32_t>() + f * V * 3, (size_t)rows * 3 * 4,
                                cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h_num_points, h->vox_stage_out[2].as<int32_t>() + f * V, (size_t)rows * 4,
                                cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  return LV_OK;
}

static int vx_host_run(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt, const float* h_points,
                       int32_t n_frames, const int64_t* h_frame_offsets, float* h_voxels, int32_t* h_coords,
                       int32_t* h_num_points, int32_t* h_voxel_num) {
  LV_CHECK(lv_voxelize_host_begin(h, cfg, flt, h_points, n_frames, h_frame_offsets, h_voxel_num));
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_voxels && h_coords && h_num_points, "lv_voxelize_host: null output");
  const size_t V = cfg->max_voxels, T = cfg->max_points, C = cfg->num_features;
  // with zero_tail == 0 only the live rows are brought back
  for (int32_t f = 0; f < n_frames; ++f)
    LV_CHECK(lv_voxelize_host_fetch(h, f, cfg->zero_tail ? (int32_t)V : h_voxel_num[f], h_voxels + (size_t)f * V * T * C,
                                    h_coords + (size_t)f * V * 3, h_num_points + (size_t)f * V));
  return LV_OK;
}

extern "C" int lv_voxelize_host(lv_handle* h, const lv_voxel_config* cfg, const float* h_points, int32_t n_frames,
                                const int64_t* h_frame_offsets, float* h_voxels, int32_t* h_coords,
                                int32_t* h_num_points, int32_t* h_voxel_num) {
  return vx_host_run(h, cfg, nullptr, h_points, n_frames, h_frame_offsets, h_voxels, h_coords, h_num_points, h_voxel_num);
}

extern "C" int lv_voxelize_filtered_host(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                                         const float* h_points, int32_t n_frames, const int64_t* h_frame_offsets,
                                         float* h_voxels, int32_t* h_coords, int32_t* h_num_points,
                                         int32_t* h_voxel_num) {
  LV_REQUIRE(flt != nullptr, "lv_voxelize_filtered_host: null filter");
  return vx_host_run(h, cfg, flt, h_points, n_frames, h_frame_offsets, h_voxels, h_coords, h_num_points, h_voxel_num);
}

extern "C" int lv_unpad_batch(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                              const int32_t* d_num_voxels, int32_t batch_size, int32_t max_voxels, int32_t max_points,
                              int32_t num_features, int32_t coor_cols, float* d_out_voxels, int32_t* d_out_num_points,
                              int32_t* d_out_coors, int64_t* d_out_total, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_unpad_batch: null handle");
  LV_REQUIRE(batch_size >= 0 && max_voxels > 0 && max_points > 0 && num_features > 0 && coor_cols > 0 && coor_cols <= 32,
             "lv_unpad_batch: bad sizes");
  if (batch_size == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors && d_num_voxels && d_out_voxels && d_out_num_points && d_out_coors &&
                 d_out_total, "lv_unpad_batch: null pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  int gx = (int)lv_div_up((int64_t)h->num_sms * 8, batch_size);
  const int max_gx = (int)lv_div_up(max_voxels, 8);
  if (gx > max_gx) gx = max_gx;
  vx_unpad_kernel<<<dim3((unsigned)gx, (unsigned)batch_size), 256, 0, (cudaStream_t)stream_>>>(
      d_voxels, d_num_points, d_coors, d_num_voxels, batch_size, max_voxels, max_points * num_features, coor_cols,
      d_out_voxels, d_out_num_points, d_out_coors, d_out_total);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_voxelize_mean_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points, int32_t n_frames,
                                       const int64_t* h_frame_offsets, int64_t capacity_rows, int32_t num_features_out,
                                       float* d_mean, int32_t* d_coords4, int32_t* d_num_points, int32_t* d_voxel_num,
                                       int64_t* d_voxel_offsets, lv_stream stream) {
  LV_REQUIRE(cfg != nullptr, "lv_voxelize_mean_concat: null config");
  LV_REQUIRE(capacity_rows >= 0, "lv_voxelize_mean_concat: negative capacity");
  LV_REQUIRE(n_frames == 0 || d_voxel_offsets, "lv_voxelize_mean_concat: null voxel_offsets");
  LV_REQUIRE(num_features_out >= 1 && num_features_out <= 4, "lv_voxelize_mean_concat: num_features_out must be 1..4");
  return vx_run(h, cfg, d_points, n_frames, h_frame_offsets, nullptr, d_coords4, d_num_points, d_voxel_num, 1,
                capacity_rows, d_voxel_offsets, nullptr, d_mean, nullptr, stream, num_features_out);
}
