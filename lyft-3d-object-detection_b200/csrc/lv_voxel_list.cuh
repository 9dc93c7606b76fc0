// Hard voxelizer, "list path" (sm_100a): three kernels instead of six for grids whose dense map is small
// (the pillar configs).  Included by lv_voxel.cu; same semantics, same outputs, bit for bit
// (second/second/utils/simplevis.py:9-61).
//
// The six-kernel path rebuilds point order inside a voxel by a stable two-level split (K3-K5) and an
// in-shared-memory ranking (K6).  Here the order is rebuilt where it is cheap - inside the warp that writes the
// voxel's row - and everything before that only has to get every point into ITS VOXEL'S LIST, in any order:
//
//   L1 vl_cells   cell(i) as K1; on the 8-byte map entry {first, count} of the cell: RED.MIN(first, i) and
//                 a = ATOM.ADD(count, 1) (both warp-aggregated) - `a` is the point's ARRIVAL rank in its cell.
//   L2 vl_assign  creator(i) := first[cell(i)] == i.  ONE single-pass scan (decoupled look-back) over the
//                 creators in point order carries two sums: creators so far (= voxel id, first-come order) and
//                 points of the voxels created so far (= where the voxel's list starts).  Creators publish
//                 {~id, list start} in the map entry with one 64-bit store; every other point reads its cell's
//                 entry (spinning until the creator - always an earlier point, hence an earlier or the same
//                 CTA - has published) and drops its index into list[start + arrival rank].  max_voxels:
//                 `continue` keeps ids < V; `break` also turns every point whose inclusive creator count
//                 exceeds V into a sentinel (simplevis.py:46-50).
//   L3 vl_rows    one warp per voxel (two small voxels share a warp): reads the voxel's list, sorts it in
//                 registers (rank by counting over shuffles; lists longer than 64 by repeated minimum), keeps
//                 the first T indices, gathers the points and writes the output row (raw voxel, fused
//                 PillarFeatureNet decoration, or decoration + PFNLayer) exactly as K6 does - but without a
//                 slot table in shared memory, without a CTA barrier, and with a grid that covers only
//                 voxels that exist.  The warp also puts EMPTY back into the voxel's map entry.
//
// Workspace per point: cell 4 B, arrival rank 4 B, list 4 B; per voxel 16 B; map 8 B per cell.
//
// OPT-IN (lv_set_option "vox_list_path" 1).  Bit-identical to the six-kernel path (tests/test_gpu_voxel_list.py
// runs both against the oracle), but measured SLOWER on 128 C5 frames (profiles/r02_list_path.txt): pillarize stage
// 1.03 ms against 0.67 ms.  ncu per launch: vl_cells 75 us (vx_cells 46: the ATOM that returns the arrival rank
// costs 30 us over a RED), vl_assign 127 us (vx_assign + vx_keys + vx_scan_hist + vx_scatter = 131 us: the second,
// spinning gather behind the look-back barrier runs at 43 % warp occupancy and 11 % DRAM - this kernel IS the
// "keys fused into assign" variant, and fusing buys nothing), vl_rows 800 us (vx_bins 470): three DEPENDENT global
// loads per voxel (record -> list -> points) with one pair of voxels in flight per warp leave the warps on the
// long scoreboard 80 % of the time, where vx_bins keeps records and lists in shared memory and has one.
#pragma once

#define VL_BIAS VX_EMPTY            // count field of an untouched entry (memset-able like `first`)
#define VL_SENT 0x7fffffff          // list entry of a point dropped by the `break` rule
#define VL_RANK_BITS 26             // look-back value = list offset << 26 | creators
#define VL_RANK_MASK ((1ull << VL_RANK_BITS) - 1)
#define VL_ITEMS_PER_CTA 16         // two voxels per warp and iteration

struct VlParams {
  unsigned long long* map2;         // [frames][G] entries {lo: first / ~id, hi: count / list start}
  uint32_t* arr;                    // [points] arrival rank
  int32_t* list;                    // [points] voxel lists of a frame, at the frame's first point slot
  int4* vrec;                       // [points] per voxel of a frame: cell, list start, count
  int32_t* frame_total;             // [frames] creators of the frame (may exceed V)
};

__device__ __forceinline__ unsigned long long vl_ld_entry(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void vl_st_entry(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------- L1: cells, first index, arrival rank
template <bool C4>
__global__ void __launch_bounds__(VX_THREADS) vl_cells_kernel(VoxParams p, VlParams q) {
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
  }
  int* map = reinterpret_cast<int*>(q.map2 + (int64_t)L.fl * p.G);   // entry e: map[2e] = first, map[2e+1] = count
  if (staged) {
    __syncthreads();  // the barrier is initialised
    lv_mbar_wait(&bar, 0);
  }
  const int base = L.c * VX_CHUNK + warp * (32 * VX_ITEMS);
  const unsigned lt = lv_lanemask_lt();
  int cell[VX_ITEMS], got[VX_ITEMS];
  unsigned peers[VX_ITEMS];
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int ti = warp * (32 * VX_ITEMS) + r * 32 + lane;  // row inside the chunk
    const int li = base + r * 32 + lane;                    // local point index inside the frame
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
        const float* s = tile + ti * p.C;
        x = s[0]; y = s[1]; z = s[2];
      } else {
        const float* s = p.pts + gi * p.C;
        x = __ldg(s); y = __ldg(s + 1); z = __ldg(s + 2);
      }
      // simplevis.py:38-42: c = floor((p - lo) / vs) in float32, bounds on the float
      const float cx = floorf(__fdiv_rn(__fsub_rn(x, p.lo[0]), p.vs[0]));
      const float cy = floorf(__fdiv_rn(__fsub_rn(y, p.lo[1]), p.vs[1]));
      const float cz = floorf(__fdiv_rn(__fsub_rn(z, p.lo[2]), p.vs[2]));
      if (cx >= 0.f && cx < (float)p.grid[0] && cy >= 0.f && cy < (float)p.grid[1] && cz >= 0.f &&
          cz < (float)p.grid[2])
        cell[r] = ((int)cz * p.grid[1] + (int)cy) * p.grid[0] + (int)cx;
      p.cell[gi - p.pt_lo] = cell[r];
    }
  }
  // all atomics of the eight rounds are issued before the first answer is needed
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    peers[r] = __match_any_sync(0xffffffffu, cell[r]);
    got[r] = 0;
    if (cell[r] >= 0 && lane == __ffs(peers[r]) - 1) {   // the lowest lane of a cell = its lowest index in the round
      atomicMin(map + 2 * cell[r], base + r * 32 + lane);
      got[r] = atomicAdd(map + 2 * cell[r] + 1, __popc(peers[r]));
    }
  }
#pragma unroll
  for (int r = 0; r < VX_ITEMS; ++r) {
    const int b = __shfl_sync(0xffffffffu, got[r], __ffs(peers[r]) - 1);
    const int li = base + r * 32 + lane;
    if (cell[r] >= 0) q.arr[L.start + li - p.pt_lo] = (uint32_t)(b - VL_BIAS) + __popc(peers[r] & lt);
  }
}

// ---------------------------------------------------------------- L2: voxel ids, list starts, lists
__global__ void __launch_bounds__(VX_THREADS) vl_assign_kernel(VoxParams p, VlParams q) {
  __shared__ unsigned long long sw[VX_WARPS];
  __shared__ unsigned long long s_excl;
  const ChunkLoc L = vx_locate(p);
  unsigned long long* map = q.map2 + (int64_t)L.fl * p.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread-contiguous items: thread t owns local points base + t*8 .. +7 (index order)
  const int base = L.c * VX_CHUNK + threadIdx.x * VX_ITEMS;
  const int64_t w0 = L.start - p.pt_lo;   // workspace index of the frame's first point
  int cell[VX_ITEMS], cnt[VX_ITEMS];   // cnt: points of the voxel a creator creates (0 for the other points)
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    const int li = base + k;
    cell[k] = li < L.n ? p.cell[w0 + li] : -1;
  }
  unsigned flags = 0;
  unsigned long long mine = 0;                              // creators | points of the voxels they create << 26
  {
    unsigned long long ent[VX_ITEMS];
#pragma unroll
    for (int k = 0; k < VX_ITEMS; ++k) ent[k] = cell[k] >= 0 ? map[cell[k]] : 0ull;   // L1 has finished: plain loads
#pragma unroll
    for (int k = 0; k < VX_ITEMS; ++k) {
      cnt[k] = 0;
      if (cell[k] >= 0 && (int)(unsigned)ent[k] == base + k) {
        flags |= 1u << k;
        cnt[k] = (int)(ent[k] >> 32) - VL_BIAS;
        mine += 1ull + ((unsigned long long)(unsigned)cnt[k] << VL_RANK_BITS);
      }
    }
  }
  unsigned long long inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sw[warp] = inc;
  __syncthreads();
  unsigned long long woff = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < VX_WARPS; ++w) {
    if (w < warp) woff += sw[w];
    agg += sw[w];
  }
  // decoupled look-back over the previous chunks of this frame (warp 0)
  if (warp == 0) {
    unsigned long long* st = p.chunk_state + blockIdx.x;
    unsigned long long excl = 0;
    if (L.c == 0) {
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | agg);
    } else {
      if (lane == 0) vx_st_state(st, VX_FLAG_AGG | agg);
      int look = L.c - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long v = VX_FLAG_PREFIX;  // lanes before the frame start act as a zero prefix
        if (idx >= 0) {
          do { v = vx_ld_state(st - (L.c - idx)); } while ((v >> 62) == 0);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_p = __ffs(is_prefix) - 1;
        unsigned long long contrib = (first_p < 0 || lane <= first_p) ? (v & ((1ull << 62) - 1)) : 0ull;
        if (idx < 0) contrib = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (first_p >= 0) break;
        look -= 32;
      }
      if (lane == 0) vx_st_state(st, VX_FLAG_PREFIX | (excl + agg));
    }
    if (lane == 0) {
      s_excl = excl;
      if (L.c == L.nchunks - 1) {
        const int total = (int)((excl + agg) & VL_RANK_MASK);
        p.voxel_num[L.f] = total < p.V ? total : p.V;
        q.frame_total[L.fl] = total;
      }
    }
  }
  __syncthreads();
  const unsigned long long texcl = s_excl + woff + inc - mine;
  int rank = (int)(texcl & VL_RANK_MASK);
  unsigned start = (unsigned)(texcl >> VL_RANK_BITS);
  const int rank0 = rank;
  // creators publish {~id, list start} and the voxel's record
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    if (!(flags & (1u << k))) continue;
    vl_st_entry(map + cell[k], ((unsigned long long)start << 32) | (unsigned)(~rank));
    q.vrec[w0 + rank] = make_int4(cell[k], (int)start, cnt[k], 0);   // rank < points of the frame
    ++rank;
    start += (unsigned)cnt[k];
  }
  __syncthreads();   // the creators of this CTA have published
  int32_t* list = q.list + w0;
  // first attempt at every entry and the arrival ranks: sixteen independent loads
  unsigned long long ent[VX_ITEMS];
  unsigned arr[VX_ITEMS];
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    ent[k] = cell[k] >= 0 ? vl_ld_entry(map + cell[k]) : 0ull;
    arr[k] = cell[k] >= 0 ? q.arr[w0 + base + k] : 0u;
  }
#pragma unroll
  for (int k = 0; k < VX_ITEMS; ++k) {
    if (cell[k] < 0) continue;
    unsigned long long e = ent[k];
    // the creator is an earlier point: same CTA (published above) or an earlier chunk (running or done)
    while ((int)(unsigned)e >= 0) e = vl_ld_entry(map + cell[k]);
    const int vid = ~(int)(unsigned)e;
    if (vid >= p.V) continue;                                  // voxel beyond max_voxels: no row, no list
    // creators at or before this point (simplevis.py:46-50: the point that would create voxel V ends the loop)
    const int incl = rank0 + __popc(flags & ((2u << k) - 1u));
    const bool keep = p.overflow != LV_OVERFLOW_BREAK || incl <= p.V;
    list[(unsigned)(e >> 32) + arr[k]] = keep ? base + k : VL_SENT;
  }
}

// frames without points own no chunk, so L2 never wrote their totals
__global__ void __launch_bounds__(256) vl_empty_frames_kernel(VoxParams p, VlParams q) {
  const int fl = blockIdx.x * blockDim.x + threadIdx.x;
  if (fl >= p.f1 - p.f0) return;
  const int f = p.f0 + fl;
  if (__ldg(p.frame_chunk + f + 1) == __ldg(p.frame_chunk + f)) {
    q.frame_total[fl] = 0;
    p.voxel_num[f] = 0;
  }
}

// ---------------------------------------------------------------- L3: lists -> output rows
// The T smallest indices of a list, ascending, into s[0..): returns how many (<= T).  n <= 64: every lane
// ranks its (up to two) entries by counting smaller ones over shuffles; longer lists: repeated minimum.
__device__ __forceinline__ int vl_sort_list(const int32_t* __restrict__ lst, int n, int T, int* s, int lane) {
  if (n <= 64) {
    const int a = lane < n ? lst[lane] : VL_SENT, b = lane + 32 < n ? lst[lane + 32] : VL_SENT;
    int ra = 0, rb = 0;
    if (n <= 32) {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) ra += __shfl_sync(0xffffffffu, a, k) < a;
    } else {
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const int va = __shfl_sync(0xffffffffu, a, k), vb = __shfl_sync(0xffffffffu, b, k);
        ra += (va < a) + (vb < a);
        rb += (va < b) + (vb < b);
      }
    }
    if (a != VL_SENT && ra < T) s[ra] = a;
    if (b != VL_SENT && rb < T) s[rb] = b;
    const int valid = __popc(__ballot_sync(0xffffffffu, a != VL_SENT)) + __popc(__ballot_sync(0xffffffffu, b != VL_SENT));
    __syncwarp();
    return valid < T ? valid : T;
  }
  int last = -1, t = 0;
  for (; t < T; ++t) {
    int m = VL_SENT;
    for (int i = lane; i < n; i += 32) {
      const int v = lst[i];
      if (v > last && v < m) m = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (m == VL_SENT) break;
    if (lane == 0) s[t] = m;
    last = m;
  }
  __syncwarp();
  return t;
}

// One voxel by the whole warp (lists longer than 16, or a voxel without a small partner).  Out of line: the
// common path of vl_rows_kernel - two small pillars per warp - keeps its registers.
template <int MODE>
__device__ __forceinline__ void vl_row_body(const VoxParams& p, const int32_t* __restrict__ lst, int cnt, int cell, int f,
                                            int64_t fstart, long long row, const DecoCfg& d, float* __restrict__ decorated,
                                            const PfnCfg& pfn, const PfnRegs<9, 2>& pfn_regs, int* srt, float* st, int lane) {
  constexpr bool DECO = MODE == VX_OUT_DECORATE || MODE == VX_OUT_PFN;
  const int per = DECO ? d.T * d.C_out : 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = vl_sort_list(lst, cnt, p.T, srt, lane);
      const float4* pts4 = reinterpret_cast<const float4*>(p.pts) + fstart;
      int cx, cy, cz;
      vx_cell_coords(p, cell, cz, cy, cx);
      if (lane == 0) {
        p.num_points[row] = n;
        if (p.coord_cols == 4) {
          *reinterpret_cast<int4*>(p.coords + row * 4) = make_int4(f, cz, cy, cx);  // preprocess.py:44-50
        } else {
          int32_t* co = p.coords + row * 3;
          co[0] = cz; co[1] = cy; co[2] = cx;  // reversed (z,y,x), simplevis.py:42
        }
      }
      if (DECO) {
        float* dst = decorated + row * per;
        if (MODE == VX_OUT_DECORATE) lv_decorate_zero_tail(n, d, dst, lane);
        float4 a = z4, b = z4;
        if (lane < n) a = __ldg(pts4 + srt[lane]);
        if (lane + 32 < n) b = __ldg(pts4 + srt[lane + 32]);
        if (MODE == VX_OUT_DECORATE) {
          lv_decorate_warp(a, b, n, cy, cx, d, st, dst, lane);
        } else {
          const int live = lv_decorate_stage(a, b, n, cy, cx, d, st, lane, LV_PFN_STRIDE);
          lv_pfn_warp<9, 2>(st, live, d.T, pfn_regs, decorated + row * pfn.units, lane);
        }
      } else {
        // raw voxel row (T, 4): live slots gather their point, padding slots are zeros from registers
        float4* o4 = reinterpret_cast<float4*>(p.voxels) + row * p.T;
        for (int t = lane; t < p.T; t += 32) {
          float4 v = z4;
          if (t < n) v = __ldg(pts4 + srt[t]);
          lv_st_stream_f4(o4 + t, v);
        }
      }
}

template <int MODE>
__device__ __noinline__ void vl_row_general(const VoxParams& p, const int32_t* __restrict__ lst, int cnt, int cell, int f,
                                            int64_t fstart, long long row, const DecoCfg& d, float* __restrict__ decorated,
                                            const PfnCfg& pfn, int* srt, float* st, int lane) {
  PfnRegs<9, 2> none;   // only the PFN mode reads it, and that mode inlines vl_row_body
  vl_row_body<MODE>(p, lst, cnt, cell, f, fstart, row, d, decorated, pfn, none, srt, st, lane);
}

template <int MODE>
__global__ void __launch_bounds__(VX_THREADS, MODE == VX_OUT_PFN ? 3 : VX_BINS_MINB)
vl_rows_kernel(const __grid_constant__ VoxParams p, const __grid_constant__ VlParams q, const __grid_constant__ DecoCfg d,
               float* __restrict__ decorated, const __grid_constant__ PfnCfg pfn) {
  constexpr bool DECO = MODE == VX_OUT_DECORATE || MODE == VX_OUT_PFN;
  extern __shared__ __align__(16) int smem[];
  const int nf = p.f1 - p.f0;
  int* s_work = smem;                               // [nf+1] creators before frame fl
  int* s_rows = s_work + nf + 1;                    // [nf+1] kept voxels before frame fl
  int* s_sort = smem + ((2 * (nf + 1) + 3) & ~3);   // [8][64] sorted point indices (16-byte aligned)
  float* stage = reinterpret_cast<float*>(s_sort + VX_WARPS * 64);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    int cw = 0, cr = 0;
    for (int f0 = 0; f0 <= nf; f0 += 32) {
      const int fl = f0 + lane;
      const int tw = fl < nf ? q.frame_total[fl] : 0;
      const int tr = fl < nf ? (tw < p.V ? tw : p.V) : 0;
      int iw = tw, ir = tr;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, iw, o), b = __shfl_up_sync(0xffffffffu, ir, o);
        if (lane >= o) { iw += a; ir += b; }
      }
      if (fl <= nf) { s_work[fl] = cw + iw - tw; s_rows[fl] = cr + ir - tr; }
      cw += __shfl_sync(0xffffffffu, iw, 31);
      cr += __shfl_sync(0xffffffffu, ir, 31);
    }
  }
  __syncthreads();
  const int total_work = s_work[nf];
  const long long row00 = p.concat ? p.row_base[p.f0] : 0;
  if (blockIdx.x == 0 && p.concat)
    for (int fl = threadIdx.x + 1; fl <= nf; fl += VX_THREADS) p.row_base[p.f0 + fl] = row00 + s_rows[fl];
  int* srt = s_sort + warp * 64;
  const int per = DECO ? d.T * d.C_out : 0;
  float* st = stage + warp * (MODE == VX_OUT_PFN ? d.T * LV_PFN_STRIDE : per + 4);
  PfnRegs<9, 2> pfn_regs;
  if (MODE == VX_OUT_PFN) pfn_regs.load(pfn, lane);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  const int hi = lane >> 4, sub = lane & 15;
  for (int blk = blockIdx.x; blk * VL_ITEMS_PER_CTA < total_work; blk += gridDim.x) {
    // every half-warp looks up ITS item: lanes 0-15 item w, lanes 16-31 item w + 1
    const int w = blk * VL_ITEMS_PER_CTA + warp * 2 + hi;
    int m_f = -1, m_cell = 0, m_start = 0, m_cnt = 0;
    long long m_row = -1;
    int64_t m_fstart = 0;
    if (w < total_work) {
      int lo = 0, up = nf;                              // largest fl with s_work[fl] <= w
      while (up - lo > 1) {
        const int mid = (lo + up) >> 1;
        if (s_work[mid] <= w) lo = mid; else up = mid;
      }
      m_f = p.f0 + lo;
      const int vid = w - s_work[lo];
      m_fstart = __ldg(p.frame_off + m_f);
      const int4 rec = q.vrec[m_fstart - p.pt_lo + vid];
      m_cell = rec.x; m_start = rec.y; m_cnt = rec.z;
      if (vid < p.V) {
        const long long row = p.concat ? row00 + s_rows[lo] + vid : (long long)m_f * p.V + vid;
        if (row < p.capacity) m_row = row;
      }
      // touched-cell reset: this half-warp is the last user of the entry
      if (sub == 0) q.map2[(int64_t)lo * p.G + m_cell] = ((unsigned long long)(unsigned)VL_BIAS << 32) | (unsigned)VX_EMPTY;
    }
    const bool ok = m_row >= 0 && m_cnt <= 16;
    const bool small = DECO && p.T >= 32 && __all_sync(0xffffffffu, ok);
    if (small) {
      // two small pillars share the warp (bit-identical sums, see lv_decorate_half_stage)
      const int32_t* lst = q.list + (m_fstart - p.pt_lo) + m_start;
      const int a_idx = sub < m_cnt ? lst[sub] : VL_SENT;
      const int n0 = __popc(__ballot_sync(0xffffffffu, a_idx != VL_SENT && hi == 0));
      const int n1 = __popc(__ballot_sync(0xffffffffu, a_idx != VL_SENT && hi == 1));
      const int nm = hi ? n1 : n0;                     // <= 16 <= T
      const long long row0 = __shfl_sync(0xffffffffu, m_row, 0), row1 = __shfl_sync(0xffffffffu, m_row, 16);
      if (MODE == VX_OUT_DECORATE) {
        lv_decorate_zero_tail(n0, d, decorated + row0 * per, lane);
        lv_decorate_zero_tail(n1, d, decorated + row1 * per, lane);
      }
      int ra = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) ra += __shfl_sync(0xffffffffu, a_idx, (lane & 16) | k) < a_idx;
      if (a_idx != VL_SENT) srt[hi * 16 + ra] = a_idx;
      __syncwarp();
      float4 a = z4;
      if (sub < nm) a = __ldg(reinterpret_cast<const float4*>(p.pts) + m_fstart + srt[hi * 16 + sub]);
      int cx, cy, cz;
      vx_cell_coords(p, m_cell, cz, cy, cx);
      if (sub == 0) {
        p.num_points[m_row] = nm;
        *reinterpret_cast<int4*>(p.coords + m_row * 4) = make_int4(m_f, cz, cy, cx);  // preprocess.py:44-50
      }
      if (MODE == VX_OUT_DECORATE) {
        lv_decorate_half(a, nm, cy, cx, d, st + hi * (16 * d.C_out + 4), decorated + m_row * per, lane);
      } else {
        lv_decorate_half_stage(a, nm, cy, cx, d, st + hi * (16 * LV_PFN_STRIDE), lane, LV_PFN_STRIDE);
        lv_pfn_warp<9, 2>(st, n0, d.T, pfn_regs, decorated + row0 * pfn.units, lane);
        lv_pfn_warp<9, 2>(st + 16 * LV_PFN_STRIDE, n1, d.T, pfn_regs, decorated + row1 * pfn.units, lane);
      }
      __syncwarp();
      continue;
    }
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      // the whole warp works on item j: its half-warp's metadata is broadcast
      const int src = j * 16;
      const long long row = __shfl_sync(0xffffffffu, m_row, src);
      if (row < 0) continue;   // warp-uniform
      const int f = __shfl_sync(0xffffffffu, m_f, src), cell = __shfl_sync(0xffffffffu, m_cell, src);
      const int start = __shfl_sync(0xffffffffu, m_start, src), cnt = __shfl_sync(0xffffffffu, m_cnt, src);
      const int64_t fstart = __shfl_sync(0xffffffffu, m_fstart, src);
      if (MODE == VX_OUT_PFN)
        vl_row_body<MODE>(p, q.list + (fstart - p.pt_lo) + start, cnt, cell, f, fstart, row, d, decorated, pfn, pfn_regs, srt, st, lane);
      else
        vl_row_general<MODE>(p, q.list + (fstart - p.pt_lo) + start, cnt, cell, f, fstart, row, d, decorated, pfn, srt, st, lane);
      __syncwarp();
    }
  }
}
