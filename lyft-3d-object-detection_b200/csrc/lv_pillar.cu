// PointPillars decoration, scatter and voxel-mean kernels (sm_100a).
//
// Replaces second/second/pytorch/models/pointpillars.py:203-231 (+ variants
// :117-145, :290-319, :378-411), :444-476 and voxel_encoder.py:219-225.
// Semantics in include/lyft_voxel.h and SURVEY.md Appendix A.3 / A.4.
//
// pillar_decorate_kernel  one warp per pillar, one CTA = 8 pillars.  The ~12 torch
//   elementwise/reduce/cat kernels of the reference collapse into one pass: read
//   (P,T,C) once (float4 per point when C == 4), write (P,T,C_out) once through a
//   shared-memory stage so that the global stores are contiguous.
//   Algorithmic bytes per pillar: T*C*4 + 20 read, T*C_out*4 written.
// pillar_scatter_*        scatter-as-gather: pillars first drop their index into a
//   dense cell->pillar map (kept all -1 between calls), then every canvas element
//   is written exactly once by a tile kernel (128 cells x C channels per CTA) with
//   128-bit streaming stores; the memset of the reference is fused away.
//   Algorithmic bytes: P*(C*4+16) read, B*C*ny*nx*4 written.
#include <algorithm>

#include "lv_common.cuh"
#include "lv_decorate.cuh"

#define PIL_WARPS 8

struct DecorateParams {
  const float* voxels;
  const int32_t* num;
  const int32_t* coors;
  int64_t P;
  int T, C, C_out;
  float vx, vy, x_off, y_off;
  int variant, with_distance;
  float* out;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <bool C4>
__global__ void __launch_bounds__(PIL_WARPS * 32) pillar_decorate_kernel(DecorateParams p) {
  extern __shared__ float stage[];  // [PIL_WARPS][T*C_out]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = p.T * p.C_out;
  const int64_t pil0 = (int64_t)blockIdx.x * PIL_WARPS;
  const int64_t pil = pil0 + warp;
  float* my = stage + warp * per;
  if (pil < p.P) {
    const float* v = p.voxels + pil * p.T * p.C;
    const int num = __ldg(p.num + pil);
    // pass 1: sum of xyz over ALL T slots (padding zeros included), z range
    float sx = 0.f, sy = 0.f, sz = 0.f, zmin = INFINITY, zmax = -INFINITY;
    for (int t = lane; t < p.T; t += 32) {
      float x, y, z;
      if (C4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(v) + t);
        x = q.x; y = q.y; z = q.z;
      } else {
        x = __ldg(v + t * p.C); y = __ldg(v + t * p.C + 1); z = __ldg(v + t * p.C + 2);
      }
      sx += x; sy += y; sz += z;
      zmin = fminf(zmin, z); zmax = fmaxf(zmax, z);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const float fn = (float)num;
    const float mx = __fdiv_rn(sx, fn), my_ = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);  // :208-209
    float height = 0.f;
    if (p.variant == LV_PILLAR_RADIUS_HEIGHT) height = warp_max(zmax) - warp_min(zmin);     // :387-389
    // pillar centre (:213-217): coors*vx + x_offset, separate multiply and add
    const float cx = __fadd_rn(__fmul_rn((float)__ldg(p.coors + pil * 4 + 3), p.vx), p.x_off);
    const float cy = __fadd_rn(__fmul_rn((float)__ldg(p.coors + pil * 4 + 2), p.vy), p.y_off);
    // pass 2: emit
    for (int t = lane; t < p.T; t += 32) {
      float* o = my + t * p.C_out;
      if (t >= num) {  // padding mask (:226-231)
        for (int c = 0; c < p.C_out; ++c) o[c] = 0.f;
        continue;
      }
      float f[8];
      const int nf = p.C < 8 ? p.C : 8;
      if (C4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(v) + t);
        f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
      } else {
        for (int c = 0; c < nf; ++c) f[c] = __ldg(v + t * p.C + c);
      }
      const float x = f[0], y = f[1], z = f[2];
      const float px = x - cx, py = y - cy;
      int k = 0;
      if (p.variant == LV_PILLAR_PFN || p.variant == LV_PILLAR_OLD) {
        const bool old = p.variant == LV_PILLAR_OLD;
        o[k++] = old ? px : x;   // F7: Old overwrites x,y in place before the concat
        o[k++] = old ? py : y;
        o[k++] = z;
        for (int c = 3; c < p.C; ++c) o[k++] = (c < 8) ? f[c] : __ldg(v + t * p.C + c);
      } else {
        o[k++] = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));  // :303-306
        o[k++] = z;
        for (int c = 3; c < p.C; ++c) o[k++] = (c < 8) ? f[c] : __ldg(v + t * p.C + c);
      }
      o[k++] = x - mx; o[k++] = y - my_; o[k++] = z - mz;   // f_cluster (:210)
      o[k++] = px; o[k++] = py;                             // f_center
      if (p.variant == LV_PILLAR_RADIUS_HEIGHT) o[k++] = height;
      if (p.with_distance) {
        const float dx = (p.variant == LV_PILLAR_OLD) ? px : x, dy = (p.variant == LV_PILLAR_OLD) ? py : y;
        o[k++] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(z, z)));
      }
    }
  }
  __syncthreads();
  // contiguous write-out of the CTA's pillars
  int64_t n_here = p.P - pil0;
  if (n_here > PIL_WARPS) n_here = PIL_WARPS;
  const int64_t total = n_here * per;
  float* dst = p.out + pil0 * per;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) dst[i] = stage[i];
}

// Fast path (C == 4, T <= 64): the pillar lives in registers (two float4 per lane), is read
// exactly once, and leaves through the warp's shared-memory stage as 128-bit stores.
__global__ void __launch_bounds__(PIL_WARPS * 32, 6) pillar_decorate_fast_kernel(DecorateParams p, DecoCfg d) {
  extern __shared__ float stage[];  // [PIL_WARPS][T*C_out + 4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = d.T * d.C_out;
  const int64_t warps_total = (int64_t)gridDim.x * PIL_WARPS;
  float* st = stage + warp * (per + 4);
  for (int64_t pil = (int64_t)blockIdx.x * PIL_WARPS + warp; pil < p.P; pil += warps_total) {
    const float4* v = reinterpret_cast<const float4*>(p.voxels) + pil * d.T;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (lane < d.T) a = lv_ld_stream_f4(v + lane);
    if (lane + 32 < d.T) b = lv_ld_stream_f4(v + lane + 32);
    const int num = __ldg(p.num + pil);
    lv_decorate_zero_tail(num, d, p.out + pil * per, lane);  // independent of the loads above
    const int4 co = __ldg(reinterpret_cast<const int4*>(p.coors) + pil);  // b, z, y, x
    lv_decorate_warp(a, b, num, co.z, co.w, d, st, p.out + pil * per, lane);
  }
}

// Decoration + PFNLayer (inference) in one pass over (P,T,4) voxels: the (P,T,C_out) decorated
// tensor is never materialised.  Algorithmic bytes per pillar: T*16 + 20 read, units*4 written.
template <int CIN, int UJ>
__global__ void __launch_bounds__(PIL_WARPS * 32) pillar_pfn_kernel(DecorateParams p, DecoCfg d, PfnCfg c) {
  extern __shared__ __align__(16) float stage[];  // [PIL_WARPS][T * LV_PFN_STRIDE]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * PIL_WARPS;
  float* st = stage + warp * (d.T * LV_PFN_STRIDE);
  PfnRegs<CIN, UJ> regs;
  regs.load(c, lane);
  for (int64_t pil = (int64_t)blockIdx.x * PIL_WARPS + warp; pil < p.P; pil += warps_total) {
    const float4* v = reinterpret_cast<const float4*>(p.voxels) + pil * d.T;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (lane < d.T) a = lv_ld_stream_f4(v + lane);
    if (lane + 32 < d.T) b = lv_ld_stream_f4(v + lane + 32);
    const int num = __ldg(p.num + pil);
    const int4 co = __ldg(reinterpret_cast<const int4*>(p.coors) + pil);  // b, z, y, x
    const int live = lv_decorate_stage(a, b, num, co.z, co.w, d, st, lane, LV_PFN_STRIDE);
    lv_pfn_warp<CIN, UJ>(st, live, d.T, regs, p.out + pil * c.units, lane);
  }
}

template <int CIN>
static void pillar_pfn_launch(int units, unsigned blocks, size_t smem, cudaStream_t stream, const DecorateParams& p,
                              const DecoCfg& d, const PfnCfg& c) {
  if (units == 32) pillar_pfn_kernel<CIN, 1><<<blocks, PIL_WARPS * 32, smem, stream>>>(p, d, c);
  else if (units == 64) pillar_pfn_kernel<CIN, 2><<<blocks, PIL_WARPS * 32, smem, stream>>>(p, d, c);
  else pillar_pfn_kernel<CIN, 4><<<blocks, PIL_WARPS * 32, smem, stream>>>(p, d, c);
}

// ---------------------------------------------------------------- scatter
#define SC_TILE 128   // cells per CTA: every lane owns 4 consecutive cells (one float4 per channel row)
__global__ void __launch_bounds__(256) pillar_index_kernel(const int32_t* __restrict__ coords, int64_t P,
                                                         const int64_t* __restrict__ d_P, int B, int ny, int nx,
                                                         int32_t* __restrict__ map) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d_P) {   // the pillar count lives on the device (no host round trip): P is only the bound
    const int64_t n = __ldg(d_P);
    if (n < P) P = n;
  }
  if (p >= P) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + p);  // b, z, y, x
  if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  map[(int64_t)c.x * ny * nx + (int64_t)c.z * nx + c.w] = (int32_t)p;
}

#define SC_THREADS 256
#define SC_WARPS (SC_THREADS / 32)
#define SC_AHEAD 2368  // tiles between a CTA and the one it prefetches for (2 x 148 SMs x 8 CTAs;
                       // measured 0.937 / 0.935 / 0.945 / 1.018 ms at 1184 / 2368 / 4736 / 9472)

// One CTA = 128 consecutive cells x all C channels of one sample.  No shared memory and at most
// 32 registers, so eight CTAs (64 warps) fit on an SM: the kernel is a store stream whose rate
// follows the number of CTAs in flight (measured: 1.55 / 1.32 / 1.16 / 1.04 ms at 3 / 4 / 5 / 6
// CTAs per SM for the shared-memory-tile version of this kernel).
//   - every lane reads the cell->pillar indices of ITS four cells (a "quad") into registers; the
//     eight warps read the same 512 bytes, so all of them take the same branch below;
//   - no occupied cell in the tile: one 128-bit zero store per lane and channel row;
//   - otherwise every lane gathers four features per row - an empty cell reads pillar 0's row,
//     which stays in L1, and discards it - so the warp never diverges inside the row loop;
//   - every feature row is read exactly once, so each CTA also prefetches, into L2, the rows the
//     CTA SC_AHEAD launches later will gather (0.99 -> 0.935 ms).
// Warp 0 leaves the map empty for the next call once every warp has read it (the barrier sits
// in front of the stores and waits for nothing but the index loads).
template <bool VEC>
__global__ void __launch_bounds__(SC_THREADS, 8) pillar_canvas_kernel(const float* __restrict__ feats, int32_t* __restrict__ map,
                                                                    int C, int64_t ncell, int tiles_per_sample,
                                                                    float* __restrict__ canvas) {
  const int b = blockIdx.x / tiles_per_sample;
  const int t = blockIdx.x - b * tiles_per_sample;
  const int64_t cell0 = (int64_t)t * SC_TILE;
  const int n_here = (int)((ncell - cell0) < SC_TILE ? (ncell - cell0) : SC_TILE);
  int32_t* m = map + (int64_t)b * ncell + cell0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = lane * 4;
  const bool full = VEC && n_here == SC_TILE;
  int4 occ = make_int4(-1, -1, -1, -1);
  if (full) {
    occ = *reinterpret_cast<const int4*>(m + j0);
  } else {
    if (j0 < n_here) occ.x = m[j0];
    if (j0 + 1 < n_here) occ.y = m[j0 + 1];
    if (j0 + 2 < n_here) occ.z = m[j0 + 2];
    if (j0 + 3 < n_here) occ.w = m[j0 + 3];
  }
  const bool mine = (occ.x & occ.y & occ.z & occ.w) >= 0;   // any of the four >= 0 (sign bit clear)
  const bool any = __syncthreads_or(mine);
  if (any && warp == 0 && mine) {
    if (occ.x >= 0) m[j0] = -1;
    if (occ.y >= 0) m[j0 + 1] = -1;
    if (occ.z >= 0) m[j0 + 2] = -1;
    if (occ.w >= 0) m[j0 + 3] = -1;
  }
  // (behind the barrier, so that only this warp's own rows wait for the extra index load)
  // prefetch for a FUTURE CTA: the last warp reads the indices of the tile SC_AHEAD launches ahead
  // and pulls the feature rows of its pillars into L2, so that tile's gathers find them there
  // instead of paying a DRAM round trip per row (every row is read exactly once otherwise)
  if (VEC && warp == SC_WARPS - 1) {
    const int64_t ta = (int64_t)blockIdx.x + SC_AHEAD;
    if (ta < (int64_t)gridDim.x) {
      const int ba = (int)(ta / tiles_per_sample);
      const int64_t ca = (ta - (int64_t)ba * tiles_per_sample) * SC_TILE + j0;
      if (ca + 3 < ncell) {
        const int4 oa = *reinterpret_cast<const int4*>(map + (int64_t)ba * ncell + ca);
        const int pa[4] = {oa.x, oa.y, oa.z, oa.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (pa[k] >= 0)
            for (int c = 0; c < C; c += 32) lv_prefetch_l2(feats + (int64_t)pa[k] * C + c);
      }
    }
  }
  float* row = canvas + ((int64_t)b * C + warp) * ncell + cell0 + j0;
  const int64_t row_step = (int64_t)SC_WARPS * ncell;
  if (!any) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = warp; c < C; c += SC_WARPS, row += row_step) {
      if (full) {
        lv_st_stream_f4(reinterpret_cast<float4*>(row), z4);
      } else {
        if (j0 < n_here) row[0] = 0.f;
        if (j0 + 1 < n_here) row[1] = 0.f;
        if (j0 + 2 < n_here) row[2] = 0.f;
        if (j0 + 3 < n_here) row[3] = 0.f;
      }
    }
    return;
  }
  // feature offsets (pillar * C fits 32 bits: checked by the host); empty cells point at pillar 0
  const unsigned ox = occ.x < 0 ? 0u : (unsigned)occ.x * (unsigned)C, oy = occ.y < 0 ? 0u : (unsigned)occ.y * (unsigned)C,
                 oz = occ.z < 0 ? 0u : (unsigned)occ.z * (unsigned)C, ow = occ.w < 0 ? 0u : (unsigned)occ.w * (unsigned)C;
  // software pipeline: the gathers of the next row are in flight while this row is stored
  float4 nx;
  nx.x = __ldg(feats + ox + warp);
  nx.y = __ldg(feats + oy + warp);
  nx.z = __ldg(feats + oz + warp);
  nx.w = __ldg(feats + ow + warp);
#pragma unroll 1
  for (int c = warp; c < C; c += SC_WARPS, row += row_step) {
    float4 v = nx;
    const int cn = c + SC_WARPS < C ? c + SC_WARPS : c;
    nx.x = __ldg(feats + ox + cn);
    nx.y = __ldg(feats + oy + cn);
    nx.z = __ldg(feats + oz + cn);
    nx.w = __ldg(feats + ow + cn);
    v.x = occ.x >= 0 ? v.x : 0.f;
    v.y = occ.y >= 0 ? v.y : 0.f;
    v.z = occ.z >= 0 ? v.z : 0.f;
    v.w = occ.w >= 0 ? v.w : 0.f;
    if (full) {
      lv_st_stream_f4(reinterpret_cast<float4*>(row), v);
    } else {
      if (j0 < n_here) row[0] = v.x;
      if (j0 + 1 < n_here) row[1] = v.y;
      if (j0 + 2 < n_here) row[2] = v.z;
      if (j0 + 3 < n_here) row[3] = v.w;
    }
  }
}

// The kernel used when C % 32 == 0 and the sample is whole 128-cell tiles (the PointPillars shapes): warp w
// owns the CONTIGUOUS channels [w*C/8, (w+1)*C/8), so a cell's features arrive as 128-bit loads of four
// consecutive channels (one load per occupied cell and four rows - empty cells load nothing), transposed in
// registers into four row stores; PIPE keeps the next four rows' loads in flight behind those stores.
// 40 registers, six CTAs per SM.  Measured on 128 C5 frames (tools/ab_canvas.py): 0.927 ms against 0.934 ms
// for pillar_canvas_kernel (its L1 traffic - four 32-bit loads per lane and row, mostly dummies - is gone);
// <PIPE, 8 CTAs> spills and takes 1.03 ms, <no PIPE, 6> 0.931 ms, a one-word-per-tile "empty" flag that lets
// empty tiles skip the map read changed nothing (0.931 ms): the kernel follows the DRAM write stream.
// prefetch for a FUTURE CTA (see pillar_canvas_kernel): indices of the tile SC_AHEAD launches ahead, feature
// rows of its pillars into L2
__device__ __forceinline__ void sc_prefetch_ahead(const float* __restrict__ feats, const int32_t* __restrict__ map, int C,
                                                  int64_t ncell, int tiles_per_sample, int j0) {
  const int64_t ta = (int64_t)blockIdx.x + SC_AHEAD;
  if (ta >= (int64_t)gridDim.x) return;
  const int ba = (int)(ta / tiles_per_sample);
  const int64_t ca = (ta - (int64_t)ba * tiles_per_sample) * SC_TILE + j0;
  const int4 oa = *reinterpret_cast<const int4*>(map + (int64_t)ba * ncell + ca);
  const int pa[4] = {oa.x, oa.y, oa.z, oa.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (pa[k] >= 0)
      for (int c = 0; c < C; c += 32) lv_prefetch_l2(feats + (int64_t)pa[k] * C + c);
}

template <bool PIPE, int MINB>
__global__ void __launch_bounds__(SC_THREADS, MINB) pillar_canvas_q_kernel(const float* __restrict__ feats, int32_t* __restrict__ map,
                                                                      int C, int64_t ncell, int tiles_per_sample,
                                                                      float* __restrict__ canvas) {
  const int b = blockIdx.x / tiles_per_sample;
  const int t = blockIdx.x - b * tiles_per_sample;
  const int64_t cell0 = (int64_t)t * SC_TILE;
  int32_t* m = map + (int64_t)b * ncell + cell0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = lane * 4;
  const int cpw = C / SC_WARPS;
  float* row = canvas + ((int64_t)b * C + warp * cpw) * ncell + cell0 + j0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int4 occ = *reinterpret_cast<const int4*>(m + j0);
  const bool mine = (occ.x & occ.y & occ.z & occ.w) >= 0;
  const bool any = __syncthreads_or(mine);
  if (any && warp == 0 && mine) {
    if (occ.x >= 0) m[j0] = -1;
    if (occ.y >= 0) m[j0 + 1] = -1;
    if (occ.z >= 0) m[j0 + 2] = -1;
    if (occ.w >= 0) m[j0 + 3] = -1;
  }
  if (warp == SC_WARPS - 1) sc_prefetch_ahead(feats, map, C, ncell, tiles_per_sample, j0);
  if (!any) {
    for (int r = 0; r < cpw; ++r, row += ncell) lv_st_stream_f4(reinterpret_cast<float4*>(row), z4);
    return;
  }
  // offsets in float4 units (pillar * C < 2^32: checked by the host)
  const float4* f4 = reinterpret_cast<const float4*>(feats) + ((warp * cpw) >> 2);
  const unsigned c4 = (unsigned)C >> 2;
  const unsigned ox = (unsigned)occ.x * c4, oy = (unsigned)occ.y * c4, oz = (unsigned)occ.z * c4, ow = (unsigned)occ.w * c4;
  const int nb = cpw >> 2;
  float4 a = z4, bq = z4, c = z4, d = z4;
  if (occ.x >= 0) a = __ldg(f4 + ox);
  if (occ.y >= 0) bq = __ldg(f4 + oy);
  if (occ.z >= 0) c = __ldg(f4 + oz);
  if (occ.w >= 0) d = __ldg(f4 + ow);
#pragma unroll 1
  for (int q = 1; q <= nb; ++q) {
    if (PIPE) {
      float4 na = z4, nbq = z4, nc = z4, nd = z4;
      if (q < nb) {
        if (occ.x >= 0) na = __ldg(f4 + ox + q);
        if (occ.y >= 0) nbq = __ldg(f4 + oy + q);
        if (occ.z >= 0) nc = __ldg(f4 + oz + q);
        if (occ.w >= 0) nd = __ldg(f4 + ow + q);
      }
      lv_st_stream_f4(reinterpret_cast<float4*>(row), make_float4(a.x, bq.x, c.x, d.x));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + ncell), make_float4(a.y, bq.y, c.y, d.y));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + 2 * ncell), make_float4(a.z, bq.z, c.z, d.z));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + 3 * ncell), make_float4(a.w, bq.w, c.w, d.w));
      a = na; bq = nbq; c = nc; d = nd;
    } else {
      lv_st_stream_f4(reinterpret_cast<float4*>(row), make_float4(a.x, bq.x, c.x, d.x));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + ncell), make_float4(a.y, bq.y, c.y, d.y));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + 2 * ncell), make_float4(a.z, bq.z, c.z, d.z));
      lv_st_stream_f4(reinterpret_cast<float4*>(row + 3 * ncell), make_float4(a.w, bq.w, c.w, d.w));
      if (q < nb) {
        if (occ.x >= 0) a = __ldg(f4 + ox + q);
        if (occ.y >= 0) bq = __ldg(f4 + oy + q);
        if (occ.z >= 0) c = __ldg(f4 + oz + q);
        if (occ.w >= 0) d = __ldg(f4 + ow + q);
      }
    }
    row += 4 * ncell;
  }
}

// ---------------------------------------------------------------- voxel mean
__global__ void __launch_bounds__(256) voxel_mean_kernel(const float* __restrict__ voxels, const int32_t* __restrict__ num,
                                                       int64_t P, int T, int C, int Cout, float* __restrict__ out) {
  // one thread per (voxel, channel): T is small (5) for the SECOND configs
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * Cout) return;
  const int64_t p = i / Cout;
  const int c = (int)(i - p * Cout);
  const float* v = voxels + p * T * C + c;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += __ldg(v + t * C);
  out[i] = __fdiv_rn(s, (float)__ldg(num + p));
}

extern "C" int lv_pillar_out_channels(int32_t num_features, int32_t variant, int32_t with_distance) {
  if (num_features < 3) return LV_E_INVALID;
  int c;
  switch (variant) {
    case LV_PILLAR_PFN:
    case LV_PILLAR_OLD: c = num_features + 5; break;             // pointpillars.py:176 / :90
    case LV_PILLAR_RADIUS: c = num_features + 5 - 1; break;      // :262-263
    case LV_PILLAR_RADIUS_HEIGHT: c = num_features + 6 - 1; break;  // :350-351
    default: return LV_E_INVALID;
  }
  return c + (with_distance ? 1 : 0);
}

extern "C" int lv_pillar_decorate(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                                  const int32_t* d_coors, int64_t n_pillars, int32_t max_points,
                                  int32_t num_features, float vx, float vy, float x_offset, float y_offset,
                                  int32_t variant, int32_t with_distance, float* d_out, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_decorate: null handle");
  LV_REQUIRE(n_pillars >= 0 && max_points > 0, "lv_pillar_decorate: bad sizes");
  const int c_out = lv_pillar_out_channels(num_features, variant, with_distance);
  LV_REQUIRE(c_out > 0, "lv_pillar_decorate: bad num_features %d / variant %d", num_features, variant);
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors && d_out, "lv_pillar_decorate: null pointer");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coors) & 15) == 0, "lv_pillar_decorate: coors must be 16-byte aligned");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  DecorateParams p{d_voxels, d_num_points, d_coors, n_pillars, max_points, num_features, c_out,
                   vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, d_out};
  const size_t smem = (size_t)PIL_WARPS * (max_points * c_out + 4) * sizeof(float);
  LV_REQUIRE(smem <= 200 * 1024, "lv_pillar_decorate: max_points*channels too large for the shared-memory stage");
  const int64_t grid = lv_div_up(n_pillars, PIL_WARPS);
  LV_REQUIRE(grid < (1ll << 31), "lv_pillar_decorate: too many pillars");
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool c4 = num_features == 4 && (reinterpret_cast<uintptr_t>(d_voxels) & 15) == 0;
  if (c4 && max_points <= 64) {
    DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, max_points, c_out};
    if (smem > 48 * 1024)
      LV_CHECK_CUDA(cudaFuncSetAttribute(pillar_decorate_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (int64_t)h->num_sms * 12;   // persistent-style: every warp loops over pillars
    if (blocks > grid) blocks = grid;
    pillar_decorate_fast_kernel<<<(unsigned)blocks, PIL_WARPS * 32, smem, stream>>>(p, d);
  } else if (c4) {
    if (smem > 48 * 1024)
      LV_CHECK_CUDA(cudaFuncSetAttribute(pillar_decorate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pillar_decorate_kernel<true><<<(unsigned)grid, PIL_WARPS * 32, smem, stream>>>(p);
  } else {
    if (smem > 48 * 1024)
      LV_CHECK_CUDA(cudaFuncSetAttribute(pillar_decorate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pillar_decorate_kernel<false><<<(unsigned)grid, PIL_WARPS * 32, smem, stream>>>(p);
  }
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_pillar_pfn(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, const int32_t* d_coors,
                             int64_t n_pillars, int32_t max_points, int32_t num_features, float vx, float vy,
                             float x_offset, float y_offset, int32_t variant, int32_t with_distance,
                             const float* d_weight, const float* d_scale, const float* d_shift, int32_t units,
                             float* d_out, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_pfn: null handle");
  LV_REQUIRE(n_pillars >= 0 && max_points > 0 && max_points <= 64 && num_features == 4,
             "lv_pillar_pfn: needs 4 features per point and max_points <= 64 (got %d, %d)", num_features, max_points);
  const int c_out = lv_pillar_out_channels(num_features, variant, with_distance);
  LV_REQUIRE(c_out >= 8 && c_out <= 10, "lv_pillar_pfn: bad variant %d", variant);
  LV_REQUIRE(units == 32 || units == 64 || units == 128, "lv_pillar_pfn: units must be 32, 64 or 128, got %d", units);
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors && d_weight && d_scale && d_shift && d_out, "lv_pillar_pfn: null pointer");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coors) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_voxels) & 15) == 0,
             "lv_pillar_pfn: voxels and coors must be 16-byte aligned");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  DecorateParams p{d_voxels, d_num_points, d_coors, n_pillars, max_points, num_features, c_out,
                   vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, d_out};
  DecoCfg d{vx, vy, x_offset, y_offset, variant, with_distance ? 1 : 0, max_points, c_out};
  PfnCfg c{d_weight, d_scale, d_shift, units};
  const size_t smem = (size_t)PIL_WARPS * max_points * LV_PFN_STRIDE * sizeof(float);   // <= 24 KB
  const unsigned blocks = (unsigned)std::min<int64_t>((int64_t)h->num_sms * 8, lv_div_up(n_pillars, PIL_WARPS));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (c_out == 8) pillar_pfn_launch<8>(units, blocks, smem, stream, p, d, c);        // radius
  else if (c_out == 9) pillar_pfn_launch<9>(units, blocks, smem, stream, p, d, c);   // pfn, old, radius_height, radius+distance
  else pillar_pfn_launch<10>(units, blocks, smem, stream, p, d, c);                  // +distance
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

static int pillar_scatter_run(lv_handle* h, const float* d_feats, const int32_t* d_coords, int64_t n_pillars,
                              const int64_t* d_n_pillars, int32_t channels, int32_t batch_size, int32_t ny, int32_t nx,
                              float* d_canvas, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_scatter: null handle");
  LV_REQUIRE(n_pillars >= 0 && channels > 0 && batch_size >= 0 && ny > 0 && nx > 0, "lv_pillar_scatter: bad sizes");
  if (batch_size == 0) return LV_OK;
  LV_REQUIRE(d_canvas != nullptr, "lv_pillar_scatter: null canvas");
  LV_REQUIRE(n_pillars == 0 || (d_feats && d_coords), "lv_pillar_scatter: null pointer");
  LV_REQUIRE(n_pillars < (1ll << 31), "lv_pillar_scatter: too many pillars");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coords) & 15) == 0, "lv_pillar_scatter: coords must be 16-byte aligned");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t ncell = (int64_t)ny * nx;
  LV_CHECK(h->pil_map.ensure((size_t)batch_size * ncell * sizeof(int32_t), stream, 0xff));
  const int tiles = (int)lv_div_up(ncell, SC_TILE);
  const bool vec = (ncell % 4 == 0) && (reinterpret_cast<uintptr_t>(d_canvas) & 15) == 0;
  // canvas_variant: 0 = auto (128-bit gathers when the shape allows), 1 = always the row-per-step kernel (A/B)
  const bool quad = vec && h->canvas_variant != 1 && ncell % SC_TILE == 0 && channels % 32 == 0 &&
                    (reinterpret_cast<uintptr_t>(d_feats) & 15) == 0;
  // every argument check sits in front of the first launch: a failing call must leave the cell->pillar map all -1
  LV_REQUIRE(n_pillars * (int64_t)channels < (1ll << 32), "lv_pillar_scatter: %lld pillars x %d channels exceed 2^32 features",
             (long long)n_pillars, channels);
  const int64_t grid = (int64_t)tiles * batch_size;
  LV_REQUIRE(grid < (1ll << 31), "lv_pillar_scatter: canvas too large");
  if (n_pillars > 0) {
    pillar_index_kernel<<<(unsigned)lv_div_up(n_pillars, 256), 256, 0, stream>>>(d_coords, n_pillars, d_n_pillars,
                                                                               batch_size, ny, nx, h->pil_map.as<int32_t>());
    LV_LAUNCH_CHECK(h);
  }
  if (quad)
    pillar_canvas_q_kernel<true, 6><<<(unsigned)grid, SC_THREADS, 0, stream>>>(d_feats, h->pil_map.as<int32_t>(), channels,
                                                                             ncell, tiles, d_canvas);
  else if (vec)
    pillar_canvas_kernel<true><<<(unsigned)grid, SC_THREADS, 0, stream>>>(d_feats, h->pil_map.as<int32_t>(), channels, ncell,
                                                                        tiles, d_canvas);
  else
    pillar_canvas_kernel<false><<<(unsigned)grid, SC_THREADS, 0, stream>>>(d_feats, h->pil_map.as<int32_t>(), channels, ncell,
                                                                         tiles, d_canvas);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_pillar_scatter(lv_handle* h, const float* d_feats, const int32_t* d_coords, int64_t n_pillars,
                                 int32_t channels, int32_t batch_size, int32_t ny, int32_t nx, float* d_canvas,
                                 lv_stream stream) {
  return pillar_scatter_run(h, d_feats, d_coords, n_pillars, nullptr, channels, batch_size, ny, nx, d_canvas, stream);
}

extern "C" int lv_pillar_scatter_dev(lv_handle* h, const float* d_feats, const int32_t* d_coords,
                                     const int64_t* d_n_pillars, int64_t max_pillars, int32_t channels,
                                     int32_t batch_size, int32_t ny, int32_t nx, float* d_canvas, lv_stream stream) {
  LV_REQUIRE(d_n_pillars != nullptr, "lv_pillar_scatter_dev: null pillar count");
  return pillar_scatter_run(h, d_feats, d_coords, max_pillars, d_n_pillars, channels, batch_size, ny, nx, d_canvas, stream);
}

extern "C" int lv_voxel_mean(lv_handle* h, const float* d_voxels, const int32_t* d_num_points, int64_t n_voxels,
                             int32_t max_points, int32_t num_features, int32_t num_features_out, float* d_out,
                             lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_voxel_mean: null handle");
  LV_REQUIRE(n_voxels >= 0 && max_points > 0 && num_features > 0 && num_features_out > 0 &&
                 num_features_out <= num_features, "lv_voxel_mean: bad sizes");
  if (n_voxels == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_out, "lv_voxel_mean: null pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  const int64_t items = n_voxels * num_features_out;
  voxel_mean_kernel<<<(unsigned)lv_div_up(items, 256), 256, 0, (cudaStream_t)stream_>>>(
      d_voxels, d_num_points, n_voxels, max_points, num_features, num_features_out, d_out);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}
