// Warp-level pillar decoration shared by the standalone decorate kernel
// (lv_pillar.cu) and the fused voxelize+decorate gather (lv_voxel.cu).
//
// Reference arithmetic: second/second/pytorch/models/pointpillars.py:203-231
// (+ variants :117-145, :290-319, :378-411); SURVEY.md Appendix A.3.
#pragma once
#include "lv_common.cuh"

struct DecoCfg {
  float vx, vy, x_off, y_off;  // pointpillars.py:198-201, rounded to float32 like torch does
  int variant, with_distance;
  int T, C_out;
};

__device__ __forceinline__ float lv_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lv_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float lv_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one LIVE slot of one pillar -> C_out floats at o (C == 4 input features)
__device__ __forceinline__ void lv_decorate_slot(const float4 q, float mx, float my, float mz, float cx, float cy,
                                                 float height, const DecoCfg& d, float* o) {
  const float x = q.x, y = q.y, z = q.z;
  const float px = x - cx, py = y - cy;  // f_center (:213-217)
  int k = 0;
  if (d.variant == LV_PILLAR_PFN) {
    o[0] = x; o[1] = y; o[2] = z; o[3] = q.w;
    k = 4;
  } else if (d.variant == LV_PILLAR_OLD) {  // x,y overwritten in place before the concat
    o[0] = px; o[1] = py; o[2] = z; o[3] = q.w;
    k = 4;
  } else {  // radius variants (:303-306): |xy|, z, features
    o[0] = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    o[1] = z; o[2] = q.w;
    k = 3;
  }
  o[k] = x - mx; o[k + 1] = y - my; o[k + 2] = z - mz;  // f_cluster (:210)
  o[k + 3] = px; o[k + 4] = py;
  k += 5;
  if (d.variant == LV_PILLAR_RADIUS_HEIGHT) o[k++] = height;
  if (d.with_distance) {
    const float dx = (d.variant == LV_PILLAR_OLD) ? px : x, dy = (d.variant == LV_PILLAR_OLD) ? py : y;
    o[k] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(z, z)));
  }
}

// An output row is [num*C_out floats of data][zeros] (padding mask, :226-231): most of a
// pillar is padding (5.4 of 60 slots are live on a Lyft sweep), so the zero part is stored
// straight from registers - it depends on nothing but `num` and is issued BEFORE the point
// gathers are consumed - and only the live slots go through the shared-memory stage.
__device__ __forceinline__ bool lv_row_vec4(const DecoCfg& d, const float* dst) {
  return ((d.T * d.C_out) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
}
__device__ __forceinline__ int lv_live_slots(int num, const DecoCfg& d) { return num < 0 ? 0 : (num > d.T ? d.T : num); }
__device__ __forceinline__ void lv_decorate_zero_tail(int num, const DecoCfg& d, float* __restrict__ dst, int lane) {
  const int per = d.T * d.C_out, nd = lv_live_slots(num, d) * d.C_out;
  if (lv_row_vec4(d, dst)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = ((nd + 3) >> 2) + lane; j < (per >> 2); j += 32) lv_st_stream_f4(d4 + j, z4);
  } else {
    for (int i = nd + lane; i < per; i += 32) dst[i] = 0.f;
  }
}

// Decorates the live part of one pillar (T <= 64, C == 4) held in registers by one warp:
// lane owns slot `lane` (a) and slot `lane + 32` (b); slots >= num must be passed as zeros.
// The live slots land in the warp's shared-memory region `st`, slot t at st + t*slot_stride
// (C_out for the dense output layout); returns the number of live slots.  Ends with __syncwarp: `st` is readable by every lane.
__device__ __forceinline__ int lv_decorate_stage(const float4 a, const float4 b, int num, int coor_y, int coor_x,
                                                 const DecoCfg& d, float* st, int lane, int slot_stride) {
  // mean over ALL T slots, padding zeros included (:208-209)
  const float sx = lv_warp_sum(a.x + b.x), sy = lv_warp_sum(a.y + b.y), sz = lv_warp_sum(a.z + b.z);
  const float fn = (float)num;
  const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
  float height = 0.f;
  if (d.variant == LV_PILLAR_RADIUS_HEIGHT) {  // :387-389, min/max over all T slots
    const bool ha = lane < d.T, hb = lane + 32 < d.T;
    const float zmax = fmaxf(ha ? a.z : -INFINITY, hb ? b.z : -INFINITY);
    const float zmin = fminf(ha ? a.z : INFINITY, hb ? b.z : INFINITY);
    height = lv_warp_max(zmax) - lv_warp_min(zmin);
  }
  // pillar centre: coors*vx + x_offset as a separate multiply and add (:213-217)
  const float cx = __fadd_rn(__fmul_rn((float)coor_x, d.vx), d.x_off);
  const float cy = __fadd_rn(__fmul_rn((float)coor_y, d.vy), d.y_off);
  const int live = lv_live_slots(num, d);
  if (lane < live) lv_decorate_slot(a, mx, my, mz, cx, cy, height, d, st + lane * slot_stride);
  if (lane + 32 < live) lv_decorate_slot(b, mx, my, mz, cx, cy, height, d, st + (lane + 32) * slot_stride);
  const int nd = live * slot_stride;
  if (lane < (((nd + 3) >> 2) << 2) - nd) st[nd + lane] = 0.f;  // pad the last quad of the data part
  __syncwarp();
  return live;
}

// lv_decorate_stage + contiguous (128-bit when aligned) stores of the live part to `dst`.  The
// caller has already issued lv_decorate_zero_tail for the rest of the row.
__device__ __forceinline__ void lv_decorate_warp(const float4 a, const float4 b, int num, int coor_y, int coor_x,
                                                 const DecoCfg& d, float* st, float* __restrict__ dst, int lane) {
  const int nd = lv_decorate_stage(a, b, num, coor_y, coor_x, d, st, lane, d.C_out) * d.C_out;
  if (lv_row_vec4(d, dst)) {
    const float4* s4 = reinterpret_cast<const float4*>(st);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < ((nd + 3) >> 2); i += 32) lv_st_stream_f4(d4 + i, s4[i]);
  } else {
    for (int i = lane; i < nd; i += 32) dst[i] = st[i];
  }
  __syncwarp();
}

// Two small pillars (<= 16 points each) decorated by ONE warp at once: lanes 0-15 own pillar A,
// lanes 16-31 pillar B, lane (sub) holds slot `sub` of its pillar in `a` (zeros beyond num).  The
// sums run over the 16 lanes of a half (xor 8,4,2,1 never crosses bit 4) - the same additions, in
// the same order, as the 32-lane tree of lv_decorate_stage sees for a pillar of <= 16 points, so
// the result is bit-identical.  `st` is the half's own stage, `dst` the half's own output row;
// the caller has already issued lv_decorate_zero_tail for both rows.  Needs T > 16 (padding slots
// exist, which is what makes the z range of the RadiusHeight variant include 0).
__device__ __forceinline__ void lv_decorate_half_stage(const float4 a, int num, int coor_y, int coor_x, const DecoCfg& d,
                                                       float* st, int lane, int slot_stride) {
  const int sub = lane & 15;
  float sx = a.x, sy = a.y, sz = a.z;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  const float fn = (float)num;
  const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
  float height = 0.f;
  if (d.variant == LV_PILLAR_RADIUS_HEIGHT) {  // :387-389: the padded slots contribute z = 0
    float zmax = fmaxf(a.z, 0.f), zmin = fminf(a.z, 0.f);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
      zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    }
    height = zmax - zmin;
  }
  const float cx = __fadd_rn(__fmul_rn((float)coor_x, d.vx), d.x_off);
  const float cy = __fadd_rn(__fmul_rn((float)coor_y, d.vy), d.y_off);
  if (sub < num) lv_decorate_slot(a, mx, my, mz, cx, cy, height, d, st + sub * slot_stride);
  const int nd = num * slot_stride;
  if (sub < (((nd + 3) >> 2) << 2) - nd) st[nd + sub] = 0.f;  // pad the last quad of the data part
  __syncwarp();
}

__device__ __forceinline__ void lv_decorate_half(const float4 a, int num, int coor_y, int coor_x, const DecoCfg& d,
                                                 float* st, float* __restrict__ dst, int lane) {
  const int sub = lane & 15;
  lv_decorate_half_stage(a, num, coor_y, coor_x, d, st, lane, d.C_out);
  const int nd = num * d.C_out;
  if (lv_row_vec4(d, dst)) {
    const float4* s4 = reinterpret_cast<const float4*>(st);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = sub; i < ((nd + 3) >> 2); i += 16) lv_st_stream_f4(d4 + i, s4[i]);
  } else {
    for (int i = sub; i < nd; i += 16) dst[i] = st[i];
  }
  __syncwarp();
}

// ---- PFNLayer (second/second/pytorch/models/pointpillars.py:51-65), inference form, fused
// behind the decoration: y = W f (Linear 9 -> units, no bias), z = relu(y * scale + shift)
// (BatchNorm1d in eval mode folded to scale/shift on the host), out = max over the T slots.
// A padded slot has f = 0, so it contributes relu(shift) to the max whenever live < T.
#define LV_PFN_MAX_IN 12
#define LV_PFN_STRIDE 12   // floats per staged slot: three 128-bit broadcast reads fetch a slot
struct PfnCfg {
  const float* weight;   // (units, C_in) row-major = torch Linear.weight
  const float* scale;    // (units)
  const float* shift;    // (units)
  int units;             // multiple of 32, <= 128
};

// lane owns output channels lane + 32 j; its CIN*UJ weights live in registers
template <int CIN, int UJ>
struct PfnRegs {
  float w[CIN][UJ], sc[UJ], sh[UJ];
  __device__ __forceinline__ void load(const PfnCfg& c, int lane) {
#pragma unroll
    for (int j = 0; j < UJ; ++j) {
      const int ch = lane + 32 * j;
      sc[j] = __ldg(c.scale + ch);
      sh[j] = __ldg(c.shift + ch);
#pragma unroll
      for (int k = 0; k < CIN; ++k) w[k][j] = __ldg(c.weight + ch * CIN + k);
    }
  }
};

// `st` holds the live slots at stride LV_PFN_STRIDE (16-byte aligned)
template <int CIN, int UJ>
__device__ __forceinline__ void lv_pfn_warp(const float* st, int live, int T, const PfnRegs<CIN, UJ>& r,
                                            float* __restrict__ out, int lane) {
  float m[UJ];
#pragma unroll
  for (int j = 0; j < UJ; ++j) m[j] = live < T ? fmaxf(r.sh[j], 0.f) : 0.f;  // relu >= 0 either way
  for (int t = 0; t < live; ++t) {
    const float4* f4 = reinterpret_cast<const float4*>(st + t * LV_PFN_STRIDE);  // broadcast reads
    float f[LV_PFN_STRIDE];
    *reinterpret_cast<float4*>(f) = f4[0];
    *reinterpret_cast<float4*>(f + 4) = f4[1];
    if (CIN > 8) *reinterpret_cast<float4*>(f + 8) = f4[2];
    float y[UJ];
#pragma unroll
    for (int j = 0; j < UJ; ++j) y[j] = 0.f;
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
#pragma unroll
      for (int j = 0; j < UJ; ++j) y[j] = fmaf(f[k], r.w[k][j], y[j]);
    }
#pragma unroll
    for (int j = 0; j < UJ; ++j) m[j] = fmaxf(m[j], fmaf(y[j], r.sc[j], r.sh[j]));
  }
#pragma unroll
  for (int j = 0; j < UJ; ++j) out[lane + 32 * j] = m[j];
  __syncwarp();
}
