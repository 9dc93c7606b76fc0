// Internal header shared by the kernels of liblyftvoxel_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/lyft_voxel.h"

#define LV_NUM_SMS_B200 148

// ------------------------------------------------------------------ errors
void lv_set_error(const char* fmt, ...);

#define LV_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      lv_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LV_E_CUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define LV_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      lv_set_error(__VA_ARGS__);  \
      return LV_E_INVALID;        \
    }                             \
  } while (0)

#define LV_CHECK(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != LV_OK) return _r; \
  } while (0)

// ------------------------------------------------------------------ workspace
// A growable device buffer owned by the handle.  ensure() reallocates only when
// the request exceeds the capacity (never on a steady-state call); `fill` is the
// byte the fresh allocation is initialised with (maps that must start "empty").
struct lv_buffer {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, cudaStream_t stream, int fill_byte = -1, bool* grew = nullptr);
  void release();
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(ptr); }
};

// Host-side cache of a small table mirrored on the device (segment offsets,
// transforms): re-uploaded only when the host content changes.
struct lv_mirror {
  lv_buffer dev;
  std::vector<unsigned char> host;
  // returns device pointer (uploaded on `stream` if the content changed)
  int sync(const void* src, size_t bytes, cudaStream_t stream, const void** out);
};

struct lv_handle {
  int device = 0;
  int num_sms = LV_NUM_SMS_B200;
  int64_t launches = 0;
  cudaStream_t own_stream = nullptr;  // used by the *_host entry points

  // options
  int64_t bev_frames_in_flight = 0;   // 0 = auto
  int64_t vox_dense_map_limit_bytes = 0;
  int64_t disable_tma = 0;            // 1: no kernel stages point tiles by TMA (A/B measurements)
  int64_t bev_tma = 0;                // 1: bev_hist_kernel stages its tiles by TMA.  Off by default: measured
                                      // slower than plain loads (tools/ab_tma.py: 0.176 vs 0.154 ms stride 4,
                                      // 0.203 vs 0.158 ms stride 5, 128 frames) - the kernel is bound by the
                                      // L2 atomics, not by the loads

  int64_t bev_u16 = 0;                // 1: 16-bit BEV counts, two per word, whenever every frame of the call has < 65,536 points (128 frames
                                      // of 336x336x3: 87 MB instead of 173 MB, L2-resident).  Off by default: measured SLOWER, 0.153 vs
                                      // 0.130 ms (histogram 73 vs 55 us: the three z-cells of an (x, y) column now share words, and
                                      // same-word atomics serialise in L2; the finalize gains nothing, 71 vs 65 us)
  int64_t bev_fused_zero = 0;         // 1: bev_hist_kernel streams the zeros of the dense outputs beside its atomics (pass A of the
                                      // finalize).  Off by default: BEV stage 0.127 vs 0.129 ms, but the pipelined step gets SLOWER
                                      // (1.655 vs 1.640 ms) - the histogram slows by what the finalize saves and overlaps its
                                      // neighbours worse
  int64_t vox_frame_kernel = 0;       // 1: small grids run K1-K5 in one cluster per frame (vx_frame_kernel, distributed shared
                                      // memory map).  Off by default: measured slower than the five kernels (0.71 vs 0.67 ms per
                                      // 128 pillar frames) - scattered 4-byte DSMEM accesses run at ~0.5 per cycle and SM
  int64_t vox_list_path = 0;          // 1: grids with a small dense map run the three-kernel list path (lv_voxel_list.cuh).  Off by
                                      // default: bit-identical but measured slower (1.03 vs 0.67 ms per 128 pillar frames)
  int64_t vox_fused_prologue = 0;     // 1: frames of <= 64 chunks run K1-K5 as ONE kernel (vx_fused_kernel: per-frame arrival counters,
                                      // look-back and polling instead of kernel boundaries).  Off by default: bit-identical, but measured
                                      // slower (0.724 vs 0.671 ms per 128 pillar frames; see the kernel's header in lv_voxel.cu)
  int64_t vox_generic_rows = 0;       // 1: the fused decoration always runs the generic row writer (A/B against the T = 60 / 9-channel one)
  int64_t vox_small_bins = 0;         // -1: never shrink the bins of K6 for calls with few frames (A/B)
  int64_t vox_two_level_scan = 0;     // K4: 0 = automatic (two levels when a frame's [bin][chunk] table exceeds 16 k counters),
                                      // 1 = always, -1 = never
  int64_t vox_hash_map = 0;           // first[] as an open-addressing table sized by the points instead of a dense map sized by the
                                      // grid: 0 = automatic (grids whose dense map exceeds 48 MB per frame), 1 = always, -1 = never
  int64_t vox_rows_waves = 0;         // CTAs of vl_rows_kernel per resident slot (0 = 4)
  int64_t canvas_variant = 0;         // 0 = auto (pillar_canvas_q_kernel when the shape allows), 1 = pillar_canvas_kernel (A/B)

  // BEV
  lv_buffer bev_counts;               // u32 [frames_in_flight][cells], kept all-zero between calls
  lv_buffer bev_dirty;                // 1 bit per 4 counts, kept all-zero between calls
  lv_mirror bev_seg_offsets, bev_seg_frame, bev_seg_tm;
  lv_buffer bev_stage_points, bev_stage_out[5], bev_stage_map;  // *_host staging
  lv_mirror draw_offsets;             // target rasterisation: box offsets per frame
  lv_buffer draw_recs, draw_stage[3]; // per-box line / scanline-edge records; *_host staging

  // voxelizer
  lv_buffer vox_map;                  // i32 [frames_in_flight][grid cells], kept all-INT_MAX between calls
  lv_buffer vox_cell, vox_aux, vox_keys[2], vox_vals[2], vox_hist, vox_chunk, vox_frame_state;
  lv_buffer vox_row_base, vox_vrec, vox_bin_start, vox_fused;
  lv_mirror vox_frame_offsets, vox_chunk_table, vox_chunk_frame;
  lv_buffer vox_stage_points, vox_stage_out[4];
  int32_t vox_host_frames = 0;        // frames held by the staging buffers since the last lv_voxelize_host_begin
  lv_voxel_config vox_host_cfg;       // its configuration
  lv_buffer flt_ranges, flt_dst, flt_tmp[4];   // block filter: z ranges per block, destination rows, unfiltered voxels

  // pillar
  lv_buffer pil_map;                  // i32 [B][ny*nx]

  // multi-sweep ingest
  lv_mirror ing_offsets, ing_tm, ing_lag, ing_has;

  lv_buffer pfn_acc;                  // training-mode PFN backward: float64 (units, 2 + C_in) accumulators

  // PNG encoder: look-back descriptors and records per CTA; *_host staging (images, files, sizes)
  lv_buffer png_state, png_rec, png_stage[3];
};

// every device buffer a handle owns (for lv_destroy / lv_workspace_bytes)
inline std::vector<lv_buffer*> lv_all_buffers(lv_handle* h) {
  return {&h->bev_counts, &h->bev_dirty, &h->bev_seg_offsets.dev, &h->bev_seg_frame.dev, &h->bev_seg_tm.dev, &h->bev_stage_points,
          &h->bev_stage_out[0], &h->bev_stage_out[1], &h->bev_stage_out[2], &h->bev_stage_out[3], &h->bev_stage_out[4],
          &h->bev_stage_map, &h->draw_offsets.dev, &h->draw_recs, &h->draw_stage[0], &h->draw_stage[1], &h->draw_stage[2],
          &h->vox_map, &h->vox_cell, &h->vox_aux, &h->vox_keys[0], &h->vox_keys[1], &h->vox_vals[0],
          &h->vox_vals[1], &h->vox_hist, &h->vox_chunk, &h->vox_frame_state, &h->vox_row_base, &h->vox_vrec, &h->vox_bin_start, &h->vox_fused,
          &h->vox_frame_offsets.dev, &h->vox_chunk_table.dev, &h->vox_chunk_frame.dev, &h->vox_stage_points, &h->vox_stage_out[0],
          &h->vox_stage_out[1], &h->vox_stage_out[2], &h->vox_stage_out[3], &h->flt_ranges, &h->flt_dst, &h->flt_tmp[0], &h->flt_tmp[1],
          &h->flt_tmp[2], &h->flt_tmp[3], &h->pil_map, &h->ing_offsets.dev,
          &h->ing_tm.dev, &h->ing_lag.dev, &h->ing_has.dev, &h->pfn_acc, &h->png_state, &h->png_rec, &h->png_stage[0], &h->png_stage[1],
          &h->png_stage[2]};
}

inline size_t lv_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t lv_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define LV_LAUNCH_CHECK(h)                 \
  do {                                     \
    (h)->launches += 1;                    \
    LV_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float4 lv_ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void lv_st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: stages tiles of point rows
// in shared memory.  Source and destination 16-byte aligned, size a multiple of 16 bytes.
__device__ __forceinline__ uint32_t lv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lv_mbar_init(uint64_t* bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lv_smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy; follow with __syncthreads()
__device__ __forceinline__ void lv_mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void lv_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lv_tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   lv_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(lv_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void lv_mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LV_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LV_DONE_%=;\n"
      "bra LV_WAIT_%=;\n"
      "LV_DONE_%=:\n"
      "}\n" ::"r"(lv_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void lv_prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned lv_lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
#endif
