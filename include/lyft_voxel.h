/* lyft_voxel.h - C ABI of the B200-native lidar voxelization / BEV rasterization
 * engine (liblyftvoxel_b200.so, sm_100a only).
 *
 * The reference (jionie/Lyft-3D-Object-Detection) has no C/FFI plugin API on this
 * path; its boundary is three Python surfaces.  Each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - Every function returns int: 0 = LV_OK, negative = LV_E_*.  lv_last_error()
 *     returns a thread-local message.  Nothing aborts or throws across the ABI.
 *   - "d_" pointers are device pointers on the handle's device, "h_" pointers are
 *     host pointers.  The CALLER allocates every input and output; the library
 *     owns only its workspace inside the handle (grown on first use / when a
 *     larger problem arrives, never on a steady-state call).
 *   - All work is enqueued on the caller's stream (a cudaStream_t passed as
 *     void*); device-pointer entry points never synchronise.  Outputs are valid
 *     once the caller synchronises that stream.  *_host entry points take host
 *     buffers, do the H2D/D2H copies on the handle's stream and return after the
 *     results are in the caller's host memory.
 *   - A handle is bound to one device and is not thread-safe: one handle per
 *     (process, device, stream).
 *   - There is no CPU fallback: without a CUDA device lv_create fails with
 *     LV_E_NODEVICE.
 */
#ifndef LYFT_VOXEL_H_
#define LYFT_VOXEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_OK 0
#define LV_E_INVALID (-1)     /* bad argument (shape, null pointer, unsupported size) */
#define LV_E_CUDA (-2)        /* a CUDA runtime call failed; see lv_last_error */
#define LV_E_NOMEM (-3)       /* workspace allocation failed */
#define LV_E_UNSUPPORTED (-4) /* feature not built / not supported on this device */
#define LV_E_NODEVICE (-5)    /* no CUDA device / wrong architecture */

#define LV_ABI_VERSION 1

typedef struct lv_handle lv_handle;
typedef void* lv_stream; /* cudaStream_t */

/* ------------------------------------------------------------------ lifecycle */

int lv_abi_version(void);
const char* lv_version_string(void);
/* Number of visible CUDA devices (0 when none / no driver); never fails. */
int lv_device_count(void);
/* Creates a handle bound to `device` (must be compute capability 10.x). */
int lv_create(int device, lv_handle** out);
int lv_destroy(lv_handle* h);
/* Thread-local message of the last failing call ("" if none).  h may be NULL. */
const char* lv_last_error(lv_handle* h);
/* Bytes of device workspace currently owned by the handle. */
int64_t lv_workspace_bytes(lv_handle* h);
/* Number of kernels this handle has launched since creation (bench.py's
 * gpu_launches claim is read from here). */
int64_t lv_launch_count(lv_handle* h);
/* Pinned host memory for callers of the *_host entry points.  The allocation is MAPPED and portable: under
 * unified addressing the same pointer is valid in kernels, which is what lets lv_png_encode write its
 * files straight into host memory. */
int lv_host_alloc(size_t bytes, void** out);
int lv_host_free(void* p);
/* Tuning knobs (0 = library default). */
int lv_set_option(lv_handle* h, const char* name, int64_t value);

/* ------------------------------------------------------------------ BEV rasteriser
 *
 * Replaces, per frame, the chain
 *   LidarPointCloud.transform(car_from_sensor)      nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py:188-195
 *   create_voxel_pointcloud(points, shape, voxel_size, z_offset)
 *                                                   generating-dataset/generating_train_bev.py:84-101
 *     (car_to_voxel_coords :73-82, transform_points :64-70,
 *      create_transformation_matrix_to_voxel_space :47-62)
 *   normalize_voxel_intensities(bev, max_intensity) generating-dataset/generating_train_bev.py:103-104
 *   np.round(bev*255).astype(np.uint8)              generating-dataset/generating_train_bev.py:213
 *   BEVImageDataset concat + /255 + HWC->CHW        deeplab_v3_baseline/dataset/dataset.py:83-106
 *
 * Points: `point_stride` float32 per point (x,y,z first; 4 for (N,4) arrays, 5
 * for raw .bin sweeps).  They are grouped in `n_segments` consecutive segments
 * (sweeps); segment s covers points [h_seg_offsets[s], h_seg_offsets[s+1]) and
 * belongs to frame h_seg_frame[s] (NULL: segment s is frame s).  h_seg_tm (NULL:
 * identity) holds one row-major float64 4x4 sensor->car transform per segment,
 * applied in float64 and rounded to float32 exactly like PointCloud.transform.
 *
 * Arithmetic: u_k = fl64(fl64(m_k*p_k) + t_k) with
 * m_k = 1/voxel_size[k], t = shape/2 + (0,0,z_offset)/voxel_size; c_k = trunc(u_k)
 * (NOT floor); keep iff 0 <= c_k < shape[k]; bev[c_1, c_0, c_2] += 1.
 * shape[0] must equal shape[1].
 *
 * Outputs (any may be NULL), each n_frames x ...:
 *   d_raw   float32 (S0,S1,S2)  raw counts            (create_voxel_pointcloud)
 *   d_norm  float32 (S0,S1,S2)  clip(raw/max_int,0,1) (normalize_voxel_intensities)
 *   d_u8    uint8   (S0,S1,S2)  rint(norm*255)
 *   d_chw   float32 (S2+3,S0,S1) [u8 | map]/255 in CHW; needs d_map_u8
 *           (n_frames x (S0,S1,3) uint8) and S2 == 3.
 */
int lv_bev_rasterize(lv_handle* h, const float* d_points, int32_t point_stride,
                     int32_t n_segments, const int64_t* h_seg_offsets,
                     const int32_t* h_seg_frame, const double* h_seg_tm, int32_t n_frames,
                     const int32_t shape[3], const double voxel_size[3], double z_offset,
                     float max_intensity, float* d_raw, float* d_norm, uint8_t* d_u8,
                     const uint8_t* d_map_u8, float* d_chw, lv_stream stream);

/* Same, host buffers in and out (pinned memory from lv_host_alloc recommended).
 * Copies points H2D, runs the kernels, copies every non-NULL output D2H, and
 * returns after the handle's stream is idle. */
int lv_bev_rasterize_host(lv_handle* h, const float* h_points, int32_t point_stride,
                          int32_t n_segments, const int64_t* h_seg_offsets,
                          const int32_t* h_seg_frame, const double* h_seg_tm, int32_t n_frames,
                          const int32_t shape[3], const double voxel_size[3], double z_offset,
                          float max_intensity, float* h_raw, float* h_norm, uint8_t* h_u8,
                          const uint8_t* h_map_u8, float* h_chw);

/* Target rasterisation: draw_boxes of generating-dataset/generating_train_bev.py:127-139, the
 * class-index image written to "{token}_target.png" (:214-223).  Frame f paints boxes
 * [h_box_offsets[f], h_box_offsets[f+1]) in that order, later boxes over earlier ones:
 *   d_corners  float64 (n_boxes, 3, 4)  Box.bottom_corners() in car space (rows x, y, z; the four
 *                                       bottom corners in order around the footprint)
 *   d_colors   int32   (n_boxes)        classes.index(box.name) + 1  (1..255)
 *   d_target   uint8   (n_frames, shape[0], shape[1])  target[:, :, 0]; 0 = background.  Every byte
 *                                       is written.
 * Corners go through car_to_voxel_coords (:73-82; fp64) and np.int0 (C truncation);
 * the filled polygon is cv2.drawContours(..., -1): OpenCV's Line + scanline fill rule, restated
 * in oracle/draw_oracle.py and pinned against cv2 4.13.  Vertices are clamped to +-2^20 pixels. */
int lv_draw_boxes(lv_handle* h, const double* d_corners, const int32_t* d_colors, int32_t n_frames,
                  const int64_t* h_box_offsets, const int32_t shape[3], const double voxel_size[3],
                  double z_offset, uint8_t* d_target, lv_stream stream);

int lv_draw_boxes_host(lv_handle* h, const double* h_corners, const int32_t* h_colors,
                       int32_t n_frames, const int64_t* h_box_offsets, const int32_t shape[3],
                       const double voxel_size[3], double z_offset, uint8_t* h_target);

/* PNG encoding on the device: cv2.imwrite(path, image) of generating-dataset/generating_train_bev.py:215
 * ("{token}_input.png", (H,W,3) uint8), :224 ("{token}_target.png", (H,W) uint8) and :229, for images that
 * already live in HBM - an encoded sparse BEV image is ~16x smaller than the dense array that would cross
 * PCIe otherwise.  PNG is lossless; the contract is decode-exactness under cv2.imread(..., IMREAD_UNCHANGED)
 * (deeplab_v3_baseline/dataset/dataset.py:83-90), and the byte stream is the one stated by
 * oracle/png_oracle.py: filter type 0, ONE fixed-Huffman deflate block (run-length matches at distance 1),
 * Adler-32, CRC-32, a single IDAT.
 *   d_images  uint8 (n_frames, height, width, channels)  channels 1 (grey) or 3; width*channels < 4096
 *   swap_rb   non-zero: a 3-channel image is in OpenCV's B,G,R order and is written R,G,B (= cv2.imwrite)
 *   out       frame f's file starts at out + f*out_stride.  DEVICE memory or MAPPED PINNED HOST memory
 *             (lv_host_alloc): in the second case the file bytes cross PCIe as they are produced and
 *             nothing else is transferred.  Bytes of a slot beyond the file size are not written.
 *   sizes     int32 (n_frames), device or mapped host: file size of frame f, or MINUS the size it needs
 *             when out_stride is too small (lv_png_max_bytes is always enough). */
int64_t lv_png_max_bytes(int32_t height, int32_t width, int32_t channels);
int lv_png_encode(lv_handle* h, const uint8_t* d_images, int32_t n_frames, int32_t height, int32_t width,
                  int32_t channels, int32_t swap_rb, uint8_t* out, int64_t out_stride, int32_t* sizes,
                  lv_stream stream);
/* Host images in, host files out (a caller that holds numpy images: the cv2.imwrite / cv2.imencode call). */
int lv_png_encode_host(lv_handle* h, const uint8_t* h_images, int32_t n_frames, int32_t height, int32_t width,
                       int32_t channels, int32_t swap_rb, uint8_t* h_out, int64_t out_stride, int32_t* h_sizes);

/* normalize_voxel_intensities alone (generating_train_bev.py:103-104) on a
 * device array of n float32: out = clip(in / max_intensity, 0, 1). */
int lv_bev_normalize(lv_handle* h, const float* d_in, int64_t n, float max_intensity,
                     float* d_out, lv_stream stream);

/* transform_points (generating-dataset/generating_train_bev.py:64-70) and
 * car_to_voxel_coords (:73-82): out(3,N) float64 = (tm . [x;y;z;1])[:3], accumulated
 * in k order with FMA from 0 (what the BLAS dgemm the reference calls does; for
 * the diagonal voxel-space matrix this is the un-fused multiply-then-add of
 * SURVEY.md Appendix A.1, bit for bit).  d_points is (N, point_stride) float32
 * rows, h_tm16 a row-major float64 4x4 on the host, d_out 3 planes of N doubles. */
int lv_transform_points(lv_handle* h, const float* d_points, int32_t point_stride, int64_t n,
                        const double* h_tm16, double* d_out, lv_stream stream);

/* ------------------------------------------------------------------ multi-sweep ingest
 *
 * The step in front of VoxelGenerator.generate: n_sweeps raw sweeps (5 float32 per point:
 * x, y, z, intensity, ring - LidarPointCloud.from_file, data_classes.py:269-284), back to back in
 * d_raw, become one cloud in the key frame.
 *   LV_INGEST_SECOND  second/second/data/nuscenes_dataset.py:196-223: intensity /= 255; sweeps with a
 *     transform: xyz = xyz @ R.T (float64, stored to float32) then xyz += T (float64, stored to
 *     float32); column 4 := time lag.  out_cols 4 -> [x,y,z,lag] (the reference's column select
 *     [0,1,2,4]), 5 -> [x,y,z,intensity/255,lag].
 *   LV_INGEST_DEVKIT  lyft_dataset_sdk/utils/data_classes.py:99-137: xyz = (M . [x;y;z;1])[:3] in
 *     float64, stored once; remove_close(close_radius) turns rows with |x| < r and |y| < r into NaN
 *     rows (dropped by every downstream kernel; close_radius < 0 disables it).  out_cols 4 ->
 *     [x,y,z,intensity], 5 -> [.., time lag].
 * h_sweep_tm: n_sweeps row-major float64 4x4 (SECOND reads R = [:3,:3], T = [:3,3]); NULL = no
 * transform at all.  h_has_tm: NULL or n_sweeps flags (0 = leave the sweep untouched, as the
 * reference does for the key sweep).  h_time_lag: n_sweeps float32 (ts - sweep_ts). */
#define LV_INGEST_SECOND 0
#define LV_INGEST_DEVKIT 1
int lv_ingest_sweeps(lv_handle* h, const float* d_raw, int32_t n_sweeps,
                     const int64_t* h_sweep_offsets, const double* h_sweep_tm,
                     const uint8_t* h_has_tm, const float* h_time_lag, int32_t mode,
                     float close_radius, int32_t out_cols, float* d_out, lv_stream stream);

/* ------------------------------------------------------------------ hard voxelizer
 *
 * Replaces spconv.utils.VoxelGeneratorV2.generate / generate_multi_gpu as called at
 *   second/second/builder/voxel_builder.py:23-32      (constructor arguments)
 *   second/second/data/preprocess.py:299-317           (generate, generate_multi_gpu)
 *   second/second/inference.py:67-79                   (single-sample inference)
 * and the legacy tuple API points_to_voxel (second/second/kittiviewer/viewer.py:389-396).
 * Algorithm: the in-tree sibling second/second/utils/simplevis.py:9-61:
 *   grid = rint((hi-lo)/vs) in float32; per point in index order
 *   c_k = floorf((p_k - lo_k)/vs_k), reject unless 0 <= c_k < grid_k; coordinate
 *   stored (z,y,x); voxel ids in order of first appearance, at most max_voxels;
 *   a point is stored in slot num[id] iff num[id] < max_points.
 */
#define LV_OVERFLOW_CONTINUE 0 /* spconv >= 1.1: skip the point, keep scanning       */
#define LV_OVERFLOW_BREAK 1    /* simplevis.py:48-49: stop at the first overflowing point */

typedef struct lv_voxel_config {
  float voxel_size[3];       /* x, y, z */
  float coors_range[6];      /* xmin ymin zmin xmax ymax zmax */
  int32_t max_points;        /* T: max points per voxel */
  int32_t max_voxels;        /* V: max voxels per frame */
  int32_t num_features;      /* C: float32 per point (>= 3) */
  int32_t overflow_mode;     /* LV_OVERFLOW_* */
  int32_t zero_tail;         /* 1: rows [voxel_num, max_voxels) are zero-filled
                                (generate_multi_gpu padding, preprocess.py:311-317);
                                0: they are left untouched (caller slices) */
} lv_voxel_config;

/* grid_size (x,y,z) for a config; pure host arithmetic. */
int lv_voxel_grid_size(const lv_voxel_config* cfg, int32_t grid_xyz[3]);

/* Voxelizes n_frames independent clouds.  Frame f covers points
 * [h_frame_offsets[f], h_frame_offsets[f+1]) of d_points (row-major, C floats).
 * Outputs are padded per frame:
 *   d_voxels     float32 (n_frames, V, T, C)
 *   d_coords     int32   (n_frames, V, 3)   zyx
 *   d_num_points int32   (n_frames, V)
 *   d_voxel_num  int32   (n_frames)
 */
int lv_voxelize(lv_handle* h, const lv_voxel_config* cfg, const float* d_points,
                int32_t n_frames, const int64_t* h_frame_offsets, float* d_voxels,
                int32_t* d_coords, int32_t* d_num_points, int32_t* d_voxel_num,
                lv_stream stream);

/* Same voxelization, batch-assembled on the device the way merge_second_batch does
 * on the host (second/second/data/preprocess.py:21-55): the voxels of all frames
 * back to back, coordinates with the batch index prepended.
 *   d_voxels        float32 (capacity_rows, T, C)
 *   d_coords4       int32   (capacity_rows, 4)   [batch, z, y, x]   (16-byte aligned)
 *   d_num_points    int32   (capacity_rows)
 *   d_voxel_num     int32   (n_frames)
 *   d_voxel_offsets int64   (n_frames + 1)  first row of every frame; [n_frames] = total rows.
 * Rows beyond capacity_rows are dropped: the caller checks d_voxel_offsets[n_frames]. */
int lv_voxelize_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points,
                       int32_t n_frames, const int64_t* h_frame_offsets, int64_t capacity_rows,
                       float* d_voxels, int32_t* d_coords4, int32_t* d_num_points,
                       int32_t* d_voxel_num, int64_t* d_voxel_offsets, lv_stream stream);

/* The un-padding VoxelNet.forward does for multi-GPU batches
 * (second/second/pytorch/models/voxelnet.py:346-358): padded (B,V,T,C) voxels, (B,V) num_points,
 * (B,V,coor_cols) coordinates (merge_second_batch_multigpu, preprocess.py:60-88) and num_voxels (B)
 * become the concatenated (sum, T, C) / (sum) / (sum, coor_cols) lists; *d_out_total = sum.
 * Outputs need room for B*V rows (or the known total). */
int lv_unpad_batch(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                   const int32_t* d_coors, const int32_t* d_num_voxels, int32_t batch_size,
                   int32_t max_voxels, int32_t max_points, int32_t num_features, int32_t coor_cols,
                   float* d_out_voxels, int32_t* d_out_num_points, int32_t* d_out_coors,
                   int64_t* d_out_total, lv_stream stream);

/* lv_voxelize_concat followed by SimpleVoxel.forward (second/second/pytorch/models/voxel_encoder.py:219-225)
 * without writing the (rows, T, C) voxels: d_mean (capacity_rows, num_features_out) float32 holds
 * sum_t voxels[p, t, c] / num_points[p] for the leading num_features_out (1..4) channels.  Needs
 * num_features == 4.  The mean VFE of the 0.05 m SECOND configs. */
int lv_voxelize_mean_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points,
                            int32_t n_frames, const int64_t* h_frame_offsets, int64_t capacity_rows,
                            int32_t num_features_out, float* d_mean, int32_t* d_coords4,
                            int32_t* d_num_points, int32_t* d_voxel_num, int64_t* d_voxel_offsets,
                            lv_stream stream);

int lv_voxelize_host(lv_handle* h, const lv_voxel_config* cfg, const float* h_points,
                     int32_t n_frames, const int64_t* h_frame_offsets, float* h_voxels,
                     int32_t* h_coords, int32_t* h_num_points, int32_t* h_voxel_num);

/* Block-filtering variant: VoxelGeneratorV2(block_filtering=True, block_factor, block_size,
 * height_threshold) - kwargs at second/second/builder/voxel_builder.py:28-31, fields of
 * second/second/protos/voxel_generator.proto:10-13, used by
 * second/second/configs/nuscenes/all.fhd.config:9-12.  PARITY UNPINNED: the rule lives in spconv 1.x
 * (points_to_voxel_3d_with_filtering, source not in the reference tree); restated in
 * oracle/voxel_oracle.c: every stored point updates the z-min/z-max of its
 * (y / block_factor, x / block_factor) block; a voxel survives iff, over the block_size x block_size
 * window of blocks around its own, height = max - min satisfies
 * height > height_threshold && height < height_high_threshold (+inf: the rule before spconv 1.2);
 * survivors keep their first-come order.  The xy grid must be divisible by block_factor. */
typedef struct lv_block_filter {
  int32_t block_factor;
  int32_t block_size;
  float height_threshold;
  float height_high_threshold;
} lv_block_filter;

/* The filter alone, on padded per-frame voxelizer output (layouts of lv_voxelize).  Outputs must not
 * alias the inputs.  d_out_voxel_num (n_frames) = survivors per frame; rows beyond it are zero-filled
 * iff cfg->zero_tail.  d_out_mask (optional, (n_frames, V) int32) = keep flag of every unfiltered voxel. */
int lv_voxel_block_filter(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                          int32_t n_frames, const float* d_voxels, const int32_t* d_coords,
                          const int32_t* d_num_points, const int32_t* d_voxel_num, float* d_out_voxels,
                          int32_t* d_out_coords, int32_t* d_out_num_points, int32_t* d_out_voxel_num,
                          int32_t* d_out_mask, lv_stream stream);

/* lv_voxelize (overflow_mode must be LV_OVERFLOW_CONTINUE) followed by the filter; the unfiltered
 * voxels live in the handle's workspace. */
int lv_voxelize_filtered(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                         const float* d_points, int32_t n_frames, const int64_t* h_frame_offsets,
                         float* d_voxels, int32_t* d_coords, int32_t* d_num_points,
                         int32_t* d_voxel_num, int32_t* d_mask, lv_stream stream);

int lv_voxelize_filtered_host(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                              const float* h_points, int32_t n_frames, const int64_t* h_frame_offsets,
                              float* h_voxels, int32_t* h_coords, int32_t* h_num_points,
                              int32_t* h_voxel_num);

/* The host-buffer call in two phases, for callers that want arrays of exactly voxel_num rows
 * (VoxelGeneratorV2.generate slices to voxel_num anyway, second/second/data/preprocess.py:305-310):
 * _begin copies the points in, runs the kernels (block filter when flt != NULL) and returns the voxel
 * count of every frame; the results stay in the handle's staging buffers until the next *_host call.
 * _fetch copies the first `rows` rows of one frame into caller arrays (rows, T, C) / (rows, 3) / (rows). */
int lv_voxelize_host_begin(lv_handle* h, const lv_voxel_config* cfg, const lv_block_filter* flt,
                           const float* h_points, int32_t n_frames, const int64_t* h_frame_offsets,
                           int32_t* h_voxel_num);
int lv_voxelize_host_fetch(lv_handle* h, int32_t frame, int32_t rows, float* h_voxels,
                           int32_t* h_coords, int32_t* h_num_points);

/* ------------------------------------------------------------------ PointPillars
 *
 * lv_pillar_decorate replaces the decoration part of
 *   PillarFeatureNet.forward             second/second/pytorch/models/pointpillars.py:203-231
 *   PillarFeatureNetOld.forward          :117-145  (x,y aliasing kept)
 *   PillarFeatureNetRadius.forward       :290-319
 *   PillarFeatureNetRadiusHeight.forward :378-411
 * including get_paddings_indicator (second/second/pytorch/models/voxel_encoder.py:27-48).
 * Input d_voxels (P,T,C) float32, d_num_points (P) int32, d_coors (P,4) int32
 * [batch,z,y,x]; output (P,T,C_out) float32, C_out = lv_pillar_out_channels().
 * vx, vy, x_offset, y_offset are the module's scalars (pointpillars.py:198-201),
 * already rounded to float32 the way torch rounds a Python scalar.
 */
#define LV_PILLAR_PFN 0
#define LV_PILLAR_OLD 1
#define LV_PILLAR_RADIUS 2
#define LV_PILLAR_RADIUS_HEIGHT 3

int lv_pillar_out_channels(int32_t num_features, int32_t variant, int32_t with_distance);

int lv_pillar_decorate(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                       const int32_t* d_coors, int64_t n_pillars, int32_t max_points,
                       int32_t num_features, float vx, float vy, float x_offset,
                       float y_offset, int32_t variant, int32_t with_distance, float* d_out,
                       lv_stream stream);

/* Voxelization fused with the decoration (north_star: "pillar decoration is fused into
 * the gather"): lv_voxelize_concat followed by lv_pillar_decorate, without ever writing
 * the (P,T,C) voxel tensor.  d_decorated is (capacity_rows, T, C_out) float32; the other
 * outputs are those of lv_voxelize_concat.  Needs num_features == 4 and max_points <= 64.
 * Replaces second/second/data/preprocess.py:299-317 + :21-55 and
 * second/second/pytorch/models/pointpillars.py:203-231 in one call. */
int lv_pillarize_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points,
                        int32_t n_frames, const int64_t* h_frame_offsets, int64_t capacity_rows,
                        float vx, float vy, float x_offset, float y_offset, int32_t variant,
                        int32_t with_distance, float* d_decorated, int32_t* d_coords4,
                        int32_t* d_num_points, int32_t* d_voxel_num, int64_t* d_voxel_offsets,
                        lv_stream stream);

/* PFNLayer in inference form (second/second/pytorch/models/pointpillars.py:51-65, last layer):
 * Linear(C_out -> units, weight (units, C_out) row-major, no bias) -> BatchNorm1d in eval mode
 * folded by the caller to y*scale + shift (scale = gamma/sqrt(var + eps), shift = beta - mean*scale;
 * scale = 1, shift = bias when the layer has no norm) -> ReLU -> max over the T slots (a padded
 * slot contributes relu(shift)).  Fused behind the decoration: the (P,T,C_out) tensor of
 * lv_pillar_decorate is never materialised.  d_out is (P, units) float32;
 * units in {32, 64, 128}; needs num_features == 4 and max_points <= 64. */
int lv_pillar_pfn(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                  const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                  float vx, float vy, float x_offset, float y_offset, int32_t variant,
                  int32_t with_distance, const float* d_weight, const float* d_scale,
                  const float* d_shift, int32_t units, float* d_out, lv_stream stream);

/* PFNLayer in TRAINING form (second/second/pytorch/models/pointpillars.py:51-65 with BatchNorm1d batch
 * statistics, and its backward) behind the fused decoration, in two device passes that never write the
 * (P,T,C_out) decorated tensor or the (P,T,units) activations:
 *   lv_pillar_pfn_moments   d_moments (float64, zeroed by the call) = [S1 (C_out) | upper triangle of
 *                           M = sum f f^T row by row (C_out (C_out+1)/2)], sums over the live slots of the
 *                           decorated features f.  The Linear has no bias and padded slots are zero, so the
 *                           batch mean / variance of y = W f over all N = P*T slots are W S1 / N and
 *                           diag(W M W^T) / N - mean^2; the caller turns them into the scale / shift of
 *                           lv_pillar_pfn (the forward pass proper).
 *   lv_pillar_pfn_backward  d_acc (float64 (units, 2 + C_out), zeroed by the call) = per channel
 *                           [dbeta = sum g | dgamma = sum g * yhat | A = sum g * f at the slot that won the
 *                           max], over the (pillar, channel) pairs with a positive maximum; g = d_grad_out
 *                           (P, units).  d_mean / d_invstd are the batch statistics of the forward.
 * Same restrictions as lv_pillar_pfn (4 features per point, max_points <= 64, units in {32, 64, 128}). */
int lv_pillar_pfn_moments(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                          const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                          float vx, float vy, float x_offset, float y_offset, int32_t variant,
                          int32_t with_distance, double* d_moments, lv_stream stream);
int lv_pillar_pfn_backward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                           const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                           float vx, float vy, float x_offset, float y_offset, int32_t variant,
                           int32_t with_distance, const float* d_weight, const float* d_scale,
                           const float* d_shift, const float* d_mean, const float* d_invstd, int32_t units,
                           const float* d_grad_out, double* d_acc, lv_stream stream);

/* The two passes as complete operations (what lyft3d_b200.pointpillars._FusedPfnTrain binds): the 64-number ends
 * - batch statistics from the moments; dW, dgamma, dbeta from the accumulators - run on the device too.
 *   forward   d_out (P, units) = max over T of relu(BatchNorm_train(W f)); d_stats float [5][units] = scale
 *             (gamma * invstd), shift (beta - mean * scale), batch mean, invstd, biased batch variance (the caller
 *             updates running_mean / running_var from rows 2 and 4); d_moments float64 [C_out + C_out (C_out+1)/2]
 *             is kept by the caller for the backward.
 *   backward  d_dweight (units, C_out), d_dgamma (units), d_dbeta (units) for d_grad_out (P, units). */
int lv_pillar_pfn_train_forward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                                const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                                float vx, float vy, float x_offset, float y_offset, int32_t variant,
                                int32_t with_distance, const float* d_weight, const float* d_gamma,
                                const float* d_beta, double eps, int32_t units, float* d_out, float* d_stats,
                                double* d_moments, lv_stream stream);
int lv_pillar_pfn_train_backward(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                                 const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                                 float vx, float vy, float x_offset, float y_offset, int32_t variant,
                                 int32_t with_distance, const float* d_weight, const float* d_gamma, double eps,
                                 const float* d_stats, const double* d_moments, int32_t units,
                                 const float* d_grad_out, float* d_dweight, float* d_dgamma, float* d_dbeta,
                                 lv_stream stream);

/* Backward of PointPillarsScatter (second/second/pytorch/models/pointpillars.py:444-476 under autograd):
 * d_grad_feats (P, channels) [p, c] = d_grad_canvas (B, channels, ny, nx) [b, c, y, x] for coords[p] = (b, ., y, x); rows
 * whose coordinates lie outside the canvas get zeros.  half != 0: both tensors are float16. */
int lv_pillar_scatter_backward(lv_handle* h, const void* d_grad_canvas, const int32_t* d_coords, int64_t n_pillars,
                               int32_t channels, int32_t batch_size, int32_t ny, int32_t nx, int32_t half,
                               void* d_grad_feats, lv_stream stream);

/* lv_voxelize_concat + lv_pillar_pfn in one call: points in, (capacity_rows, units) pillar
 * features out (units == 64), ready for lv_pillar_scatter.  Replaces preprocess.py:299-317 +
 * :21-55, pointpillars.py:203-231 and :51-65 without writing voxels or decorated points. */
int lv_pillarize_pfn_concat(lv_handle* h, const lv_voxel_config* cfg, const float* d_points,
                            int32_t n_frames, const int64_t* h_frame_offsets, int64_t capacity_rows,
                            float vx, float vy, float x_offset, float y_offset, int32_t variant,
                            int32_t with_distance, const float* d_weight, const float* d_scale,
                            const float* d_shift, int32_t units, float* d_features,
                            int32_t* d_coords4, int32_t* d_num_points, int32_t* d_voxel_num,
                            int64_t* d_voxel_offsets, lv_stream stream);

/* lv_pillar_scatter replaces PointPillarsScatter.forward
 * (second/second/pytorch/models/pointpillars.py:444-476):
 * canvas[b, c, y, x] = feats[p, c] for the pillar p with coords[p] = (b, ., y, x),
 * 0 elsewhere.  d_feats (P,C) float32, d_coords (P,4) int32, canvas (B,C,ny,nx).
 * Every canvas element is written exactly once (no separate memset). */
int lv_pillar_scatter(lv_handle* h, const float* d_feats, const int32_t* d_coords,
                      int64_t n_pillars, int32_t channels, int32_t batch_size, int32_t ny,
                      int32_t nx, float* d_canvas, lv_stream stream);

/* Same, with the pillar count read on the device (*d_n_pillars, e.g. d_voxel_offsets[n_frames] of
 * lv_voxelize_concat; rows at and beyond min(*d_n_pillars, max_pillars) are ignored): the host
 * never waits for the voxelizer, so consecutive frames pipeline across streams. */
int lv_pillar_scatter_dev(lv_handle* h, const float* d_feats, const int32_t* d_coords,
                          const int64_t* d_n_pillars, int64_t max_pillars, int32_t channels,
                          int32_t batch_size, int32_t ny, int32_t nx, float* d_canvas,
                          lv_stream stream);

/* Half-precision forms (apex O2: second/second/pytorch/train.py:34-47 casts "voxels" to half,
 * second/configs/nuscenes/all.pp.mida.config:348; the canvas follows voxel_features.dtype,
 * pointpillars.py:449-452).  float16 arrays are passed as uint16_t*.
 * lv_pillar_decorate_half reproduces torch's half arithmetic - every op computes in float32 and rounds
 * its result to half - statement by statement (pointpillars.py:203-231 and the variants): needs
 * num_features == 4 and max_points <= 64.  lv_pillar_scatter_half writes the (B,C,ny,nx) float16 canvas
 * in one pass, every element exactly once, bit-exact. */
int lv_pillar_decorate_half(lv_handle* h, const uint16_t* d_voxels, const int32_t* d_num_points,
                            const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                            float vx, float vy, float x_offset, float y_offset, int32_t variant,
                            int32_t with_distance, uint16_t* d_out, lv_stream stream);
int lv_pillar_scatter_half(lv_handle* h, const uint16_t* d_feats, const int32_t* d_coords, int64_t n_pillars,
                           int32_t channels, int32_t batch_size, int32_t ny, int32_t nx, uint16_t* d_canvas,
                           lv_stream stream);

/* SimpleVoxel.forward (second/second/pytorch/models/voxel_encoder.py:219-225):
 * out[p, c] = sum_t voxels[p, t, c] / num[p] for c < num_features_out. */
int lv_voxel_mean(lv_handle* h, const float* d_voxels, const int32_t* d_num_points,
                  int64_t n_voxels, int32_t max_points, int32_t num_features,
                  int32_t num_features_out, float* d_out, lv_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* LYFT_VOXEL_H_ */
