/* Hard-voxelizer oracle (plain C, CPU).  TEST INFRASTRUCTURE - see oracle/__init__.py.
 *
 * PARITY UNPINNED at the spconv boundary: the reference calls
 * spconv.utils.VoxelGeneratorV2 (second/second/builder/voxel_builder.py:3,23-32),
 * an un-vendored, un-pinned third-party package whose source is not under
 * /root/reference.  This file restates the algorithm of its in-tree sibling
 *     second/second/utils/simplevis.py:9-61  (_points_to_bevmap_reverse_kernel)
 * - same coordinate rule (:38-42), same first-come voxel ids (:45-51), same
 * `break` on max_voxels (:48-49) - extended with the output contract of the call
 * sites (second/second/data/preprocess.py:305-317): voxels (V,T,C) zero padded,
 * coordinates (V,3) zyx int32, num_points_per_voxel (V,) int32.
 *
 * overflow_mode 1 = `break`   (in-tree rule, simplevis.py:48-49)
 * overflow_mode 0 = `continue` (spconv >= 1.1 rule, SURVEY.md F6: a point that
 *                   would open voxel number max_voxels is skipped, later points
 *                   still fill voxels that already exist)
 *
 * All arithmetic is float32: c = floorf((p - lo) / vs) with an IEEE divide, the
 * bounds test is done on the float before the int cast (simplevis.py:38-40).
 * Build: gcc -O2 -fPIC -shared (never -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* grid_size = round_half_even((hi - lo) / vs) in float32 -> int32, xyz order.
 * simplevis.py:27-30 (np.round on a float32 array). */
void lvo_grid_size(const float* voxel_size, const float* coors_range, int32_t* grid_xyz) {
  for (int j = 0; j < 3; ++j) {
    volatile float span = coors_range[3 + j] - coors_range[j];
    volatile float g = span / voxel_size[j];
    grid_xyz[j] = (int32_t)rintf(g);
  }
}

/* Returns voxel_num.  coor_to_voxelidx is a dense int32 map of D*H*W cells that
 * must hold -1 everywhere on entry; every touched cell is reset to -1 on exit
 * (what spconv does), so it can be reused.
 * voxels / coors / num_points must be zero-filled by the caller
 * (max_voxels*T*C, max_voxels*3, max_voxels).
 * kept_points (may be NULL) receives the number of points consumed into voxels
 * before caps (the "break keeps 80,856 / continue keeps 118,024" probe of F6). */
int32_t lvo_points_to_voxel(const float* points, int64_t n, int32_t num_features,
                            const float* voxel_size, const float* coors_range,
                            int32_t max_points, int32_t max_voxels, int32_t overflow_mode,
                            int32_t* coor_to_voxelidx, float* voxels, int32_t* coors,
                            int32_t* num_points_per_voxel, int64_t* kept_points) {
  int32_t grid[3];
  lvo_grid_size(voxel_size, coors_range, grid);
  const int64_t H = grid[1], W = grid[0];
  int32_t voxel_num = 0;
  int64_t kept = 0;
  int32_t coor[3];
  for (int64_t i = 0; i < n; ++i) {
    const float* p = points + i * num_features;
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      volatile float d = p[j] - coors_range[j];
      volatile float q = d / voxel_size[j];
      float c = floorf(q);
      if (!(c >= 0.0f) || c >= (float)grid[j]) { /* NaN fails too */
        failed = 1;
        break;
      }
      coor[2 - j] = (int32_t)c;
    }
    if (failed) continue;
    int64_t cell = ((int64_t)coor[0] * H + coor[1]) * W + coor[2];
    int32_t voxelidx = coor_to_voxelidx[cell];
    if (voxelidx == -1) {
      voxelidx = voxel_num;
      if (voxel_num >= max_voxels) {
        if (overflow_mode == 1) break;
        continue;
      }
      voxel_num += 1;
      coor_to_voxelidx[cell] = voxelidx;
      coors[voxelidx * 3 + 0] = coor[0];
      coors[voxelidx * 3 + 1] = coor[1];
      coors[voxelidx * 3 + 2] = coor[2];
    }
    ++kept;
    int32_t num = num_points_per_voxel[voxelidx];
    if (num < max_points) {
      memcpy(voxels + ((int64_t)voxelidx * max_points + num) * num_features, p,
             sizeof(float) * (size_t)num_features);
      num_points_per_voxel[voxelidx] = num + 1;
    }
  }
  for (int32_t v = 0; v < voxel_num; ++v) {
    int64_t cell = ((int64_t)coors[v * 3] * H + coors[v * 3 + 1]) * W + coors[v * 3 + 2];
    coor_to_voxelidx[cell] = -1;
  }
  if (kept_points) *kept_points = kept;
  return voxel_num;
}

/* BEV histogram (same arithmetic as oracle/bev_oracle.py, used only as a fast
 * cross-check of the numpy restatement on large inputs; the numpy version is the
 * oracle of record).  generating-dataset/generating_train_bev.py:84-101.
 * points: (N, stride) float32 row-major (x,y,z first).  m[k], t[k]: the float64
 * diagonal and translation of the voxel-space 4x4.  counts: (S1? no) laid out
 * [y][x][z] with dims (S0,S1,S2) square in S0,S1. */
void lvo_bev_counts(const float* points, int64_t n, int32_t stride, const double* m,
                    const double* t, const int32_t* shape, uint32_t* counts) {
  for (int64_t i = 0; i < n; ++i) {
    const float* p = points + i * stride;
    int64_t c[3];
    int ok = 1;
    for (int k = 0; k < 3; ++k) {
      volatile double prod = m[k] * (double)p[k];
      volatile double u = prod + t[k];
      if (!(u > -9.3e18 && u < 9.3e18)) { ok = 0; break; } /* NaN/Inf -> INT64_MIN -> out */
      c[k] = (int64_t)u; /* C truncation == np.intp cast */
      if (c[k] < 0 || c[k] >= shape[k]) { ok = 0; break; }
    }
    if (!ok) continue;
    counts[(c[1] * shape[1] + c[0]) * shape[2] + c[2]] += 1u;
  }
}

/* Block-filtering voxelizer (configs/nuscenes/all.fhd.config:9-12,
 * kwargs at second/second/builder/voxel_builder.py:28-31).  PARITY UNPINNED: the
 * arithmetic lives in spconv 1.x (points_to_voxel_3d_with_filtering, C++), which is
 * neither vendored nor installed; this restates its published algorithm as recalled:
 *   - the same first-come loop as above with the `continue` overflow rule;
 *   - every STORED point (num < max_points at its turn) updates the z-minimum and
 *     z-maximum of its block (y_cell / block_factor, x_cell / block_factor);
 *   - voxel i is kept iff, over the window of block_size x block_size blocks
 *     [b - block_size/2, b + block_size - block_size/2) clipped to the block grid,
 *     height = max - min satisfies height > height_threshold (and, from spconv 1.2 on,
 *     height < height_high_threshold; pass +inf for the 1.1 rule);
 *   - mins / maxs start at +-99999999 (float32).
 * voxel_mask (max_voxels int32) receives the keep flags; the caller compacts
 * (spconv does `coors[voxel_mask]` in Python).  mins / maxs: (H/bf)*(W/bf) floats of
 * scratch.  Returns the UNFILTERED voxel_num. */
int32_t lvo_points_to_voxel_filtered(const float* points, int64_t n, int32_t num_features,
                                     const float* voxel_size, const float* coors_range,
                                     int32_t max_points, int32_t max_voxels, int32_t block_factor,
                                     int32_t block_size, float height_threshold,
                                     float height_high_threshold, int32_t* coor_to_voxelidx,
                                     float* voxels, int32_t* coors, int32_t* num_points_per_voxel,
                                     int32_t* voxel_mask, float* mins, float* maxs) {
  int32_t grid[3];
  lvo_grid_size(voxel_size, coors_range, grid);
  const int64_t H = grid[1], W = grid[0];
  const int32_t BH = grid[1] / block_factor, BW = grid[0] / block_factor;
  for (int64_t i = 0; i < (int64_t)BH * BW; ++i) {
    mins[i] = 99999999.0f;
    maxs[i] = -99999999.0f;
  }
  int32_t voxel_num = 0;
  int32_t coor[3];
  for (int64_t i = 0; i < n; ++i) {
    const float* p = points + i * num_features;
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      volatile float d = p[j] - coors_range[j];
      volatile float q = d / voxel_size[j];
      float c = floorf(q);
      if (!(c >= 0.0f) || c >= (float)grid[j]) {
        failed = 1;
        break;
      }
      coor[2 - j] = (int32_t)c;
    }
    if (failed) continue;
    int64_t cell = ((int64_t)coor[0] * H + coor[1]) * W + coor[2];
    int32_t voxelidx = coor_to_voxelidx[cell];
    if (voxelidx == -1) {
      voxelidx = voxel_num;
      if (voxel_num >= max_voxels) continue;
      voxel_num += 1;
      coor_to_voxelidx[cell] = voxelidx;
      coors[voxelidx * 3 + 0] = coor[0];
      coors[voxelidx * 3 + 1] = coor[1];
      coors[voxelidx * 3 + 2] = coor[2];
    }
    int32_t num = num_points_per_voxel[voxelidx];
    if (num < max_points) {
      memcpy(voxels + ((int64_t)voxelidx * max_points + num) * num_features, p,
             sizeof(float) * (size_t)num_features);
      const int64_t b = (int64_t)(coor[1] / block_factor) * BW + coor[2] / block_factor;
      if (p[2] < mins[b]) mins[b] = p[2];
      if (p[2] > maxs[b]) maxs[b] = p[2];
      num_points_per_voxel[voxelidx] = num + 1;
    }
  }
  for (int32_t v = 0; v < voxel_num; ++v) {
    int64_t cell = ((int64_t)coors[v * 3] * H + coors[v * 3 + 1]) * W + coors[v * 3 + 2];
    coor_to_voxelidx[cell] = -1;
    const int32_t by = coors[v * 3 + 1] / block_factor, bx = coors[v * 3 + 2] / block_factor;
    float mn = mins[(int64_t)by * BW + bx], mx = maxs[(int64_t)by * BW + bx];
    int32_t y0 = by - block_size / 2, y1 = by + block_size - block_size / 2;
    int32_t x0 = bx - block_size / 2, x1 = bx + block_size - block_size / 2;
    if (y0 < 0) y0 = 0;
    if (y1 > BH) y1 = BH;
    if (x0 < 0) x0 = 0;
    if (x1 > BW) x1 = BW;
    for (int32_t j = y0; j < y1; ++j)
      for (int32_t k = x0; k < x1; ++k) {
        const float a = mins[(int64_t)j * BW + k], b2 = maxs[(int64_t)j * BW + k];
        if (a < mn) mn = a;
        if (b2 > mx) mx = b2;
      }
    volatile float height = mx - mn;
    voxel_mask[v] = (height > height_threshold) && (height < height_high_threshold);
  }
  return voxel_num;
}
