// placeholder - replaced by the real voxelizer
#include <math.h>
#include "lv_common.cuh"
extern "C" int lv_voxel_grid_size(const lv_voxel_config* cfg, int32_t grid_xyz[3]) {
  LV_REQUIRE(cfg && grid_xyz, "lv_voxel_grid_size: null argument");
  for (int j = 0; j < 3; ++j) {
    volatile float span = cfg->coors_range[3 + j] - cfg->coors_range[j];
    volatile float g = span / cfg->voxel_size[j];
    grid_xyz[j] = (int32_t)rintf(g);
  }
  return LV_OK;
}
extern "C" int lv_voxelize(lv_handle*, const lv_voxel_config*, const float*, int32_t, const int64_t*, float*, int32_t*, int32_t*, int32_t*, lv_stream) { lv_set_error("not built"); return LV_E_UNSUPPORTED; }
extern "C" int lv_voxelize_host(lv_handle*, const lv_voxel_config*, const float*, int32_t, const int64_t*, float*, int32_t*, int32_t*, int32_t*) { lv_set_error("not built"); return LV_E_UNSUPPORTED; }
