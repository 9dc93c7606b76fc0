"""B200-native lidar voxelization / BEV rasterization engine.

Drop-in surfaces of jionie/Lyft-3D-Object-Detection's preprocessing hot path
(SURVEY.md 8b), all backed by hand-written sm_100a CUDA kernels behind the
C-ABI of include/lyft_voxel.h (csrc/ -> liblyftvoxel_b200.so):

  bev              create_voxel_pointcloud / normalize_voxel_intensities / ...
  voxel_generator  VoxelGeneratorV2 / VoxelGenerator / points_to_voxel
  pointpillars     PillarFeatureNet* / PointPillarsScatter + the VFE/middle registry
  engine           batched multi-frame engine (frames sharded over GPUs)

There is no CPU fallback: importing is cheap, the first call loads the CUDA
library and raises if it (or a GPU) is missing.
"""
__version__ = "0.1.0"
