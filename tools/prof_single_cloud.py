#!/usr/bin/env python
"""ncu driver: one 1.06 M-point cloud through the voxelizer at the SECOND (C2) and the pillar (C3) configuration."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import synth, voxel_generator as vg  # noqa: E402

cloud = torch.from_numpy(synth.multisweep_cloud(20)).cuda()
offs = np.array([0, cloud.shape[0]], dtype=np.int64)
for vs, rg, T, V in ((synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000), (synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)):
    for _ in range(3):
        vg.voxelize_frames(cloud, offs, vs, rg, T, V, zero_tail=False)
torch.cuda.synchronize()
print("ok")
