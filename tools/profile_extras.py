#!/usr/bin/env python
"""Small driver for ncu captures of the kernels bench.py's default step does not launch: the target raster
(db_setup / db_raster) and the opt-in cluster-per-frame prologue (vx_frame_kernel)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import _native as nat, bev, synth, voxel_generator as vg  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    scenes = [synth.box_scene(7000 + f, 60) for f in range(128)]
    d_c = torch.from_numpy(np.concatenate([c for c, _ in scenes])).to(dev)
    d_k = torch.from_numpy(np.concatenate([k + 1 for _, k in scenes]).astype(np.int32)).to(dev)
    offs = np.arange(129, dtype=np.int64) * 60
    for _ in range(2):
        bev.rasterize_targets(d_c, d_k, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    frames = np.concatenate([synth.c5_frame(i) for i in range(32)])
    pts = torch.from_numpy(frames).to(dev)
    foffs = np.arange(33, dtype=np.int64) * 53146
    h = nat.get_handle(0)
    h.set_option("vox_frame_kernel", 1)
    for _ in range(2):
        vg.voxelize_frames(pts, foffs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000, zero_tail=False)
    h.set_option("vox_frame_kernel", 0)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
