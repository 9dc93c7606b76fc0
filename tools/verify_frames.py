#!/usr/bin/env python
"""Scans frames of the C5 set: both paths on the GPU against the oracle, frame by frame; prints every frame that is not
bit-exact where it must be, and the largest f_cluster deviation (absolute, and as a fraction of the summation-order bound)
(python tools/verify_frames.py [first [count]])."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lyft3d_b200 import engine as eng_mod, synth  # noqa: E402
from oracle import bev_oracle, pillar_oracle, voxel_oracle  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = bench.workload_cfg()
F = 16
worst_all, abs_all, bad = 0.0, 0.0, 0
for b0 in range(first, first + count, F):
    ids = list(range(b0, min(b0 + F, first + count)))
    frames = [synth.c5_frame(f) for f in ids]
    n = frames[0].shape[0]
    eng = eng_mod.FrameBatchEngine(0, len(ids), n)
    eng.step(torch.from_numpy(np.concatenate(frames)).cuda())
    torch.cuda.synchronize()
    offs = eng.voxel_offsets.cpu().numpy()
    for i, f in enumerate(ids):
        fr = frames[i]
        raw = bev_oracle.create_voxel_pointcloud(np.ascontiguousarray(fr.T), cfg["bev_shape"], cfg["bev_voxel_size"], cfg["bev_z_offset"])
        u8 = bev_oracle.quantize_u8(bev_oracle.normalize_voxel_intensities(raw))
        v, c, nn = voxel_oracle.points_to_voxel(fr, cfg["voxel_size"], cfg["pc_range"], cfg["max_points"], cfg["max_voxels"])
        dec = pillar_oracle.decorate(v, nn, pillar_oracle.merge_batch_coords([c]), cfg["voxel_size"], cfg["pc_range"])
        r0, r1 = int(offs[i]), int(offs[i + 1])
        exact = (np.array_equal(eng.bev_u8[i].cpu().numpy(), u8) and r1 - r0 == v.shape[0] and
                 np.array_equal(eng.coords[r0:r1, 1:].cpu().numpy(), c) and np.array_equal(eng.num_points[r0:r1].cpu().numpy(), nn))
        ok, worst, max_abs = pillar_oracle.decorate_mismatch(eng.decorated[r0:r1].cpu().numpy(), dec, v, nn)
        worst_all, abs_all = max(worst_all, worst), max(abs_all, max_abs)
        if not (exact and ok):
            bad += 1
            print(json.dumps({"frame": f, "bit_exact_outputs": bool(exact), "decoration_ok": bool(ok), "worst": worst, "max_abs": max_abs}))
print("frames %d..%d: %d not ok; largest f_cluster deviation %.3g = %.3f of the summation-order bound" %
      (first, first + count - 1, bad, abs_all, worst_all))
