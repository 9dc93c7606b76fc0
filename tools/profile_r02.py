#!/usr/bin/env python
"""Driver for the round-2 ncu captures of the kernels bench.py's default step does not launch (or launches in another
shape): batched SECOND clouds through the open-addressing table, a single BEV frame, the training-mode PFN (moments,
forward, backward), the PNG encoder, the fp16 decoration / scatter, and the opt-in fused prologue."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import _native as nat, bev, engine as eng_mod, pointpillars as pp, synth, voxel_generator as vg  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    h = nat.get_handle(0)
    cloud = torch.from_numpy(synth.multisweep_cloud(20)).to(dev)
    n = cloud.shape[0]
    nb = 6
    batch = torch.cat([cloud] * nb).contiguous()
    boffs = np.arange(nb + 1, dtype=np.int64) * n
    for _ in range(2):   # batched C2: vx_cells_kernel<1,1> (hash) + K2-K6 at T = 5
        vg.voxelize_frames(batch, boffs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000, zero_tail=False)
    del batch
    rows1 = torch.from_numpy(synth.fixture_points_nx4()).to(dev)
    for _ in range(2):   # C1, one frame
        bev.rasterize_frames(rows1, np.array([0, rows1.shape[0]], dtype=np.int64), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                             synth.BEV_Z_OFFSET, want=("raw", "u8"))
    # 32 C5 frames: pillarize (five-kernel and fused prologue), training PFN on its rows, PNG of the BEV images
    F = 32
    base = [synth.c5_frame(f) for f in range(8)]
    pts = torch.from_numpy(np.concatenate([base[f % 8] for f in range(F)])).to(dev)
    eng = eng_mod.FrameBatchEngine(0, F, base[0].shape[0])
    eng.voxelize(pts)
    rows = eng.read_total_rows()
    h.set_option("vox_fused_prologue", 1)
    eng.voxelize(pts)
    h.set_option("vox_fused_prologue", 0)
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).to(dev).train()
    for _ in range(2):
        net.zero_grad()
        out = net(eng.voxels[:rows], eng.num_points[:rows], eng.coords[:rows])
        out.sum().backward()
    eng.bev(pts)
    bev.encode_png_frames(eng.bev_u8)
    half = eng.voxels[:rows].half()
    dec = pp.decorate_pillars(half, eng.num_points[:rows], eng.coords[:rows], net.vx, net.vy, net.x_offset, net.y_offset)
    feats = torch.rand((rows, 64), device=dev).half()
    pp.scatter_pillars(feats, eng.coords[:rows], F, 400, 400)
    torch.cuda.synchronize()
    print("ok", rows, tuple(dec.shape))


if __name__ == "__main__":
    main()
