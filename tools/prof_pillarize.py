#!/usr/bin/env python
"""Small driver for ncu captures of the voxelizer: `reps` fused voxelize+decorate passes over 128 C5 frames
(python tools/prof_pillarize.py [reps] [frames])."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import engine as eng_mod, synth  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    base = [synth.c5_frame(f) for f in range(16)]
    pts = torch.from_numpy(np.concatenate([base[f % 16] for f in range(F)])).cuda()
    V = int(sys.argv[3]) if len(sys.argv) > 3 else synth.PILLAR_MAX_VOXELS
    eng = eng_mod.FrameBatchEngine(0, F, base[0].shape[0], max_voxels=V)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(reps):
        flush.zero_()
        ev[0].record()
        eng.pillarize(pts)
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    print("pillarize (V=%d): median %.4f ms of %d (L2 flushed), %d pillars" % (V, sorted(ts)[len(ts) // 2], reps, eng.read_total_rows()))


if __name__ == "__main__":
    main()
