#!/usr/bin/env python
"""A/B timing of the TMA-staged point tiles against plain loads (run on a B200).

    python tools/ab_tma.py            # BEV stride 4 / 5 and voxelizer C 4 / 5, TMA on / off

Prints one line per case: CUDA-event ms over 128 C5-shaped frames, median of 20 calls.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import _native as nat, bev, synth, voxel_generator as vg  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    F = 128
    raw5 = synth.load_fixture_raw()
    n = raw5.shape[0]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    base = torch.from_numpy(raw5).to(dev)
    rows5 = base[None].repeat(F, 1, 1)
    ang = torch.rand(F, generator=g, device=dev) * 6.2831853
    c, s = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    x, y = rows5[:, :, 0].clone(), rows5[:, :, 1].clone()
    rows5[:, :, 0] = c * x - s * y
    rows5[:, :, 1] = s * x + c * y
    rows5 = rows5.reshape(F * n, 5).contiguous()
    rows4 = rows5[:, :4].contiguous()
    offs = np.arange(F + 1, dtype=np.int64) * n
    h = nat.get_handle(0)
    out = {"norm": torch.empty((F,) + tuple(synth.BEV_SHAPE), dtype=torch.float32, device=dev),
           "u8": torch.empty((F,) + tuple(synth.BEV_SHAPE), dtype=torch.uint8, device=dev)}
    for stride, rows in ((4, rows4), (5, rows5)):
        for off in (0, 1):
            h.set_option("bev_tma", 1 - off)
            ms = timed(lambda: bev.rasterize_frames(rows, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                                    synth.BEV_Z_OFFSET, want=("norm", "u8"), out=out))
            print("bev      stride %d  tma %-3s  %.4f ms  (%d frames x %d points)" %
                  (stride, "off" if off else "on", ms, F, n))
    for C, rows in ((4, rows4), (5, rows5)):
        for off in (0, 1):
            h.set_option("disable_tma", off)
            ms = timed(lambda: vg.voxelize_frames(rows, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 8, 12000,
                                                  zero_tail=False))
            print("voxelize C %d       tma %-3s  %.4f ms" % (C, "off" if off else "on", ms))
    h.set_option("disable_tma", 0)
    h.set_option("bev_tma", 0)


if __name__ == "__main__":
    main()
