#!/usr/bin/env python
"""A/B of the pillar_canvas kernel variants (lv_set_option "canvas_variant"): 128 C5 frames,
CUDA events around lv_pillar_scatter, canvases compared bit for bit with variant 0."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench import make_pool_frames  # noqa: E402
from lyft3d_b200.engine import FrameBatchEngine  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    F = 128
    pts, n = make_pool_frames(F, dev, 0)
    eng = FrameBatchEngine(0, F, n)
    eng.pillarize(pts)
    rows = eng.read_total_rows()
    eng.features.normal_()
    ref = None
    variants = [int(v) for v in sys.argv[1:]] or [1, 0, 1, 0]   # 1 = row-per-step kernel, 0 = 128-bit gathers
    for v in variants:
        eng.h.set_option("canvas_variant", v)
        for _ in range(3):
            eng.scatter(rows)
        torch.cuda.synchronize()
        ms = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.scatter(rows)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        if ref is None:
            ref = eng.canvas.clone()
            same = True
        else:
            same = bool(torch.equal(ref, eng.canvas))
        print("canvas_variant %d: median %.4f ms  min %.4f ms  identical=%s  (%d pillars)" %
              (v, float(np.median(ms)), min(ms), same, rows))
    eng.h.set_option("canvas_variant", 0)


if __name__ == "__main__":
    main()
