#!/usr/bin/env python
"""Do latency-bound kernels hide behind the canvas stream?  Times, over 20 repetitions on a B200:
canvas alone, BEV alone, voxelize+decorate alone, and pairs of them on two streams (the second one
high-priority).  Measured (profiles/README.md): the pairs take the SUM of their parts - kernels of
different streams do not overlap usefully here, the canvas kernel fills every SM."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import _native as nat, synth  # noqa: E402
from lyft3d_b200.engine import FrameBatchEngine  # noqa: E402


def main():
    F = 128
    frames = np.concatenate([synth.c5_frame(f) for f in range(8)])
    n = frames.shape[0] // 8
    pts = torch.from_numpy(np.tile(frames, (F // 8, 1))).cuda()
    eng = FrameBatchEngine(0, F, n)
    eng.features.normal_()
    eng.pillarize(pts)
    rows = eng.read_total_rows()
    s1 = torch.cuda.Stream(priority=0)
    s2 = torch.cuda.Stream(priority=-1)
    R = 20

    def canvas(s):
        with torch.cuda.stream(s):
            eng.scatter(rows)

    def bev(s):
        with torch.cuda.stream(s):
            eng.bev(pts)

    def pillarize(s):
        with torch.cuda.stream(s):
            eng.pillarize(pts)

    def timed(name, fa, fb=None):
        for _ in range(3):
            fa(s1)
            if fb:
                fb(s2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(R):
            fa(s1)
            if fb:
                fb(s2)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1)
        e2.record(s2)
        cur = torch.cuda.current_stream()
        cur.wait_event(e1)
        cur.wait_event(e2)
        b.record()
        torch.cuda.synchronize()
        print("%-44s %.4f ms per repetition" % (name, a.elapsed_time(b) / R), flush=True)

    timed("canvas alone", canvas)
    timed("BEV alone", bev)
    timed("voxelize+decorate alone", pillarize)
    timed("canvas || BEV (high priority)", canvas, bev)
    timed("canvas || voxelize+decorate (high priority)", canvas, pillarize)
    timed("voxelize+decorate || BEV (high priority)", pillarize, bev)


if __name__ == "__main__":
    main()
