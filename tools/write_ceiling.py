#!/usr/bin/env python
"""Write-only HBM ceiling on this GPU: cudaMemset / torch fill of a canvas-sized buffer
(5.24 GB) timed with CUDA events, next to pillar_canvas_kernel's time on the same buffer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=10):
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
    x = torch.empty((128, 64, 400, 400), dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    nbytes = x.numel() * 4
    for name, fn, moved in (("memset (zero_)", lambda: x.zero_(), nbytes), ("fill_(1.0)", lambda: x.fill_(1.0), nbytes),
                            ("copy_ (read+write)", lambda: y.copy_(x), 2 * nbytes)):
        ms = timed(fn)
        print("%-22s %.4f ms  %.1f GB/s" % (name, ms, moved / ms / 1e6))


if __name__ == "__main__":
    main()
