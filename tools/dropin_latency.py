#!/usr/bin/env python
"""Wall-clock latency of the drop-in calls with HOST numpy arrays in and out (what a DataLoader worker of the
reference does: second/second/data/preprocess.py:299-317), next to the CPU oracle on the same inputs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyft3d_b200 import synth, voxel_generator as vg  # noqa: E402
from oracle import voxel_oracle  # noqa: E402


def wall(fn, reps=10):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    cloud = synth.multisweep_cloud(20)
    sweep = synth.fixture_points_nx4()
    for name, pts, vs, rg, T, V in (
            ("C2 SECOND 0.05 m, 1.06 M points", cloud, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
            ("C3 pillars 0.25 m, 1.06 M points", cloud, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
            ("C3 pillars 0.25 m, one sweep (53,146 points)", sweep, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)):
        gen = vg.VoxelGeneratorV2(vs, rg, T, max_voxels=V)
        ms = wall(lambda: gen.generate(pts, V))
        ora = voxel_oracle.VoxelOracle(vs, rg, T, V)
        cpu = wall(lambda: ora.generate(pts), reps=3)
        print("%-48s generate(): %.3f ms   CPU oracle (C, 1 core): %.2f ms" % (name, ms, cpu))
    # sixteen sweeps per call through generate_batch: launches and synchronisation paid once per batch
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, max_voxels=30000)
    sweeps = [synth.c5_frame(f) for f in range(16)]
    ms = wall(lambda: gen.generate_batch(sweeps, 30000))
    print("%-48s generate_batch(16 sweeps): %.3f ms = %.3f ms per sweep" % ("C3 pillars 0.25 m, 53,146 points each", ms, ms / 16))


if __name__ == "__main__":
    main()
