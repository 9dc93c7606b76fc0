"""Generates tests/golden/ref_simplevis_3d.npz by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden_simplevis3d        (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  The reference's in-tree sibling of the voxelizer, second/second/utils/simplevis.py:9-108
(``points_to_bev``), at THREE-DIMENSIONAL grids (several height slices) and with the voxel cap hit at different places:
for every case the per-(y, x) point-count map (bev_map[-1]) and the per-cell height maps (bev_map[:-1]: the highest
point of every occupied (z, y, x) cell above the slice floor, in units of the slice height).  The `break` rule of
:46-50 decides which cells exist at all, so every case with a small cap also pins the first-come ORDER of the cells.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")
# (name, first point, points, voxel size, range, max_voxels)
CASES = [("second_coarse", 0, 8000, (0.2, 0.2, 1.0), (0.0, -32.0, -3.0, 52.8, 32.0, 1.0), 1500),
         ("second_coarse_all", 0, 20000, (0.2, 0.2, 1.0), (0.0, -32.0, -3.0, 52.8, 32.0, 1.0), 40000),
         ("slices8", 5000, 30000, (0.5, 0.5, 0.5), (-40.0, -40.0, -3.0, 40.0, 40.0, 1.0), 700),
         ("slices8_cap1", 5000, 30000, (0.5, 0.5, 0.5), (-40.0, -40.0, -3.0, 40.0, 40.0, 1.0), 1),
         ("fine", 20000, 25000, (0.1, 0.1, 2.0), (-20.0, -20.0, -3.0, 20.0, 20.0, 1.0), 5000)]


def case_points(first, n):
    return np.ascontiguousarray(synth.fixture_points_nx4()[first:first + n])


def main():
    assert ref_loader.available(), "needs /root/reference"
    sv = ref_loader.load_simplevis()
    out = {}
    for name, first, n, vs, rg, mv in CASES:
        bm = sv.points_to_bev(case_points(first, n), vs, rg, max_voxels=mv)
        assert bm[-1].max() < 65535
        out[name + ".count"] = bm[-1].astype(np.uint16)
        out[name + ".height"] = bm[:-1].astype(np.float32)
        print(name, bm.shape, "points", int(bm[-1].sum()), "columns", int((bm[-1] > 0).sum()), "cells", int((bm[:-1] > 0).sum()))
    path = os.path.join(GOLD, "ref_simplevis_3d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
