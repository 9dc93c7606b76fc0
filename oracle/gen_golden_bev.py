"""Generates tests/golden/ref_bev.npz by EXECUTING THE REFERENCE'S OWN BEV CLOSURES
(generating-dataset/generating_train_bev.py:47-104 + the quantise statement :213) in the build
container, through oracle/ref_loader.run_bev_closures.

Run:  python -m oracle.gen_golden_bev          (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  The dense grids are stored sparsely (flat indices of the non-zero cells +
their counts) so the fixture stays small; the inputs are regenerated from lyft3d_b200.synth on
the GPU box (bundled sweep + seeded clouds), only the adversarial point set is stored.

Cases:
  c1        bundled sweep, 336x336x3 @ (0.4,0.4,1.5), z_offset -2.0           (BASELINE configs[0])
  c11_336   11-sweep aggregated cloud (584,606 points), same grid
  c11_1024  the same cloud, 1024x1024x3 @ (0.2,0.2,1.5)                       (BASELINE configs[3] grid)
  adv       adversarial points: coordinates on cell boundaries (+-1 ulp), the (-1, 0) truncation
            band on every axis, the upper bounds, NaN / +-Inf, huge magnitudes
"""
import hashlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def adversarial_points(seed=99):
    """(4, N) float32: every class of point whose cell the last fp64 rounding, the truncation or the
    bounds test decides."""
    rng = np.random.default_rng(seed)
    vs = np.array(synth.BEV_VOXEL_SIZE)
    cols = []
    # x / y on cell boundaries k*0.4 and one float32 ulp either side, across the whole grid and beyond
    k = np.arange(-172, 173)
    for base in (k * vs[0],):
        b32 = base.astype(np.float32)
        for v in (b32, np.nextafter(b32, np.float32(np.inf)), np.nextafter(b32, np.float32(-np.inf))):
            n = len(v)
            cols.append(np.stack([v, rng.uniform(-60, 60, n).astype(np.float32), rng.uniform(-1.9, 2.0, n).astype(np.float32),
                                  np.zeros(n, np.float32)]))
            cols.append(np.stack([rng.uniform(-60, 60, n).astype(np.float32), v, rng.uniform(-1.9, 2.0, n).astype(np.float32),
                                  np.zeros(n, np.float32)]))
    # z = 1.5 k - 0.25 sits exactly on a cell boundary after the affine (A.1), +- ulp
    kz = np.arange(-4, 6)
    z32 = (1.5 * kz - 0.25).astype(np.float32)
    for v in (z32, np.nextafter(z32, np.float32(np.inf)), np.nextafter(z32, np.float32(-np.inf))):
        n = len(v)
        cols.append(np.stack([rng.uniform(-60, 60, n).astype(np.float32), rng.uniform(-60, 60, n).astype(np.float32), v,
                              np.zeros(n, np.float32)]))
    # the (-1, 0) band in voxel space truncates to cell 0 (not floor): x in (-67.6, -67.2], z in (-1.75, -0.25)
    n = 2000
    cols.append(np.stack([rng.uniform(-67.7, -67.1, n).astype(np.float32), rng.uniform(-67.7, -67.1, n).astype(np.float32),
                          rng.uniform(-1.8, -0.2, n).astype(np.float32), np.zeros(n, np.float32)]))
    # the upper bounds
    cols.append(np.stack([rng.uniform(66.9, 67.3, n).astype(np.float32), rng.uniform(66.9, 67.3, n).astype(np.float32),
                          rng.uniform(2.5, 2.9, n).astype(np.float32), np.zeros(n, np.float32)]))
    # non-finite and huge
    bad = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e20, -1e20, 2.0 ** 31, -2.0 ** 31, 2.0 ** 63, 0.0, -0.0], np.float32)
    for axis in range(3):
        p = np.zeros((4, len(bad)), np.float32)
        p[axis] = bad
        cols.append(p)
    # many points in one cell (count above the clip at 16) and a cell hit exactly 8 times (round-half-even at 127.5)
    cols.append(np.tile(np.array([[1.01], [2.02], [0.3], [0.0]], np.float32), (1, 40)))
    cols.append(np.tile(np.array([[-5.01], [7.02], [0.3], [0.0]], np.float32), (1, 8)))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def cases():
    """name -> (points (4,N) float32, shape, voxel_size, z_offset)."""
    c11 = np.ascontiguousarray(synth.multisweep_cloud(11).T)
    return {
        "c1": (synth.fixture_points_4xn(), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET),
        "c11_336": (c11, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET),
        "c11_1024": (c11, synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE, synth.BEV_Z_OFFSET),
        "adv": (adversarial_points(), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET),
    }


def expand(g, name, shape):
    """Dense (raw f32, norm f32, u8) grids of a stored case."""
    cells = int(np.prod(shape))
    raw = np.zeros(cells, np.float32)
    raw[g[name + "_idx"]] = g[name + "_count"]
    norm = np.zeros(cells, np.float32)
    norm[g[name + "_idx"]] = g[name + "_norm"]
    u8 = np.zeros(cells, np.uint8)
    u8[g[name + "_idx"]] = g[name + "_u8"]
    return raw.reshape(shape), norm.reshape(shape), u8.reshape(shape)


def main():
    assert ref_loader.available(), "needs /root/reference"
    out = {}
    for name, (pts, shape, vs, zo) in cases().items():
        raw, norm, u8 = ref_loader.run_bev_closures(pts, shape, vs, zo)
        assert raw.dtype == np.float32 and norm.dtype == np.float32 and u8.dtype == np.uint8
        idx = np.flatnonzero(raw.reshape(-1)).astype(np.int32)
        out[name + "_idx"] = idx
        out[name + "_count"] = raw.reshape(-1)[idx]
        out[name + "_norm"] = norm.reshape(-1)[idx]
        out[name + "_u8"] = u8.reshape(-1)[idx]
        out[name + "_sha"] = np.array([sha16(raw), sha16(norm), sha16(u8)])
        out[name + "_npoints"] = np.int64(pts.shape[1])
        print(name, pts.shape, "cells", len(idx), "in bounds", int(raw.sum()), "max", raw.max(), out[name + "_sha"])
    out["adv_points"] = adversarial_points()
    np.savez_compressed(os.path.join(GOLD, "ref_bev.npz"), **out)


if __name__ == "__main__":
    main()
