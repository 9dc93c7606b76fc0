"""Pins oracle/voxel_oracle.py against the reference's in-tree sibling kernel
(second/second/utils/simplevis.py:9-61, outputs stored by oracle/gen_golden.py),
the pure-Python loop, the frozen C2/C3 hashes and structural properties."""
import hashlib
import json
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import ref_loader
from oracle import voxel_oracle as vo


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def density(coors, num, ny=400, nx=400):
    dm = np.zeros((ny, nx), np.int64)
    dm[coors[:, 1], coors[:, 2]] = num
    return dm


def test_against_reference_points_to_bev(fixture_nx4, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_simplevis_pillar.npz"))
    v, c, n = vo.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 100000, 40000)
    assert np.array_equal(density(c, n), g["density_full"].astype(np.int64))
    assert v.shape[0] == 8569 and n.sum() == 47732
    # `break` rule: simplevis.py:46-50 with max_voxels=3000
    v, c, n = vo.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 100000, 3000,
                                 overflow="break")
    assert np.array_equal(density(c, n), g["density_break3000"].astype(np.int64))
    assert v.shape[0] == 3000 and n.sum() == 7764
    # `continue` keeps filling existing voxels past the cap
    v2, c2, n2 = vo.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 100000, 3000,
                                    overflow="continue")
    assert np.array_equal(c2, c) and n2.sum() > n.sum() and np.all(n2 >= n)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_against_live_reference_second_config(fixture_nx4):
    """Live run of the reference kernel at a coarser SECOND-like 3-D grid (z slices > 1)."""
    sv = ref_loader.load_simplevis()
    sub = fixture_nx4[:8000]
    vs, rg = (0.2, 0.2, 1.0), (0.0, -32.0, -3.0, 52.8, 32.0, 1.0)
    bm = sv.points_to_bev(sub, vs, rg, max_voxels=1500)
    v, c, n = vo.points_to_voxel(sub, vs, rg, 100000, 1500, overflow="break")
    gs = vo.grid_size(vs, rg)
    dm = np.zeros((gs[1], gs[0]), np.int64)
    np.add.at(dm, (c[:, 1], c[:, 2]), n)
    assert np.array_equal(dm, bm[-1].astype(np.int64))


@pytest.mark.parametrize("mode", ["continue", "break"])
@pytest.mark.parametrize("cfg", ["second", "pillar"])
def test_c_equals_python_loop(fixture_nx4, mode, cfg):
    sub = fixture_nx4[10000:16000]
    if cfg == "second":
        args = (synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 700)
    else:
        args = (synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 7, 600)
    a = vo.points_to_voxel_loop(sub, *args, overflow=mode)
    b = vo.points_to_voxel(sub, *args, overflow=mode)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert a[0].shape[0] == args[3]  # cap reached -> overflow rule exercised


def test_grid_size_rounding():
    assert vo.grid_size(synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE).tolist() == [1056, 1280, 40]
    assert vo.grid_size(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).tolist() == [400, 400, 1]
    # notebook FHD variant, SURVEY.md 7 ("[41 2000 2000]")
    assert vo.grid_size((0.05, 0.05, 0.1), (-50, -50, -3, 50, 50, 1.1)).tolist() == [2000, 2000, 41]


def test_frozen_hashes(golden_dir, fixture_nx4, cloud11):
    with open(os.path.join(golden_dir, "voxel_oracle_hashes.json")) as f:
        frozen = json.load(f)
    cases = {"c3_single": (fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
             "c2_11": (cloud11, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
             "c3_11": (cloud11, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)}
    for name, (pts, vs, rg, T, mv) in cases.items():
        orc = vo.VoxelOracle(vs, rg, T, mv)
        for mode in ("continue", "break"):
            v, c, n = orc.generate(pts, overflow=mode)
            fz = frozen["%s_%s" % (name, mode)]
            assert v.shape[0] == fz["voxel_num"] and orc.last_kept_points == fz["kept_points"]
            assert (sha16(v), sha16(c), sha16(n)) == (fz["sha_voxels"], fz["sha_coords"], fz["sha_num"])
        assert np.all(orc._map == -1)  # touched cells reset


def test_properties(cloud11):
    T, mv = 5, 60000
    v, c, n = vo.points_to_voxel(cloud11, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, T, mv)
    assert v.shape == (mv, T, 4) and c.shape == (mv, 3) and n.shape == (mv,)
    assert c.dtype == np.int32 and n.dtype == np.int32 and v.dtype == np.float32
    gs = vo.grid_size(synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE)
    assert np.all(c >= 0) and np.all(c < gs[::-1][None, :])
    assert len(np.unique(c, axis=0)) == mv                      # coords unique
    assert n.min() >= 1 and n.max() <= T
    mask = np.arange(T)[None, :] < n[:, None]
    assert np.all(v[~mask] == 0)                                # zero padding
    # every stored point falls in its voxel (zyx order, floor rule)
    lo = np.asarray(synth.SECOND_RANGE[:3], np.float32)
    vs = np.asarray(synth.SECOND_VOXEL_SIZE, np.float32)
    cc = np.floor((v[..., :3] - lo) / vs).astype(np.int32)[..., ::-1]
    assert np.all((cc == c[:, None, :])[mask])


def test_empty_and_all_out_of_range():
    e = np.zeros((0, 4), np.float32)
    v, c, n = vo.points_to_voxel(e, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, 100)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    far = np.full((10, 4), 1e6, np.float32)
    v, c, n = vo.points_to_voxel(far, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, 100)
    assert v.shape[0] == 0
    nan = np.full((4, 4), np.nan, np.float32)
    v, c, n = vo.points_to_voxel(nan, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, 100)
    assert v.shape[0] == 0


def test_upper_boundary_is_exclusive():
    pts = np.array([[50.0, 0, 0, 0], [49.999, 0, 0, 0], [-50.0, 0, 0, 0], [-50.001, 0, 0, 0],
                    [0, 0, 10.0, 0], [0, 0, -10.0, 0]], np.float32)
    v, c, n = vo.points_to_voxel(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, 100)
    assert c.tolist() == [[0, 200, 399], [0, 200, 0], [0, 200, 200]]


def _maps_from_voxels(voxels, coords, num, vs, rg, grid):
    """What simplevis.points_to_bev draws (:52-57), rebuilt from a voxelizer's output with an uncapped point list:
    count per (y, x) column and, per (z, y, x) cell, the highest point above the slice floor in slice heights."""
    gx, gy, gz = (int(v) for v in grid)
    count = np.zeros((gy, gx), np.int64)
    np.add.at(count, (coords[:, 1], coords[:, 2]), num)
    lowers = np.linspace(np.float32(rg[2]), np.float32(rg[5]), gz, endpoint=False).astype(np.float32)
    T = voxels.shape[1]
    live = np.arange(T)[None, :] < num[:, None]
    hn = ((voxels[:, :, 2] - lowers[coords[:, 0]][:, None]) / np.float32(vs[2])).astype(np.float32)
    hmax = np.where(live, hn, -np.inf).max(axis=1) if voxels.shape[0] else np.zeros((0,), np.float32)
    height = np.zeros((gz, gy, gx), np.float32)
    height[coords[:, 0], coords[:, 1], coords[:, 2]] = np.maximum(hmax, 0).astype(np.float32)
    return count, height


def test_against_reference_points_to_bev_3d_grids(golden_dir):
    """tests/golden/ref_simplevis_3d.npz (oracle/gen_golden_simplevis3d.py ran the reference's points_to_bev): grids
    with several height slices, caps hit early (the `break` rule decides which cells exist, i.e. first-come order)."""
    from oracle import gen_golden_simplevis3d as gg
    g = np.load(os.path.join(golden_dir, "ref_simplevis_3d.npz"))
    for name, first, n, vs, rg, mv in gg.CASES:
        pts = gg.case_points(first, n)
        v, c, k = vo.points_to_voxel(pts, vs, rg, 4096, mv, overflow="break")
        count, height = _maps_from_voxels(v, c, k, vs, rg, vo.grid_size(vs, rg))
        assert np.array_equal(count, g[name + ".count"].astype(np.int64)), name
        assert np.array_equal(height, g[name + ".height"]), name
