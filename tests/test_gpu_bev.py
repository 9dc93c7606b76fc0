"""GPU parity: BEV rasteriser (C-ABI, sm_100a kernels) vs the oracle - bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def bev():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import bev as b
    return b


@pytest.fixture(scope="module")
def bo():
    from oracle import bev_oracle
    return bev_oracle


def test_c1_fixture_golden(bev, bo, fixture_4xn, golden_dir):
    with open(os.path.join(golden_dir, "c1_bev_known.json")) as f:
        known = json.load(f)
    got = bev.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert got.shape == synth.BEV_SHAPE and got.dtype == np.float32
    assert sha16(got) == known["sha256_16_raw_f32"]
    ref = bo.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert np.array_equal(got, ref)
    norm = bev.normalize_voxel_intensities(got)
    assert norm.dtype == np.float32 and np.array_equal(norm, bo.normalize_voxel_intensities(ref))
    assert sha16(norm) == known["sha256_16_norm_f32"]
    u8 = bev.quantize_u8(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert sha16(u8) == known["sha256_16_u8"]
    # a second call must see a clean workspace
    again = bev.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert np.array_equal(again, ref)


def test_cuda_tensor_in_out(bev, bo, fixture_4xn):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(fixture_4xn)).cuda()
    got = bev.create_voxel_pointcloud(t, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert got.is_cuda
    ref = bo.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert np.array_equal(got.cpu().numpy(), ref)
    n = bev.normalize_voxel_intensities(got, 16)
    assert np.array_equal(n.cpu().numpy(), bo.normalize_voxel_intensities(ref))


@pytest.mark.parametrize("shape,vs,zoff", [
    ((336, 336, 3), (0.4, 0.4, 1.5), -2.0),
    ((1024, 1024, 3), (0.2, 0.2, 1.5), -2.0),
    ((1024, 1024, 3), (0.4, 0.4, 1.5), -2.0),
    ((50, 50, 5), (1.0, 1.0, 0.7), 0.3),       # cells % 4 != 0 -> scalar finalize
    ((64, 64, 4), (0.5, 0.5, 1.0), 0.0),
])
def test_multisweep_cloud_shapes(bev, bo, cloud11, shape, vs, zoff):
    pts = np.ascontiguousarray(cloud11.T)
    got = bev.create_voxel_pointcloud(pts, shape, vs, zoff)
    ref = bo.create_voxel_pointcloud(pts, shape, vs, zoff)
    assert np.array_equal(got, ref)


def test_3xn_points_and_adversarial_boundaries(bev, bo):
    # z = 1.5k - 0.25 sits exactly on a cell boundary; x,y on +-edges; NaN/Inf dropped
    z = (1.5 * np.arange(-8, 8) - 0.25).astype(np.float32)
    x = np.array([-67.2, -67.20001, 67.19999, 67.2, 0.0, -0.2, -0.4, 0.39999], np.float32)
    xs, zs = np.meshgrid(x, z)
    pts = np.stack([xs.ravel(), xs.ravel()[::-1].copy(), zs.ravel()]).astype(np.float32)
    bad = np.array([[np.nan, 0, 0], [0, np.inf, 0], [0, 0, -np.inf], [1e30, 0, 0]], np.float32).T
    pts = np.concatenate([pts, bad], axis=1)
    got = bev.create_voxel_pointcloud(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    ref = bo.create_voxel_pointcloud(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert np.array_equal(got, ref) and ref.sum() > 0


def test_empty_and_errors(bev):
    e = bev.create_voxel_pointcloud(np.zeros((4, 0), np.float32), (8, 8, 3), (1, 1, 1), 0)
    assert e.shape == (8, 8, 3) and e.sum() == 0
    with pytest.raises(Exception):
        bev.create_voxel_pointcloud(np.zeros((5, 4), np.float32), (8, 8, 3))
    with pytest.raises(Exception):
        bev.create_voxel_pointcloud(np.zeros((4, 4), np.float32), (8, 8))
    with pytest.raises(Exception):
        bev.create_voxel_pointcloud(np.zeros((4, 4), np.float32), (8, 9, 3))   # non-square, SURVEY.md F8
    with pytest.raises(Exception):
        bev.transform_points(np.zeros((2, 4), np.float32), np.eye(4))


def test_transform_points_and_car_to_voxel_coords(bev, bo, fixture_4xn):
    got = bev.car_to_voxel_coords(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    ref = bo.car_to_voxel_coords(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.array_equal(got, ref)                       # diagonal matrix: bit-exact
    tm = synth.sweep_transform(4)
    g2 = bev.transform_points(fixture_4xn, tm)
    r2 = bo.transform_points(fixture_4xn, tm)
    np.testing.assert_allclose(g2, r2, rtol=1e-14, atol=1e-12)   # full rotation: BLAS order differs in the last fp64 ulp
    tmv = bev.create_transformation_matrix_to_voxel_space(synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, (0, 0, -2.0))
    assert np.array_equal(tmv, bo.create_transformation_matrix_to_voxel_space(
        synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, (0, 0, -2.0))) and tmv.dtype == np.float64


def test_batched_frames_with_sweep_transforms_u8_chw(bev, bo, fixture_nx4):
    """C4-style: 3 frames, frame f made of (f+1) sweeps each with its own sensor->car 4x4,
    raw .bin stride 5, 1024^2 grid, u8 + CHW/map outputs; more frames than frames-in-flight."""
    raw5 = synth.load_fixture_raw()
    seg_pts, seg_frame, seg_tm = [], [], []
    for f in range(3):
        for s in range(f + 1):
            seg_pts.append(raw5[: 20000 + 1000 * s])
            seg_frame.append(f)
            seg_tm.append(synth.sweep_transform(s + 2 * f))
    rows = np.concatenate(seg_pts, axis=0)
    offs = np.zeros(len(seg_pts) + 1, np.int64)
    offs[1:] = np.cumsum([p.shape[0] for p in seg_pts])
    shape, vs = synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE
    maps = np.stack([synth.map_raster(seed=4000 + f) for f in range(3)])
    res = bev.rasterize_frames(rows, offs, shape, vs, synth.BEV_Z_OFFSET, seg_frame=seg_frame,
                               seg_tm=np.stack(seg_tm), want=("raw", "norm", "u8", "chw"), map_u8=maps)
    mism = 0
    for f in range(3):
        cloud = []
        for k, (p, sf, tm) in enumerate(zip(seg_pts, seg_frame, seg_tm)):
            if sf == f:
                cloud.append(bo.sensor_to_car(np.ascontiguousarray(p[:, :4].T), tm))
        cloud = np.concatenate(cloud, axis=1)
        ref = bo.create_voxel_pointcloud(cloud, shape, vs, synth.BEV_Z_OFFSET)
        mism += int((res["raw"][f] != ref).sum())
        nref = bo.normalize_voxel_intensities(ref)
        u8 = bo.quantize_u8(nref)
        assert np.array_equal(res["norm"][f], nref)
        assert np.array_equal(res["u8"][f], u8)
        assert np.array_equal(res["chw"][f], bo.bev_concat_map_chw(u8, maps[f]))
    # fp64 FMA chain rounded once to fp32 == BLAS result after the fp32 store
    assert mism == 0


def test_many_small_frames_exceeding_frames_in_flight(bev, bo):
    frames = [synth.c5_frame(f)[:5000] for f in range(40)]
    rows = np.concatenate(frames)
    offs = np.arange(41, dtype=np.int64) * 5000
    res = bev.rasterize_frames(rows, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET,
                               want=("raw", "u8"))
    for f in (0, 1, 22, 23, 24, 39):
        ref = bo.create_voxel_pointcloud(np.ascontiguousarray(frames[f].T), synth.BEV_SHAPE,
                                         synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
        assert np.array_equal(res["raw"][f], ref)
        assert np.array_equal(res["u8"][f], bo.quantize_u8(bo.normalize_voxel_intensities(ref)))
    # size-independent property: every in-bounds point is counted exactly once
    total = sum(int(bo.create_voxel_pointcloud(np.ascontiguousarray(fr.T), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                               synth.BEV_Z_OFFSET).sum()) for fr in frames)
    assert int(res["raw"].sum(dtype=np.float64)) == total


@pytest.mark.parametrize("stride", [4, 5])
def test_tma_staged_tiles_ragged_frames_and_unaligned_views(bev, bo, stride):
    """Tiles of 2048 rows are staged by TMA bulk copies when the rows are 16-byte aligned;
    ragged frames make tiles straddle frame boundaries, a row-offset view breaks the
    alignment (stride 5) and must take the plain-load path; the TMA path (option `bev_tma`,
    off by default because it measured slower) must not change a single count."""
    import torch
    from lyft3d_b200 import _native as nat
    raw5 = synth.load_fixture_raw()
    sizes = [2048, 1, 4095, 0, 6151, 777, 20000, 33]
    rng = np.random.default_rng(77)
    frames = []
    for f, n in enumerate(sizes):
        sel = rng.integers(0, raw5.shape[0], size=n)
        fr = raw5[sel].copy()
        fr[:, 0] += 0.37 * f
        frames.append(np.ascontiguousarray(fr[:, :stride]))
    rows = np.concatenate(frames, axis=0)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    refs = [bo.create_voxel_pointcloud(np.ascontiguousarray(fr[:, :4].T), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                       synth.BEV_Z_OFFSET) for fr in frames]
    h = nat.get_handle(0)
    pad = torch.zeros((rows.shape[0] + 3, stride), dtype=torch.float32, device="cuda")
    for shift in (0, 1, 3):           # shift != 0 with stride 5 -> rows not 16-byte aligned
        view = pad[shift:shift + rows.shape[0]]
        view.copy_(torch.from_numpy(rows))
        for on in (1, 0):
            h.set_option("bev_tma", on)
            try:
                res = bev.rasterize_frames(view, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET,
                                           want=("raw", "u8"))
            finally:
                h.set_option("bev_tma", 0)
            got = res["raw"].cpu().numpy()
            for f in range(len(sizes)):
                assert np.array_equal(got[f], refs[f]), (stride, shift, on, f)
            assert np.array_equal(res["u8"][2].cpu().numpy(), bo.quantize_u8(bo.normalize_voxel_intensities(refs[2])))


def test_u16_counts_and_fused_zero_fill_change_nothing(bev, bo):
    """Options of the histogram: 16-bit packed counts (lv_set_option("bev_u16", 1): used whenever every frame of a call
    has fewer than 65,536 points; measured slower, so not the default) and the zero fill inside the histogram kernel
    ("bev_fused_zero").  Every combination gives the same bytes - including a frame with 20,000 copies of one point
    (a count far above the uint8 saturation), the 1024^2 CHW path, and a call in which one frame has 70,000 points
    (falls back to 32-bit counts by itself)."""
    import torch
    from lyft3d_b200 import _native as nat
    h = nat.get_handle(0)
    hot = np.tile(np.array([[1.3, -2.2, -0.5, 0.5]], dtype=np.float32), (20000, 1))
    small = [synth.c5_frame(f)[:40000] for f in range(3)] + [hot, synth.c5_frame(5)[:1]]
    big = small + [np.concatenate([synth.c5_frame(6), synth.c5_frame(7)])[:70000]]
    for frames in (small, big):
        rows = torch.from_numpy(np.concatenate(frames)).cuda()
        offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
        outs = []
        for u16 in (1, 0):
            for fz in (0, 1):
                h.set_option("bev_u16", u16)
                h.set_option("bev_fused_zero", fz)
                try:
                    for _ in range(2):      # twice: the counts and the bitmap must be left all-zero
                        res = bev.rasterize_frames(rows, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET,
                                                   want=("raw", "norm", "u8"))
                finally:
                    h.set_option("bev_u16", 0)
                    h.set_option("bev_fused_zero", 0)
                outs.append({k: v.clone() for k, v in res.items()})
        for o in outs[1:]:
            for k in ("raw", "norm", "u8"):
                assert torch.equal(outs[0][k], o[k]), k
        for f in (0, 3, 4, len(frames) - 1):
            ref = bo.create_voxel_pointcloud(np.ascontiguousarray(frames[f].T), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                             synth.BEV_Z_OFFSET)
            assert np.array_equal(outs[0]["raw"][f].cpu().numpy(), ref), f
        assert float(outs[0]["raw"][3].max()) == 20000.0
    # the CHW / map path (hwc3 finalize) in both count widths
    maps = torch.from_numpy(synth.map_raster(seed=4000)[None]).cuda()
    rows = torch.from_numpy(synth.c5_frame(1)).cuda()
    offs = np.array([0, rows.shape[0]], dtype=np.int64)
    chw = []
    for u16 in (1, 0):
        h.set_option("bev_u16", u16)
        try:
            res = bev.rasterize_frames(rows, offs, synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE, synth.BEV_Z_OFFSET,
                                       want=("u8", "chw"), map_u8=maps)
        finally:
            h.set_option("bev_u16", 0)
        chw.append((res["u8"].clone(), res["chw"].clone()))
    assert torch.equal(chw[0][0], chw[1][0]) and torch.equal(chw[0][1], chw[1][1])
