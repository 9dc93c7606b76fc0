"""Pins oracle/bev_oracle.py: known answers on the bundled sweep,
the order-pinned fp64 affine == BLAS dot, trunc-vs-floor, error behaviour."""
import hashlib
import json
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import bev_oracle as bo
from oracle import voxel_oracle as vo


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def known(golden_dir):
    with open(os.path.join(golden_dir, "c1_bev_known.json")) as f:
        return json.load(f)


def test_c1_known_answers(fixture_4xn, known):
    bev = bo.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert bev.dtype == np.float32 and bev.shape == synth.BEV_SHAPE
    assert fixture_4xn.shape[1] == known["n_points"] == 53146
    assert int(bev.sum()) == known["in_bounds"] == 47213
    assert int((bev > 0).sum()) == known["nonzero_cells"] == 6595
    assert int(bev.max()) == known["max_count"] == 594
    assert int((bev >= 16).sum()) == known["saturated_cells"] == 440
    assert [int(v) for v in bev.sum(axis=(0, 1))] == known["channel_sums"] == [42977, 2924, 1312]
    # survey-time hashes, SURVEY.md 8(c)
    assert sha16(bev) == known["sha256_16_raw_f32"] == "d8aa630368259baa"
    u8 = bo.quantize_u8(bo.normalize_voxel_intensities(bev))
    assert sha16(u8) == known["sha256_16_u8"] == "06fc320850db7693"


def test_matrix_is_float64_and_values():
    tm = bo.create_transformation_matrix_to_voxel_space(synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, (0, 0, synth.BEV_Z_OFFSET))
    assert tm.dtype == np.float64
    assert tm[0, 0] == 2.5 and tm[1, 1] == 2.5 and tm[2, 2] == 1 / 1.5
    assert tm[0, 3] == 168.0 and tm[1, 3] == 168.0
    assert tm[2, 3] == 1.5 + (-2.0 / 1.5)


def test_elementwise_equals_blas_dot(fixture_4xn):
    rng = np.random.default_rng(3)
    rnd = (rng.normal(size=(4, 200000)) * 60).astype(np.float32)
    adv = np.zeros((4, 64), np.float32)
    adv[2, :] = (1.5 * np.arange(64) - 0.25 - 30).astype(np.float32)  # z on cell boundaries (A.1)
    for pts in (fixture_4xn, rnd, adv):
        a = bo.car_to_voxel_coords(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
        b = bo.car_to_voxel_coords_elementwise(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
        assert np.array_equal(a, b)


def test_truncation_not_floor():
    # t_z = 3/2 - 2/1.5 = 0.1667.  z=-1.0 -> u_z=-0.5 -> trunc 0 (kept; floor would drop it)
    pts = np.array([[0.1, 0.1, -1.0, 1.0], [-0.1, 0.1, 0.0, 1.0]], np.float32).T
    bev = bo.create_voxel_pointcloud(pts, (8, 8, 3), (1.0, 1.0, 1.5), -2.0)
    assert bev.sum() == 2
    assert bev[4, 4, 0] == 1      # x=0.1 -> 4.1 -> 4 ; y=0.1 -> 4 ; z -> 0
    assert bev[4, 3, 0] == 1      # x=-0.1 -> 3.9 -> 3 ; bev[y, x, z]


def test_truncation_negative_fraction_lands_in_cell_zero():
    # u in (-1, 0) truncates to 0: in bounds (floor would drop it)
    pts = np.array([[-3.99, 0.0, 0.0, 0.0]], np.float32).T   # x*1 + 4 = 0.01 -> 0
    pts2 = np.array([[-4.5, 0.0, 0.0, 0.0]], np.float32).T   # -0.5 -> trunc 0 -> kept
    pts3 = np.array([[-5.0, 0.0, 0.0, 0.0]], np.float32).T   # -1.0 -> -1 -> dropped
    for p, n in ((pts, 1), (pts2, 1), (pts3, 0)):
        bev = bo.create_voxel_pointcloud(p, (8, 8, 3), (1.0, 1.0, 1.0), 0)
        assert bev.sum() == n


def test_bad_shapes_raise():
    with pytest.raises(Exception):
        bo.car_to_voxel_coords(np.zeros((5, 10), np.float32), (8, 8, 3), (1, 1, 1))
    with pytest.raises(Exception):
        bo.car_to_voxel_coords(np.zeros((4, 10), np.float32), (8, 8), (1, 1, 1))
    with pytest.raises(Exception):
        bo.transform_points(np.zeros((2, 10), np.float32), np.eye(4))


def test_c_crosscheck_large(cloud11):
    for shape, vs in ((synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE), (synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE)):
        ref = bo.create_voxel_pointcloud(cloud11.T, shape, vs, synth.BEV_Z_OFFSET)
        c = vo.bev_counts_c(cloud11, shape, vs, synth.BEV_Z_OFFSET)
        assert np.array_equal(c.astype(np.float32), ref)


def test_normalize_and_quantize():
    bev = np.arange(0, 40, dtype=np.float32).reshape(2, 4, 5)
    n = bo.normalize_voxel_intensities(bev)
    assert n.dtype == np.float32 and n.max() == 1.0 and n[0, 0, 1] == np.float32(1 / 16)
    u8 = bo.quantize_u8(n)
    assert u8[0, 1, 3] == 128   # count 8 -> 127.5 -> round-half-even 128
    assert u8.max() == 255


def test_sensor_to_car_rounds_once_to_f32(fixture_4xn):
    tm = synth.sweep_transform(3)
    out = bo.sensor_to_car(fixture_4xn, tm)
    assert out.dtype == np.float32
    p = fixture_4xn.astype(np.float64)
    x = tm[0, 0] * p[0] + tm[0, 1] * p[1] + tm[0, 2] * p[2] + tm[0, 3]
    # any fp64 accumulation order rounds to the same float32 ("Float semantics")
    assert np.array_equal(out[0], x.astype(np.float32))
    assert np.array_equal(out[3], fixture_4xn[3])
