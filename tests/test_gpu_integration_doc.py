"""GPU: the raw ctypes bindings printed in INTEGRATION.md (sections 1, 2, 6, 7) work as written - a
maintainer who pastes them into the reference gets the oracle's results.  No shim from the package is
used here, only the shared library."""
import ctypes
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from lyft3d_b200._native import LIB_PATH

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raw():
    lib = ctypes.CDLL(LIB_PATH)
    lib.lv_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    assert lib.lv_create(0, ctypes.byref(h)) == 0
    yield lib, h
    lib.lv_destroy(h)


class lv_voxel_config(ctypes.Structure):
    _fields_ = [("voxel_size", ctypes.c_float * 3), ("coors_range", ctypes.c_float * 6),
                ("max_points", ctypes.c_int32), ("max_voxels", ctypes.c_int32),
                ("num_features", ctypes.c_int32), ("overflow_mode", ctypes.c_int32),
                ("zero_tail", ctypes.c_int32)]


class lv_block_filter(ctypes.Structure):
    _fields_ = [("block_factor", ctypes.c_int32), ("block_size", ctypes.c_int32),
                ("height_threshold", ctypes.c_float), ("height_high_threshold", ctypes.c_float)]


def test_section_1_bev_binding(raw, fixture_4xn):
    from oracle import bev_oracle
    lib, h = raw

    def create_voxel_pointcloud(points, shape, voxel_size=(0.5, 0.5, 1), z_offset=0):
        rows = np.ascontiguousarray(points.T, dtype=np.float32)          # (N, 3|4)
        offs = np.array([0, rows.shape[0]], dtype=np.int64)
        bev = np.empty(shape, dtype=np.float32)
        rc = lib.lv_bev_rasterize_host(
            h, rows.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(rows.shape[1]), 1,
            offs.ctypes.data_as(ctypes.c_void_p), None, None, 1,
            (ctypes.c_int32 * 3)(*shape), (ctypes.c_double * 3)(*voxel_size), ctypes.c_double(z_offset),
            ctypes.c_float(16.0), bev.ctypes.data_as(ctypes.c_void_p), None, None, None, None)
        if rc:
            raise Exception(lib.lv_last_error(h).decode())
        return bev

    got = create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    ref = bev_oracle.create_voxel_pointcloud(fixture_4xn, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert np.array_equal(got, ref)


def _cfg(vs, rg, T, V):
    cfg = lv_voxel_config()
    cfg.voxel_size[:] = list(vs)
    cfg.coors_range[:] = list(rg)
    cfg.max_points, cfg.max_voxels, cfg.num_features, cfg.overflow_mode, cfg.zero_tail = T, V, 4, 0, 0
    return cfg


def test_section_2_voxel_generator_binding(raw, fixture_nx4):
    from oracle import voxel_oracle
    lib, h = raw

    def generate(points, cfg):                       # points (N, C) float32
        V, T, C = cfg.max_voxels, cfg.max_points, cfg.num_features
        voxels = np.zeros((V, T, C), np.float32)
        coords = np.zeros((V, 3), np.int32)
        num = np.zeros((V,), np.int32)
        vnum = np.zeros((1,), np.int32)
        offs = np.array([0, points.shape[0]], np.int64)
        rc = lib.lv_voxelize_host(h, ctypes.byref(cfg), points.ctypes.data_as(ctypes.c_void_p), 1,
                                  offs.ctypes.data_as(ctypes.c_void_p), voxels.ctypes.data_as(ctypes.c_void_p),
                                  coords.ctypes.data_as(ctypes.c_void_p), num.ctypes.data_as(ctypes.c_void_p),
                                  vnum.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise ValueError(lib.lv_last_error(h).decode())
        k = int(vnum[0])
        return {"voxels": voxels[:k], "coordinates": coords[:k], "num_points_per_voxel": num[:k]}

    cfg = _cfg(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    res = generate(fixture_nx4, cfg)
    v, c, n = voxel_oracle.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    assert np.array_equal(res["coordinates"], c) and np.array_equal(res["num_points_per_voxel"], n)
    assert np.array_equal(res["voxels"].view(np.uint32), v.view(np.uint32))
    # the two-phase form: arrays of exactly voxel_num rows
    vnum = np.zeros((1,), np.int32)
    offs = np.array([0, fixture_nx4.shape[0]], np.int64)
    assert lib.lv_voxelize_host_begin(h, ctypes.byref(cfg), None, fixture_nx4.ctypes.data_as(ctypes.c_void_p), 1,
                                      offs.ctypes.data_as(ctypes.c_void_p), vnum.ctypes.data_as(ctypes.c_void_p)) == 0
    k = int(vnum[0])
    vox, co, nu = np.empty((k, 60, 4), np.float32), np.empty((k, 3), np.int32), np.empty((k,), np.int32)
    assert lib.lv_voxelize_host_fetch(h, 0, k, vox.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p),
                                      nu.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(co, c) and np.array_equal(vox.view(np.uint32), v.view(np.uint32)) and np.array_equal(nu, n)
    # errors come back as codes with a message, never as an abort
    assert lib.lv_voxelize_host_fetch(h, 3, k, vox.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p),
                                      nu.ctypes.data_as(ctypes.c_void_p)) == -1
    assert b"frame 3" in lib.lv_last_error(h)


def test_section_6_block_filter_binding(raw, fixture_nx4):
    from oracle import voxel_oracle
    lib, h = raw
    vs, rg = (0.05, 0.05, 0.2), (-50, -50, -5, 50, 50, 3)
    cfg = _cfg(vs, rg, 5, 40000)
    flt = lv_block_filter(1, 8, 0.2, 2.0)
    V, T = 40000, 5
    voxels, coords = np.zeros((V, T, 4), np.float32), np.zeros((V, 3), np.int32)
    num, vnum = np.zeros((V,), np.int32), np.zeros((1,), np.int32)
    offs = np.array([0, fixture_nx4.shape[0]], np.int64)
    rc = lib.lv_voxelize_filtered_host(h, ctypes.byref(cfg), ctypes.byref(flt),
                                       fixture_nx4.ctypes.data_as(ctypes.c_void_p), 1,
                                       offs.ctypes.data_as(ctypes.c_void_p), voxels.ctypes.data_as(ctypes.c_void_p),
                                       coords.ctypes.data_as(ctypes.c_void_p), num.ctypes.data_as(ctypes.c_void_p),
                                       vnum.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.lv_last_error(h)
    v, c, n, _ = voxel_oracle.points_to_voxel_filtered(fixture_nx4, vs, rg, 5, 40000, 1, 8, 0.2, 2.0)
    k = int(vnum[0])
    assert k == v.shape[0] and np.array_equal(coords[:k], c) and np.array_equal(num[:k], n)
    assert np.array_equal(voxels[:k].view(np.uint32), v.view(np.uint32))


def test_section_7_target_binding(raw, golden_dir):
    lib, h = raw
    gold = np.load(os.path.join(golden_dir, "ref_draw_boxes.npz"))
    from oracle.gen_golden_draw import SCENES
    seed, n, shape, vs, zo, ext = SCENES[0]
    corners_f64, cls = synth.box_scene(seed, n, ext)
    colors_i32 = (cls + 1).astype(np.int32)
    box_offsets = np.array([0, n], dtype=np.int64)
    target_u8 = np.empty((1, shape[0], shape[1]), dtype=np.uint8)
    rc = lib.lv_draw_boxes_host(h, corners_f64.ctypes.data_as(ctypes.c_void_p), colors_i32.ctypes.data_as(ctypes.c_void_p),
                                1, box_offsets.ctypes.data_as(ctypes.c_void_p), (ctypes.c_int32 * 3)(*shape),
                                (ctypes.c_double * 3)(*vs), ctypes.c_double(zo),
                                target_u8.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.lv_last_error(h)
    assert np.array_equal(target_u8[0], gold["scene0"])
