"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol
include/lyft_voxel.h declares, the ctypes table covers them all, and - without a
GPU - the product path fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge
from lyft3d_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(nat.LIB_PATH):
        ge.build()
    return nat.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "lyft_voxel.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert sorted(nat.SIGNATURES.keys()) == syms


def test_version(lib):
    assert lib.lv_abi_version() == 1
    assert b"sm_100a" in lib.lv_version_string()


def test_struct_layout_matches_header():
    # 3 + 6 floats, 5 int32
    assert ctypes.sizeof(nat.VoxelConfig) == 4 * (3 + 6 + 5)


def test_pure_host_entry_points(lib):
    cfg = nat.VoxelConfig()
    cfg.voxel_size[:] = [0.05, 0.05, 0.1]
    cfg.coors_range[:] = [0.0, -32.0, -3.0, 52.8, 32.0, 1.0]
    g = (ctypes.c_int32 * 3)()
    assert lib.lv_voxel_grid_size(ctypes.byref(cfg), g) == 0
    assert list(g) == [1056, 1280, 40]
    assert lib.lv_pillar_out_channels(4, 0, 0) == 9
    assert lib.lv_pillar_out_channels(4, 1, 1) == 10
    assert lib.lv_pillar_out_channels(4, 2, 0) == 8
    assert lib.lv_pillar_out_channels(4, 3, 0) == 9
    assert lib.lv_pillar_out_channels(2, 0, 0) < 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.lv_device_count() == 0
    h = ctypes.c_void_p()
    rc = lib.lv_create(0, ctypes.byref(h))
    assert rc == nat.LV_E_NODEVICE and not h.value
    assert b"no CPU fallback" in lib.lv_last_error(None)
    from lyft3d_b200 import bev, voxel_generator as vg
    pts = np.zeros((4, 10), np.float32)
    with pytest.raises(nat.LyftVoxelError):
        bev.create_voxel_pointcloud(pts, (8, 8, 3), (1, 1, 1), 0)
    # every drop-in fails the same way: generate (plain and block-filtering), the target raster
    gen = vg.VoxelGeneratorV2([0.25, 0.25, 8], [-50, -50, -5, 50, 50, 3], 60, max_voxels=100)
    with pytest.raises(nat.LyftVoxelError):
        gen.generate(np.zeros((10, 4), np.float32))
    flt = vg.VoxelGeneratorV2([0.25, 0.25, 8], [-50, -50, -5, 50, 50, 3], 5, max_voxels=100, block_filtering=True,
                              block_factor=1, block_size=8, height_threshold=0.2)
    with pytest.raises(nat.LyftVoxelError):
        flt.generate_multi_gpu(np.zeros((10, 4), np.float32))
    with pytest.raises(nat.LyftVoxelError):
        bev.rasterize_targets(np.zeros((1, 3, 4)), np.ones(1, np.int32), np.array([0, 1]), (8, 8, 3), (1, 1, 1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lyft-3d-object-detection_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
