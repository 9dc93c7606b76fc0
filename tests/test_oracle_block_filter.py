"""Block-filtering voxelizer oracle.  PARITY UNPINNED: the rule is spconv 1.x's
(source not in the reference tree; see oracle/voxel_oracle.c).  What CAN be checked on the CPU is
that two independent statements of the restated rule agree - the serial C loop that updates the
block z-ranges point by point, and a numpy statement that works on the voxelizer's outputs - and
the structural properties the rule implies."""
import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import voxel_oracle as vo

FHD_VS = (0.05, 0.05, 0.2)             # second/second/configs/nuscenes/all.fhd.config:5
FHD_RANGE = (-50, -50, -5, 50, 50, 3)  # the config's 100 m x 100 m extent (2000 x 2000 x 40 cells)


@pytest.mark.parametrize("bf,bs,lo,hi", [(1, 8, 0.2, 2.0), (8, 3, 0.1, 2.0), (4, 1, 0.3, None), (2, 5, 0.0, 1.0)])
def test_c_loop_matches_numpy_statement(bf, bs, lo, hi):
    pts = synth.c5_frame(3)
    v, c, n, keep = vo.points_to_voxel_filtered(pts, FHD_VS, FHD_RANGE, 3, 40000, bf, bs, lo, hi)
    v0, c0, n0 = vo.points_to_voxel(pts, FHD_VS, FHD_RANGE, 3, 40000)
    assert keep.shape[0] == c0.shape[0]
    k2 = vo.block_filter_numpy(v0, c0, n0, vo.grid_size(FHD_VS, FHD_RANGE), bf, bs, lo, hi)
    assert np.array_equal(keep, k2)
    assert 0 < keep.sum() < keep.shape[0]
    # survivors keep their first-come order and their contents
    assert np.array_equal(c, c0[keep]) and np.array_equal(n, n0[keep])
    assert np.array_equal(v.view(np.uint32), v0[keep].view(np.uint32))


def test_thresholds_are_monotone():
    pts = synth.c5_frame(1)
    keeps = [vo.points_to_voxel_filtered(pts, FHD_VS, FHD_RANGE, 3, 40000, 1, 8, lo, None)[3]
             for lo in (0.0, 0.2, 0.5, 1.0)]
    for a, b in zip(keeps, keeps[1:]):
        assert np.all(a | ~b)                 # a higher threshold only removes voxels
    flat = vo.points_to_voxel_filtered(pts, FHD_VS, FHD_RANGE, 3, 40000, 1, 8, 1e9, None)[3]
    assert flat.sum() == 0


def test_flat_ground_is_removed_and_a_pole_survives():
    rng = np.random.default_rng(7)
    ground = np.column_stack([rng.uniform(-20, 20, 4000), rng.uniform(-20, 20, 4000),
                              rng.normal(-1.5, 0.005, 4000), np.zeros(4000)]).astype(np.float32)
    pole = np.column_stack([np.full(50, 5.02), np.full(50, 5.02), np.linspace(-1.5, 0.3, 50),
                            np.zeros(50)]).astype(np.float32)
    pts = np.concatenate([ground, pole])
    v, c, n, keep = vo.points_to_voxel_filtered(pts, FHD_VS, FHD_RANGE, 3, 40000, 1, 8, 0.2, None)
    assert keep.sum() > 0
    cx = np.floor((5.02 + 50) / 0.05)
    # every survivor lies within the 8-block window of the pole's column
    assert np.all(np.abs(c[:, 2] - cx) <= 4) and np.all(np.abs(c[:, 1] - cx) <= 4)
