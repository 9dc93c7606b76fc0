"""Hard-voxelizer oracle (CPU).  TEST INFRASTRUCTURE - see oracle/__init__.py.

PARITY UNPINNED at the spconv boundary: the arithmetic of
``spconv.utils.VoxelGeneratorV2`` is not in /root/reference.  The algorithm
restated here is the in-tree sibling second/second/utils/simplevis.py:9-61,
plus the output contract of second/second/data/preprocess.py:305-317.  The
restatement IS pinned against that sibling: oracle/gen_golden.py executes the
reference's ``points_to_bev`` on the bundled sweep and stores its density map;
tests/test_oracle_voxel.py checks that this oracle's per-pillar counts and voxel
count reproduce it (coordinate rule, bounds, first-come ids, ``break``); five more
cases at three-dimensional grids with the cap hit early (counts per column and the
height of every (z, y, x) cell) come from oracle/gen_golden_simplevis3d.py, and
tests/test_gpu_voxel.py holds the GPU voxelizer to the same goldens directly.  What
stays unpinned is what only spconv knows: ``continue`` instead of ``break`` on
overflow, and the order of the points inside a voxel.

Two implementations of the same loop:
  * ``points_to_voxel_loop``  - pure Python, a line-for-line walk of the
    reference loop; small inputs only.
  * ``points_to_voxel``       - the C restatement (oracle/voxel_oracle.c via
    ctypes), checked against the loop version in the tests; this is also the
    CPU baseline timed by bench.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvoxel_oracle.so")
_lib = None

OVERFLOW_CONTINUE = 0
OVERFLOW_BREAK = 1


def build(force=False):
    """gcc -O2 oracle/voxel_oracle.c -> oracle/libvoxel_oracle.so (git-ignored)."""
    src = os.path.join(_HERE, "voxel_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math",
                               "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.lvo_points_to_voxel.restype = ctypes.c_int32
        lib.lvo_points_to_voxel.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.lvo_points_to_voxel_filtered.restype = ctypes.c_int32
        lib.lvo_points_to_voxel_filtered.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
            ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.lvo_grid_size.restype = None
        lib.lvo_grid_size.argtypes = [ctypes.c_void_p] * 3
        lib.lvo_bev_counts.restype = None
        lib.lvo_bev_counts.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
        _lib = lib
    return _lib


def grid_size(voxel_size, coors_range):
    """simplevis.py:27-30: np.round((hi - lo) / vs) in float32 -> int32, xyz."""
    voxel_size = np.asarray(voxel_size, dtype=np.float32)
    coors_range = np.asarray(coors_range, dtype=np.float32)
    g = (coors_range[3:] - coors_range[:3]) / voxel_size
    return np.round(g, 0, g).astype(np.int32)


def points_to_voxel_loop(points, voxel_size, coors_range, max_points, max_voxels,
                         overflow="continue"):
    """Pure-Python walk of simplevis.py:35-51 with the voxel outputs of
    preprocess.py:305-310.  Returns (voxels, coordinates, num_points_per_voxel)
    sliced to voxel_num."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    voxel_size = np.asarray(voxel_size, dtype=np.float32)
    coors_range = np.asarray(coors_range, dtype=np.float32)
    gs = grid_size(voxel_size, coors_range)
    voxelmap_shape = tuple(int(v) for v in gs[::-1])
    coor_to_voxelidx = -np.ones(shape=voxelmap_shape, dtype=np.int32)
    N, C = points.shape
    voxels = np.zeros((max_voxels, max_points, C), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    coor = np.zeros(shape=(3,), dtype=np.int32)
    voxel_num = 0
    for i in range(N):
        failed = False
        for j in range(3):
            c = np.floor((points[i, j] - coors_range[j]) / voxel_size[j])
            if not (c >= 0) or c >= gs[j]:
                failed = True
                break
            coor[2 - j] = c
        if failed:
            continue
        voxelidx = coor_to_voxelidx[coor[0], coor[1], coor[2]]
        if voxelidx == -1:
            voxelidx = voxel_num
            if voxel_num >= max_voxels:
                if overflow == "break":
                    break
                continue
            voxel_num += 1
            coor_to_voxelidx[coor[0], coor[1], coor[2]] = voxelidx
            coors[voxelidx] = coor
        n = num[voxelidx]
        if n < max_points:
            voxels[voxelidx, n] = points[i]
            num[voxelidx] += 1
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]


class VoxelOracle:
    """Stateful C oracle that keeps the dense coor->voxelidx map between calls
    (allocated once, touched cells reset - the spconv strategy)."""

    def __init__(self, voxel_size, coors_range, max_points, max_voxels):
        self.voxel_size = np.asarray(voxel_size, dtype=np.float32).copy()
        self.coors_range = np.asarray(coors_range, dtype=np.float32).copy()
        self.max_points = int(max_points)
        self.max_voxels = int(max_voxels)
        self.grid_size = grid_size(self.voxel_size, self.coors_range)
        self._map = -np.ones(int(np.prod(self.grid_size.astype(np.int64))), dtype=np.int32)
        self.last_kept_points = 0

    def generate(self, points, max_voxels=None, overflow="continue", padded=False):
        lib = _load()
        points = np.ascontiguousarray(points, dtype=np.float32)
        n, c = points.shape
        mv = self.max_voxels if max_voxels is None else int(max_voxels)
        voxels = np.zeros((mv, self.max_points, c), dtype=np.float32)
        coors = np.zeros((mv, 3), dtype=np.int32)
        num = np.zeros((mv,), dtype=np.int32)
        kept = ctypes.c_int64(0)
        vn = lib.lvo_points_to_voxel(
            points.ctypes.data, n, c, self.voxel_size.ctypes.data, self.coors_range.ctypes.data,
            self.max_points, mv, OVERFLOW_BREAK if overflow == "break" else OVERFLOW_CONTINUE,
            self._map.ctypes.data, voxels.ctypes.data, coors.ctypes.data, num.ctypes.data,
            ctypes.addressof(kept))
        self.last_kept_points = kept.value
        if padded:
            return voxels, coors, num, vn
        return voxels[:vn], coors[:vn], num[:vn]


def points_to_voxel(points, voxel_size, coors_range, max_points, max_voxels,
                    overflow="continue"):
    """One-shot C oracle -> (voxels, coordinates, num_points_per_voxel)."""
    return VoxelOracle(voxel_size, coors_range, max_points, max_voxels).generate(
        points, overflow=overflow)


def points_to_voxel_filtered(points, voxel_size, coors_range, max_points, max_voxels, block_factor=8,
                             block_size=3, height_threshold=0.1, height_high_threshold=2.0):
    """Block-filtering voxelizer, C oracle (lvo_points_to_voxel_filtered) + the compaction spconv does
    in Python (`coors[voxel_mask]`).  PARITY UNPINNED (spconv 1.x, source absent; see voxel_oracle.c).
    height_high_threshold=None -> +inf (the rule before spconv 1.2).
    Returns (voxels, coordinates, num_points_per_voxel, voxel_mask) - the first three filtered, the
    mask over the unfiltered voxels."""
    lib = _load()
    points = np.ascontiguousarray(points, dtype=np.float32)
    voxel_size = np.asarray(voxel_size, dtype=np.float32).copy()
    coors_range = np.asarray(coors_range, dtype=np.float32).copy()
    gs = grid_size(voxel_size, coors_range)
    assert gs[0] % block_factor == 0 and gs[1] % block_factor == 0 and block_size > 0
    n, c = points.shape
    cmap = -np.ones(int(np.prod(gs.astype(np.int64))), dtype=np.int32)
    voxels = np.zeros((max_voxels, max_points, c), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    mask = np.zeros((max_voxels,), dtype=np.int32)
    nb = int(gs[1] // block_factor) * int(gs[0] // block_factor)
    mins = np.empty(nb, dtype=np.float32)
    maxs = np.empty(nb, dtype=np.float32)
    hi = np.inf if height_high_threshold is None else height_high_threshold
    vn = lib.lvo_points_to_voxel_filtered(
        points.ctypes.data, n, c, voxel_size.ctypes.data, coors_range.ctypes.data, max_points,
        max_voxels, block_factor, block_size, ctypes.c_float(height_threshold), ctypes.c_float(hi),
        cmap.ctypes.data, voxels.ctypes.data, coors.ctypes.data, num.ctypes.data, mask.ctypes.data,
        mins.ctypes.data, maxs.ctypes.data)
    keep = mask[:vn].astype(bool)
    return voxels[:vn][keep], coors[:vn][keep], num[:vn][keep], keep


def block_filter_numpy(voxels, coors, num, grid_xyz, block_factor, block_size, height_threshold,
                       height_high_threshold=2.0):
    """Independent numpy statement of the block filter on voxelizer OUTPUTS (the stored points of a
    voxel are exactly the points that updated its block's z range): returns the keep mask."""
    voxels = np.asarray(voxels, dtype=np.float32)
    V, T, _ = voxels.shape
    BH, BW = int(grid_xyz[1]) // block_factor, int(grid_xyz[0]) // block_factor
    mins = np.full((BH, BW), 99999999, dtype=np.float32)
    maxs = np.full((BH, BW), -99999999, dtype=np.float32)
    live = np.arange(T)[None, :] < np.asarray(num)[:, None]
    z = voxels[:, :, 2]
    by = np.asarray(coors)[:, 1] // block_factor
    bx = np.asarray(coors)[:, 2] // block_factor
    zmin = np.where(live, z, np.float32(np.inf)).min(axis=1) if V else np.zeros(0, np.float32)
    zmax = np.where(live, z, np.float32(-np.inf)).max(axis=1) if V else np.zeros(0, np.float32)
    np.minimum.at(mins, (by, bx), zmin.astype(np.float32))
    np.maximum.at(maxs, (by, bx), zmax.astype(np.float32))
    keep = np.zeros(V, dtype=bool)
    hi = np.float32(np.inf) if height_high_threshold is None else np.float32(height_high_threshold)
    for i in range(V):
        y0, y1 = max(0, by[i] - block_size // 2), min(BH, by[i] + block_size - block_size // 2)
        x0, x1 = max(0, bx[i] - block_size // 2), min(BW, bx[i] + block_size - block_size // 2)
        h = np.float32(maxs[y0:y1, x0:x1].max() - mins[y0:y1, x0:x1].min())
        keep[i] = (h > np.float32(height_threshold)) and (h < hi)
    return keep


def bev_counts_c(points_nx, shape, voxel_size, z_offset):
    """C cross-check of oracle.bev_oracle.create_voxel_pointcloud (raw counts)."""
    from . import bev_oracle
    lib = _load()
    pts = np.ascontiguousarray(points_nx, dtype=np.float32)
    tm = bev_oracle.create_transformation_matrix_to_voxel_space(shape, voxel_size, (0, 0, z_offset))
    m = np.ascontiguousarray(np.diag(tm)[:3], dtype=np.float64)
    t = np.ascontiguousarray(tm[:3, 3], dtype=np.float64)
    shp = np.asarray(shape, dtype=np.int32)
    counts = np.zeros(tuple(int(s) for s in shape), dtype=np.uint32)
    lib.lvo_bev_counts(pts.ctypes.data, pts.shape[0], pts.shape[1], m.ctypes.data,
                       t.ctypes.data, shp.ctypes.data, counts.ctypes.data)
    return counts
