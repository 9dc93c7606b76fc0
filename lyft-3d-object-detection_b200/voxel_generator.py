"""Hard-voxelizer drop-ins (reference boundary: spconv.utils.VoxelGeneratorV2 as used by
second/second/builder/voxel_builder.py:23-32 and second/second/data/preprocess.py:299-317).

    VoxelGeneratorV2(voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                     full_mean=False, block_filtering=False, block_factor=..., block_size=...,
                     height_threshold=...)
        .generate(points, max_voxels=None)       -> {"voxels","coordinates","num_points_per_voxel"[, ...]}
        .generate_multi_gpu(points, max_voxels)  -> same, padded to max_voxels, + "voxel_num"
        .voxel_size .point_cloud_range .grid_size .max_num_points_per_voxel
    VoxelGenerator(...).generate(points, max_voxels) -> (voxels, coordinates, num_points_per_voxel)
    points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000)
        legacy tuple API (second/second/kittiviewer/viewer.py:389-396)

The generator object is a plain picklable Python object holding only its
configuration (it is captured in functools.partial for DataLoader workers and stored
on the network); the native handle is created lazily per process.
numpy in -> numpy out; CUDA tensor in -> CUDA tensors out (no host copies).
``overflow`` selects what happens when max_voxels is hit: "continue" (spconv >= 1.1,
default) or "break" (the in-tree rule, simplevis.py:48-49) - SURVEY.md F6.
"""
import ctypes

import numpy as np

from . import _native as nat


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _make_config(voxel_size, coors_range, max_points, max_voxels, num_features, overflow, zero_tail):
    cfg = nat.VoxelConfig()
    cfg.voxel_size[:] = [float(v) for v in voxel_size]
    cfg.coors_range[:] = [float(v) for v in coors_range]
    cfg.max_points = int(max_points)
    cfg.max_voxels = int(max_voxels)
    cfg.num_features = int(num_features)
    cfg.overflow_mode = nat.OVERFLOW_BREAK if overflow == "break" else nat.OVERFLOW_CONTINUE
    cfg.zero_tail = int(bool(zero_tail))
    return cfg


def _make_filter(block_filter):
    """(block_factor, block_size, height_threshold, height_high_threshold | None) -> struct lv_block_filter."""
    bf, bs, lo, hi = block_filter
    flt = nat.BlockFilter()
    flt.block_factor = int(bf)
    flt.block_size = int(bs)
    flt.height_threshold = float(lo)
    flt.height_high_threshold = float("inf") if hi is None else float(hi)
    return flt


def voxelize_frames(points, frame_offsets, voxel_size, coors_range, max_points, max_voxels,
                    overflow="continue", zero_tail=True, handle=None, block_filter=None):
    """Batched voxelization of F independent clouds.

    points (N_total, C) float32 rows (numpy or CUDA tensor); frame_offsets int64 (F+1).
    Returns padded per-frame arrays: voxels (F,V,T,C), coordinates (F,V,3) zyx,
    num_points_per_voxel (F,V), voxel_num (F).  With zero_tail=False rows at and
    beyond voxel_num[f] are unspecified.  block_filter = (block_factor, block_size,
    height_threshold, height_high_threshold) runs the spconv block-filtering variant
    (lv_voxelize_filtered; `continue` overflow rule only).
    """
    if overflow not in ("continue", "break"):
        raise ValueError("overflow must be 'continue' or 'break'")
    lib = nat.load()
    offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    F = offs.shape[0] - 1
    V, T = int(max_voxels), int(max_points)
    if _is_cuda_tensor(points):
        import torch
        if points.dim() != 2 or points.dtype != torch.float32:
            raise ValueError("points must be a (N, C) float32 tensor")
        pts = points.contiguous()
        C = pts.shape[1]
        dev = pts.device
        cfg = _make_config(voxel_size, coors_range, T, V, C, overflow, zero_tail)
        voxels = torch.empty((F, V, T, C), dtype=torch.float32, device=dev)
        coords = torch.empty((F, V, 3), dtype=torch.int32, device=dev)
        num = torch.empty((F, V), dtype=torch.int32, device=dev)
        vnum = torch.empty((F,), dtype=torch.int32, device=dev)
        h = handle or nat.get_handle(dev.index)
        with torch.cuda.device(dev):
            if block_filter is not None:
                flt = _make_filter(block_filter)
                nat.check(lib.lv_voxelize_filtered(h.ptr, ctypes.byref(cfg), ctypes.byref(flt), pts.data_ptr(), F,
                                                   offs.ctypes.data, voxels.data_ptr(), coords.data_ptr(),
                                                   num.data_ptr(), vnum.data_ptr(), None,
                                                   nat.current_stream_ptr(dev)))
            else:
                nat.check(lib.lv_voxelize(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data,
                                          voxels.data_ptr(), coords.data_ptr(), num.data_ptr(), vnum.data_ptr(),
                                          nat.current_stream_ptr(dev)))
        return voxels, coords, num, vnum
    pts = np.ascontiguousarray(points, dtype=np.float32)
    if pts.ndim != 2:
        raise ValueError("points must be (N, C)")
    C = pts.shape[1]
    cfg = _make_config(voxel_size, coors_range, T, V, C, overflow, zero_tail)
    alloc = np.zeros if not zero_tail else np.empty
    voxels = alloc((F, V, T, C), dtype=np.float32)
    coords = alloc((F, V, 3), dtype=np.int32)
    num = alloc((F, V), dtype=np.int32)
    vnum = np.zeros((F,), dtype=np.int32)
    h = handle or nat.get_handle()
    if block_filter is not None:
        flt = _make_filter(block_filter)
        nat.check(lib.lv_voxelize_filtered_host(h.ptr, ctypes.byref(cfg), ctypes.byref(flt), pts.ctypes.data, F,
                                                offs.ctypes.data, voxels.ctypes.data, coords.ctypes.data,
                                                num.ctypes.data, vnum.ctypes.data))
    else:
        nat.check(lib.lv_voxelize_host(h.ptr, ctypes.byref(cfg), pts.ctypes.data, F, offs.ctypes.data,
                                       voxels.ctypes.data, coords.ctypes.data, num.ctypes.data, vnum.ctypes.data))
    return voxels, coords, num, vnum


def voxelize_concat_frames(points, frame_offsets, voxel_size, coors_range, max_points, max_voxels,
                           overflow="continue", capacity=None, handle=None):
    """Batched voxelization straight into the layout ``merge_second_batch`` produces
    (second/second/data/preprocess.py:21-55; lv_voxelize_concat): points (N_total, C) float32 CUDA
    tensor, frame_offsets int64 (F+1).  Returns CUDA tensors voxels (sum V, T, C), coordinates
    (sum V, 4) [b, z, y, x], num_points_per_voxel (sum V) and voxel_num (F).  One host read (the
    total), like the reference's ``voxels.shape[0]`` (preprocess.py:310)."""
    import torch
    if overflow not in ("continue", "break"):
        raise ValueError("overflow must be 'continue' or 'break'")
    if not (_is_cuda_tensor(points) and points.dim() == 2 and points.dtype == torch.float32):
        raise ValueError("points must be a (N, C) float32 CUDA tensor")
    lib = nat.load()
    pts = points.contiguous()
    offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    F = offs.shape[0] - 1
    V, T, C = int(max_voxels), int(max_points), pts.shape[1]
    cap = int(capacity if capacity is not None else F * V)
    dev = pts.device
    cfg = _make_config(voxel_size, coors_range, T, V, C, overflow, False)
    voxels = torch.empty((cap, T, C), dtype=torch.float32, device=dev)
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    vnum = torch.empty((F,), dtype=torch.int32, device=dev)
    voff = torch.zeros((F + 1,), dtype=torch.int64, device=dev)
    h = handle or nat.get_handle(dev.index)
    with torch.cuda.device(dev):
        nat.check(lib.lv_voxelize_concat(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, cap,
                                         voxels.data_ptr(), coords.data_ptr(), num.data_ptr(), vnum.data_ptr(),
                                         voff.data_ptr(), nat.current_stream_ptr(dev)))
    total = int(voff[F].item()) if F > 0 else 0
    if total > cap:
        raise nat.LyftVoxelError(nat.LV_E_INVALID, "voxel capacity %d exceeded (%d rows)" % (cap, total))
    return voxels[:total], coords[:total], num[:total], vnum


def _pinned_empty(shape, dtype):
    """numpy array over pinned host memory (owned by a torch tensor the array keeps alive)."""
    import torch
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    t = torch.empty((n,), dtype=torch.float32 if dtype == np.float32 else torch.int32, pin_memory=True)
    return t.numpy().reshape(shape)


def voxelize_host_single(points, voxel_size, coors_range, max_points, max_voxels, overflow="continue",
                         padded=False, block_filter=None, handle=None):
    """One cloud, host numpy in and out, through lv_voxelize_host_begin / _fetch: the arrays that come
    back have exactly voxel_num rows (``padded=False``, what ``generate`` returns) or max_voxels rows of which
    only the first voxel_num crossed the bus (``padded=True``, the generate_multi_gpu layout: the rest is
    the zero fill of a fresh allocation).  Returns (voxels, coordinates, num_points_per_voxel, voxel_num)."""
    lib = nat.load()
    pts = np.ascontiguousarray(points, dtype=np.float32)
    if pts.ndim != 2:
        raise ValueError("points must be (N, C)")
    C = pts.shape[1]
    V, T = int(max_voxels), int(max_points)
    cfg = _make_config(voxel_size, coors_range, T, V, C, overflow, False)
    offs = np.array([0, pts.shape[0]], dtype=np.int64)
    vnum = np.zeros((1,), dtype=np.int32)
    h = handle or nat.get_handle()
    flt = _make_filter(block_filter) if block_filter is not None else None
    nat.check(lib.lv_voxelize_host_begin(h.ptr, ctypes.byref(cfg), ctypes.byref(flt) if flt is not None else None,
                                         pts.ctypes.data, 1, offs.ctypes.data, vnum.ctypes.data))
    k = int(vnum[0])
    rows = V if padded else k
    if padded:
        voxels = np.zeros((rows, T, C), dtype=np.float32)
        coords = np.zeros((rows, 3), dtype=np.int32)
        num = np.zeros((rows,), dtype=np.int32)
    else:
        # the result arrays live in PINNED host memory (torch's caching host allocator hands the blocks out and takes
        # them back when the arrays die): the 8 MB of a sweep's voxels cross PCIe in one DMA instead of through the
        # driver's pageable staging - 0.68 -> 0.34 ms per call for the bundled sweep (tools/dropin_latency.py)
        voxels, coords, num = _pinned_empty((rows, T, C), np.float32), _pinned_empty((rows, 3), np.int32), \
            _pinned_empty((rows,), np.int32)
    nat.check(lib.lv_voxelize_host_fetch(h.ptr, 0, k, voxels.ctypes.data, coords.ctypes.data, num.ctypes.data))
    return voxels, coords, num, k


class VoxelGeneratorV2:
    """spconv.utils.VoxelGeneratorV2-compatible generator (see module docstring)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, full_mean=False,
                 block_filtering=False, block_factor=8, block_size=3, height_threshold=0.1,
                 height_high_threshold=2.0, overflow="continue"):
        assert full_mean is False, "full_mean is not supported (spconv asserts the same)"
        if overflow not in ("continue", "break"):
            raise ValueError("overflow must be 'continue' or 'break'")
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = grid_size
        self._full_mean = full_mean
        self._overflow = overflow
        # block filtering (voxel_builder.py:28-31; all.fhd.config:9-12).  The rule is spconv's
        # (PARITY UNPINNED, see include/lyft_voxel.h); spconv asserts the same divisibility.
        self._block_filter = None
        if block_filtering:
            if overflow != "continue":
                raise ValueError("block_filtering exists only with the 'continue' overflow rule")
            assert block_size > 0
            assert grid_size[0] % block_factor == 0
            assert grid_size[1] % block_factor == 0
            self._block_filter = (int(block_factor), int(block_size), float(height_threshold), height_high_threshold)

    # -- the V2 dict API (preprocess.py:305-317, inference.py:68-69)
    def _run(self, points, max_voxels, padded):
        mv = self._max_voxels if max_voxels is None else int(max_voxels)
        n = points.shape[0]
        if not _is_cuda_tensor(points):
            return voxelize_host_single(points, self._voxel_size, self._point_cloud_range, self._max_num_points, mv,
                                        overflow=self._overflow, padded=padded, block_filter=self._block_filter)
        voxels, coords, num, vnum = voxelize_frames(
            points, np.array([0, n], dtype=np.int64), self._voxel_size, self._point_cloud_range,
            self._max_num_points, mv, overflow=self._overflow, zero_tail=padded,
            block_filter=self._block_filter)
        k = int(vnum[0])
        return voxels[0], coords[0], num[0], k

    def generate(self, points, max_voxels=None):
        voxels, coords, num, k = self._run(points, max_voxels, padded=False)
        return {"voxels": voxels[:k], "coordinates": coords[:k], "num_points_per_voxel": num[:k],
                "voxel_num": k}

    def generate_multi_gpu(self, points, max_voxels=None):
        voxels, coords, num, k = self._run(points, max_voxels, padded=True)
        return {"voxels": voxels, "coordinates": coords, "num_points_per_voxel": num, "voxel_num": k}

    def generate_batch(self, clouds, max_voxels=None):
        """``[self.generate(c, max_voxels) for c in clouds]`` for host numpy clouds in ONE pass: one copy of all the
        points to the device, one run of the kernels over all the clouds (lv_voxelize_host_begin), then the rows of every
        cloud (lv_voxelize_host_fetch) - the launch and synchronisation cost of a call is paid once per batch instead of
        once per sweep."""
        clouds = [np.ascontiguousarray(c, dtype=np.float32) for c in clouds]
        if not clouds:
            return []
        if any(c.ndim != 2 or c.shape[1] != clouds[0].shape[1] for c in clouds):
            raise ValueError("clouds must be (N_i, C) arrays with the same C")
        mv = self._max_voxels if max_voxels is None else int(max_voxels)
        lib = nat.load()
        C, T, F = clouds[0].shape[1], self._max_num_points, len(clouds)
        pts = np.concatenate(clouds, axis=0)
        offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int64)
        cfg = _make_config(self._voxel_size, self._point_cloud_range, T, mv, C, self._overflow, False)
        vnum = np.zeros((F,), dtype=np.int32)
        h = nat.get_handle()
        flt = _make_filter(self._block_filter) if self._block_filter is not None else None
        nat.check(lib.lv_voxelize_host_begin(h.ptr, ctypes.byref(cfg), ctypes.byref(flt) if flt is not None else None,
                                             pts.ctypes.data, F, offs.ctypes.data, vnum.ctypes.data))
        out = []
        for f in range(F):
            k = int(vnum[f])
            voxels, coords, num = _pinned_empty((k, T, C), np.float32), _pinned_empty((k, 3), np.int32), \
                _pinned_empty((k,), np.int32)
            nat.check(lib.lv_voxelize_host_fetch(h.ptr, f, k, voxels.ctypes.data, coords.ctypes.data, num.ctypes.data))
            out.append({"voxels": voxels, "coordinates": coords, "num_points_per_voxel": num, "voxel_num": k})
        return out

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


class VoxelGenerator(VoxelGeneratorV2):
    """V1-style generator: generate() returns the tuple (voxels, coordinates, num_points_per_voxel)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, overflow="break"):
        super().__init__(voxel_size, point_cloud_range, max_num_points, max_voxels, overflow=overflow)

    def generate(self, points, max_voxels=None):
        res = super().generate(points, max_voxels)
        return res["voxels"], res["coordinates"], res["num_points_per_voxel"]


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000,
                    overflow="break"):
    """Legacy tuple API (numba point_cloud_ops.points_to_voxel; call sites
    second/second/kittiviewer/viewer.py:389-396,509-516,573-579).  reverse_index=True
    returns zyx coordinates (the only form the reference uses); False returns xyz."""
    n = points.shape[0]
    voxels, coords, num, vnum = voxelize_frames(points, np.array([0, n], dtype=np.int64), voxel_size, coors_range,
                                                max_points, max_voxels, overflow=overflow, zero_tail=False)
    k = int(vnum[0])
    co = coords[0][:k]
    if not reverse_index:
        co = co.flip(-1) if _is_cuda_tensor(co) else co[:, ::-1].copy()
    return voxels[0][:k], co, num[0][:k]


def unpad_multigpu_batch(voxels, num_points, coordinates, num_voxels):
    """The un-padding at the top of VoxelNet.forward (voxelnet.py:346-358) on the device:
    padded (B,V,T,C) / (B,V) / (B,V,4) CUDA tensors + num_voxels (B) -> concatenated
    (sum,T,C) / (sum) / (sum,4).  One host read (the total), like the reference's
    ``example["num_voxels"].cpu()``."""
    import torch
    lib = nat.load()
    if not (voxels.is_cuda and voxels.dtype == torch.float32):
        raise ValueError("voxels must be a float32 CUDA tensor")
    voxels = voxels.contiguous()
    B, V, T, C = voxels.shape
    num = num_points.to(torch.int32).contiguous()
    co = coordinates.to(torch.int32).contiguous()
    nv = num_voxels.to(device=voxels.device, dtype=torch.int32).reshape(-1).contiguous()
    cols = co.shape[-1]
    dev = voxels.device
    out_v = torch.empty((B * V, T, C), dtype=torch.float32, device=dev)
    out_n = torch.empty((B * V,), dtype=torch.int32, device=dev)
    out_c = torch.empty((B * V, cols), dtype=torch.int32, device=dev)
    total = torch.zeros((1,), dtype=torch.int64, device=dev)
    h = nat.get_handle(dev.index)
    with torch.cuda.device(dev):
        nat.check(lib.lv_unpad_batch(h.ptr, voxels.data_ptr(), num.data_ptr(), co.data_ptr(), nv.data_ptr(), B, V, T, C,
                                     cols, out_v.data_ptr(), out_n.data_ptr(), out_c.data_ptr(), total.data_ptr(),
                                     nat.current_stream_ptr(dev)))
    n = int(total.item())
    return out_v[:n], out_n[:n], out_c[:n]


def voxelize_mean_frames(points, frame_offsets, voxel_size, coors_range, max_points, max_voxels,
                         num_features_out=4, overflow="continue", capacity=None, handle=None):
    """Voxelization + SimpleVoxel mean VFE in one pipeline (lv_voxelize_mean_concat): the (V,T,C)
    voxel tensor of ``generate`` + ``SimpleVoxel.forward`` (voxel_encoder.py:219-225) is never
    written.  points: (N_total, 4) float32 CUDA tensor; returns CUDA tensors
    (mean (sum V, num_features_out), coordinates (sum V, 4) [b,z,y,x], num_points (sum V),
    voxel_num (F))."""
    import torch
    lib = nat.load()
    if not (_is_cuda_tensor(points) and points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 4):
        raise ValueError("points must be a (N, 4) float32 CUDA tensor")
    pts = points.contiguous()
    offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    F = offs.shape[0] - 1
    cap = int(capacity if capacity is not None else F * int(max_voxels))
    dev = pts.device
    cfg = _make_config(voxel_size, coors_range, int(max_points), int(max_voxels), 4, overflow, False)
    mean = torch.empty((cap, num_features_out), dtype=torch.float32, device=dev)
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    vnum = torch.empty((F,), dtype=torch.int32, device=dev)
    voff = torch.empty((F + 1,), dtype=torch.int64, device=dev)
    h = handle or nat.get_handle(dev.index)
    with torch.cuda.device(dev):
        nat.check(lib.lv_voxelize_mean_concat(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, cap,
                                              int(num_features_out), mean.data_ptr(), coords.data_ptr(), num.data_ptr(),
                                              vnum.data_ptr(), voff.data_ptr(), nat.current_stream_ptr(dev)))
    total = min(int(voff[F].item()), cap)
    return mean[:total], coords[:total], num[:total], vnum
