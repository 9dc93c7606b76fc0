"""ctypes binding of liblyftvoxel_b200.so (include/lyft_voxel.h).

The library is the product; there is NO CPU fallback.  Loading fails loudly when
the .so has not been built (``python -c 'import __graft_entry__ as g; g.build()'``
or ``lyft-3d-object-detection_b200/csrc/build.sh``) and ``get_handle`` fails
loudly when no B200 is visible.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "liblyftvoxel_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

LV_OK = 0
LV_E_INVALID = -1
LV_E_CUDA = -2
LV_E_NOMEM = -3
LV_E_UNSUPPORTED = -4
LV_E_NODEVICE = -5

OVERFLOW_CONTINUE = 0
OVERFLOW_BREAK = 1

PILLAR_VARIANTS = {"pfn": 0, "old": 1, "radius": 2, "radius_height": 3}


class LyftVoxelError(Exception):
    """Raised for every non-zero return code of the C ABI."""

    def __init__(self, code, message):
        super().__init__("liblyftvoxel_b200: %s (code %d)" % (message, code))
        self.code = code


class VoxelConfig(ctypes.Structure):
    """struct lv_voxel_config."""
    _fields_ = [("voxel_size", ctypes.c_float * 3),
                ("coors_range", ctypes.c_float * 6),
                ("max_points", ctypes.c_int32),
                ("max_voxels", ctypes.c_int32),
                ("num_features", ctypes.c_int32),
                ("overflow_mode", ctypes.c_int32),
                ("zero_tail", ctypes.c_int32)]


class BlockFilter(ctypes.Structure):
    """struct lv_block_filter."""
    _fields_ = [("block_factor", ctypes.c_int32),
                ("block_size", ctypes.c_int32),
                ("height_threshold", ctypes.c_float),
                ("height_high_threshold", ctypes.c_float)]


_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_f32 = ctypes.c_float
_f64 = ctypes.c_double

# name -> (restype, argtypes); every symbol include/lyft_voxel.h declares
SIGNATURES = {
    "lv_abi_version": (ctypes.c_int, []),
    "lv_version_string": (ctypes.c_char_p, []),
    "lv_device_count": (ctypes.c_int, []),
    "lv_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "lv_destroy": (ctypes.c_int, [_vp]),
    "lv_last_error": (ctypes.c_char_p, [_vp]),
    "lv_workspace_bytes": (_i64, [_vp]),
    "lv_launch_count": (_i64, [_vp]),
    "lv_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "lv_host_free": (ctypes.c_int, [_vp]),
    "lv_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, _i64]),
    "lv_bev_rasterize": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _f64, _f32,
                                        _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_bev_rasterize_host": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _f64, _f32,
                                             _vp, _vp, _vp, _vp, _vp]),
    "lv_draw_boxes": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _f64, _vp, _vp]),
    "lv_draw_boxes_host": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _f64, _vp]),
    "lv_png_max_bytes": (_i64, [_i32, _i32, _i32]),
    "lv_png_encode": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "lv_png_encode_host": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "lv_bev_normalize": (ctypes.c_int, [_vp, _vp, _i64, _f32, _vp, _vp]),
    "lv_transform_points": (ctypes.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "lv_ingest_sweeps": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _vp, _vp]),
    "lv_voxel_grid_size": (ctypes.c_int, [ctypes.POINTER(VoxelConfig), _vp]),
    "lv_voxelize": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_concat": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp,
                                          _vp, _vp]),
    "lv_pillarize_concat": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _i64, _f32, _f32, _f32,
                                           _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_mean_concat": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _i64, _i32, _vp, _vp,
                                               _vp, _vp, _vp, _vp]),
    "lv_unpad_batch": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_host": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxel_block_filter": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), ctypes.POINTER(BlockFilter), _i32, _vp,
                                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_filtered": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), ctypes.POINTER(BlockFilter), _vp, _i32,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_filtered_host": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), ctypes.POINTER(BlockFilter), _vp,
                                                 _i32, _vp, _vp, _vp, _vp, _vp]),
    "lv_voxelize_host_begin": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), ctypes.POINTER(BlockFilter), _vp, _i32,
                                              _vp, _vp]),
    "lv_voxelize_host_fetch": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "lv_pillar_out_channels": (ctypes.c_int, [_i32, _i32, _i32]),
    "lv_pillar_decorate": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32,
                                          _i32, _i32, _vp, _vp]),
    "lv_pillar_pfn": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                     _vp, _vp, _vp, _i32, _vp, _vp]),
    "lv_pillarize_pfn_concat": (ctypes.c_int, [_vp, ctypes.POINTER(VoxelConfig), _vp, _i32, _vp, _i64, _f32, _f32,
                                               _f32, _f32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp,
                                               _vp, _vp]),
    "lv_pillar_pfn_moments": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                             _vp, _vp]),
    "lv_pillar_pfn_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                              _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "lv_pillar_pfn_train_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                                   _vp, _vp, _vp, _f64, _i32, _vp, _vp, _vp, _vp]),
    "lv_pillar_pfn_train_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                                    _vp, _vp, _f64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lv_pillar_scatter_backward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lv_pillar_scatter": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lv_pillar_scatter_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lv_pillar_decorate_half": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32,
                                               _vp, _vp]),
    "lv_pillar_scatter_half": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lv_voxel_mean": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
}

_lib = None
_lock = threading.RLock()  # get_handle -> Handle() -> load() re-enters
_handles = {}  # (pid, device) -> handle


def load():
    """dlopen the CUDA library (no GPU needed for this step)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise LyftVoxelError(
                    LV_E_UNSUPPORTED,
                    "%s is not built; run csrc/build.sh (nvcc, sm_100a). There is no CPU fallback" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError = ABI mismatch, fail loudly
                fn.restype = res
                fn.argtypes = args
            if lib.lv_abi_version() != 1:
                raise LyftVoxelError(LV_E_UNSUPPORTED, "ABI version mismatch")
            _lib = lib
    return _lib


def last_error(handle=None):
    return load().lv_last_error(handle).decode("utf-8", "replace")


def check(rc, handle=None):
    if rc != LV_OK:
        raise LyftVoxelError(rc, last_error(handle))
    return rc


class Handle:
    """Owns one lv_handle (one per process and device, created lazily)."""

    def __init__(self, device):
        lib = load()
        h = _vp()
        check(lib.lv_create(int(device), ctypes.byref(h)))
        self.device = int(device)
        self.ptr = h
        self.pid = os.getpid()

    def launches(self):
        return int(load().lv_launch_count(self.ptr))

    def workspace_bytes(self):
        return int(load().lv_workspace_bytes(self.ptr))

    def set_option(self, name, value):
        check(load().lv_set_option(self.ptr, name.encode(), int(value)))

    def close(self):
        if self.ptr and self.pid == os.getpid():
            load().lv_destroy(self.ptr)
        self.ptr = None


def get_handle(device=None):
    """Handle for `device` (default: torch's current CUDA device), created on
    first use in the calling process - CUDA contexts do not survive fork, so a
    DataLoader worker gets its own (threading)."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise LyftVoxelError(LV_E_NODEVICE, "no CUDA device visible (this package has no CPU fallback)")
        device = torch.cuda.current_device()
    key = (os.getpid(), int(device))
    h = _handles.get(key)
    if h is None:
        with _lock:
            h = _handles.get(key)
            if h is None:
                h = Handle(device)
                _handles[key] = h
    return h


def dptr(t):
    """Raw device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
