"""Multi-sweep ingest drop-ins: raw 5-float sweeps -> one cloud in the key
frame on the GPU, in front of VoxelGenerator.generate / create_voxel_pointcloud.

    aggregate_sweeps_second(key_raw, sweeps, ts)   second/second/data/nuscenes_dataset.py:196-223
    aggregate_sweeps_devkit(sweeps_raw, matrices, time_lags, min_distance)
                                                   lyft_dataset_sdk/utils/data_classes.py:99-137

Inputs are the arrays the reference reads from disk (np.fromfile(...).reshape([-1, 5])) plus the
per-sweep calibration its info dicts hold; numpy in -> numpy out, CUDA tensors in -> CUDA tensor
out on the current stream.  All arithmetic runs in lv_ingest_sweeps (sm_100a); no CPU path.
"""
import numpy as np

from . import _native as nat

MODE_SECOND = 0
MODE_DEVKIT = 1


def _is_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def ingest_sweeps(raw_list, matrices, time_lags, mode, close_radius=-1.0, out_cols=4, has_tm=None, handle=None):
    """raw_list: per-sweep (M,5) float32 arrays (numpy or CUDA tensors; all of one kind).
    matrices: per-sweep 4x4 float64 (None entries / has_tm 0 = untouched sweep)."""
    import torch
    lib = nat.load()
    S = len(raw_list)
    cuda = S > 0 and _is_cuda(raw_list[0])
    sizes = [int(r.shape[0]) for r in raw_list]
    offs = np.zeros(S + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    n = int(offs[-1])
    if cuda:
        dev = raw_list[0].device
        raw = torch.cat([r.reshape(-1, 5) for r in raw_list], dim=0).contiguous() if S > 1 else raw_list[0].reshape(-1, 5).contiguous()
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if dev is None:
            raise nat.LyftVoxelError(nat.LV_E_NODEVICE, "no CUDA device visible (this package has no CPU fallback)")
        host = np.concatenate([np.asarray(r, dtype=np.float32).reshape(-1, 5) for r in raw_list], axis=0) if S else \
            np.zeros((0, 5), np.float32)
        raw = torch.from_numpy(np.ascontiguousarray(host)).to(dev)
    if raw.dtype != torch.float32:
        raise ValueError("sweeps must be float32")
    tm = np.zeros((S, 16), np.float64)
    flags = np.zeros(S, np.uint8)
    for s in range(S):
        m = None if matrices is None else matrices[s]
        if m is not None and (has_tm is None or has_tm[s]):
            tm[s] = np.asarray(m, dtype=np.float64).reshape(16)
            flags[s] = 1
    lag = np.ascontiguousarray(np.asarray(time_lags, dtype=np.float64).astype(np.float32))
    out = torch.empty((n, out_cols), dtype=torch.float32, device=dev)
    h = handle or nat.get_handle(dev.index)
    with torch.cuda.device(dev):
        nat.check(lib.lv_ingest_sweeps(h.ptr, raw.data_ptr(), S, offs.ctypes.data, tm.ctypes.data, flags.ctypes.data,
                                       lag.ctypes.data, int(mode), float(close_radius), int(out_cols), out.data_ptr(),
                                       nat.current_stream_ptr(dev)))
    return out if cuda else out.cpu().numpy()


def aggregate_sweeps_second(key_raw, sweeps, ts=0.0, keep_intensity=False):
    """get_sensor_data's point assembly (nuscenes_dataset.py:196-223).  `sweeps` are the info
    dicts' entries with "points" holding the (M,5) rows read from sweep["lidar_path"].
    Returns (sum N, 4) [x, y, z, time lag] (or 5 columns with intensity/255)."""
    raws, mats, lags = [key_raw], [None], [0.0]
    for sw in sweeps:
        m = np.eye(4)
        m[:3, :3] = np.asarray(sw["sweep2lidar_rotation"], dtype=np.float64)
        m[:3, 3] = np.asarray(sw["sweep2lidar_translation"], dtype=np.float64)
        raws.append(sw["points"])
        mats.append(m)
        lags.append(ts - sw["timestamp"] / 1e6)
    return ingest_sweeps(raws, mats, lags, MODE_SECOND, out_cols=5 if keep_intensity else 4)


def aggregate_sweeps_devkit(sweeps_raw, matrices, time_lags, min_distance=1.0, compact=True):
    """LidarPointCloud.from_file_multisweep (data_classes.py:99-137) given each sweep's fused
    4x4 (:118) and time lag (:125).  Returns (points (4,N) float32, times (1,N)) like the
    reference; with compact=False the removed rows stay in place as NaN rows (static shapes,
    no host synchronisation) and the result is the (N_total, 5) row matrix."""
    rows = ingest_sweeps(list(sweeps_raw), list(matrices), time_lags, MODE_DEVKIT, close_radius=float(min_distance),
                         out_cols=5)
    if not compact:
        return rows
    if _is_cuda(rows):
        keep = ~rows[:, 0].isnan()
        rows = rows[keep]
        return rows[:, :4].t().contiguous(), rows[:, 4:5].t().double().contiguous()
    keep = ~np.isnan(rows[:, 0])
    rows = rows[keep]
    return np.ascontiguousarray(rows[:, :4].T), np.ascontiguousarray(rows[:, 4:5].T.astype(np.float64))
