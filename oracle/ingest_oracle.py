"""TEST INFRASTRUCTURE - CPU restatement of the reference's multi-sweep ingest (numpy).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the
product (lyft-3d-object-detection_b200/) never does.

Restates, with in-memory arrays in place of the file reads,
  second/second/data/nuscenes_dataset.py:196-223   (get_sensor_data: SECOND's 10-sweep cloud)
  nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py:99-137, 153-165, 188-195
                                                    (from_file_multisweep, remove_close, transform)
Pinned by tests/test_oracle_ingest.py against the reference's own PointCloud class executed
through oracle/ref_loader.py when /root/reference is present (transform / remove_close), and
by the frozen hashes in tests/golden/ingest_hashes.json.
"""
import numpy as np


def second_aggregate(key_raw, sweeps, ts=0.0):
    """nuscenes_dataset.py:196-223.

    key_raw: (N,5) float32 rows of the key sweep (as np.fromfile(...).reshape([-1, 5])).
    sweeps:  list of dicts with "points" (M,5) float32, "sweep2lidar_rotation" (3,3),
             "sweep2lidar_translation" (3,), "timestamp" (microseconds); `ts` is the key
             sweep's timestamp in SECONDS (info["timestamp"] / 1e6).
    Returns (sum N, 4) float32 [x, y, z, time lag]."""
    points = np.array(key_raw, dtype=np.float32, copy=True).reshape([-1, 5])
    points[:, 3] /= 255                                   # :203
    points[:, 4] = 0                                      # :204
    sweep_points_list = [points]
    for sweep in sweeps:                                  # :208
        points_sweep = np.array(sweep["points"], dtype=np.float32, copy=True).reshape([-1, 5])
        sweep_ts = sweep["timestamp"] / 1e6               # :214
        points_sweep[:, 3] /= 255                         # :215
        points_sweep[:, :3] = points_sweep[:, :3] @ sweep["sweep2lidar_rotation"].T   # :216-217
        points_sweep[:, :3] += sweep["sweep2lidar_translation"]                       # :218
        points_sweep[:, 4] = ts - sweep_ts                # :219
        sweep_points_list.append(points_sweep)
    return np.concatenate(sweep_points_list, axis=0)[:, [0, 1, 2, 4]]                  # :223


def second_aggregate_5col(key_raw, sweeps, ts=0.0):
    """Same, without the final column select: [x, y, z, intensity/255, time lag]."""
    points = np.array(key_raw, dtype=np.float32, copy=True).reshape([-1, 5])
    points[:, 3] /= 255
    points[:, 4] = 0
    out = [points]
    for sweep in sweeps:
        ps = np.array(sweep["points"], dtype=np.float32, copy=True).reshape([-1, 5])
        ps[:, 3] /= 255
        ps[:, :3] = ps[:, :3] @ sweep["sweep2lidar_rotation"].T
        ps[:, :3] += sweep["sweep2lidar_translation"]
        ps[:, 4] = ts - sweep["timestamp"] / 1e6
        out.append(ps)
    return np.concatenate(out, axis=0)


def devkit_transform(points_4xn, transf_matrix):
    """PointCloud.transform, data_classes.py:188-195 (float64 product, float32 store)."""
    pts = np.array(points_4xn, dtype=np.float32, copy=True)
    pts[:3, :] = transf_matrix.dot(np.vstack((pts[:3, :], np.ones(pts.shape[1]))))[:3, :]
    return pts


def devkit_remove_close_mask(points_4xn, radius):
    """PointCloud.remove_close, data_classes.py:153-165 -> boolean keep mask."""
    x_filt = np.abs(points_4xn[0, :]) < radius
    y_filt = np.abs(points_4xn[1, :]) < radius
    return np.logical_not(np.logical_and(x_filt, y_filt))


def devkit_aggregate(sweeps_raw, matrices, time_lags, min_distance=1.0):
    """from_file_multisweep, data_classes.py:99-137, given the already fused 4x4 of each
    sweep (:118) and its time lag (:125).  Returns (points (4,N) float32, times (1,N) float64)
    with close points REMOVED, as the reference does."""
    all_pts = np.zeros((4, 0), dtype=np.float32)
    all_times = np.zeros((1, 0))
    for raw, m, lag in zip(sweeps_raw, matrices, time_lags):
        pc = np.ascontiguousarray(np.asarray(raw, dtype=np.float32).reshape(-1, 5)[:, :4].T)   # from_file :282-284
        pc = devkit_transform(pc, np.asarray(m, dtype=np.float64))
        pc = pc[:, devkit_remove_close_mask(pc, min_distance)]
        times = lag * np.ones((1, pc.shape[1]))
        all_times = np.hstack((all_times, times))
        all_pts = np.hstack((all_pts, pc))
    return all_pts, all_times
