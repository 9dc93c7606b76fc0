"""Generates tests/golden/ref_ingest.npz by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden_ingest        (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  What is produced and from what:
  second_points   the reference's own statements second/second/data/nuscenes_dataset.py:196-223
                  (NuScenesDataset.get_sensor_data's point assembly), executed unchanged by
                  oracle.ref_loader.run_second_get_sensor_points on temporary .bin files holding
                  lyft3d_b200.synth.ingest_case().
  devkit_points / devkit_times
                  the reference's LidarPointCloud.from_file -> transform -> remove_close chain
                  (nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py:269-284, 188-195,
                  153-165) composed as from_file_multisweep does (:99-137) on the same sweeps
                  with the per-sweep 4x4 built from the same rotation / translation.
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")


def main():
    assert ref_loader.available(), "needs /root/reference"
    key, sweeps, ts_us = synth.ingest_case()
    with tempfile.TemporaryDirectory() as td:
        kp = os.path.join(td, "key.bin")
        key.tofile(kp)
        info = {"lidar_path": kp, "timestamp": ts_us, "token": "t", "sweeps": []}
        for i, sw in enumerate(sweeps):
            sp = os.path.join(td, "sweep%d.bin" % i)
            sw["points"].tofile(sp)
            info["sweeps"].append({"lidar_path": sp, "timestamp": sw["timestamp"],
                                   "sweep2lidar_rotation": sw["sweep2lidar_rotation"],
                                   "sweep2lidar_translation": sw["sweep2lidar_translation"]})
        second_points = ref_loader.run_second_get_sensor_points(info)

        dc = ref_loader.load_devkit_data_classes()
        all_pts = np.zeros((4, 0), dtype=np.float32)
        all_times = np.zeros((1, 0))
        paths = [kp] + [s["lidar_path"] for s in info["sweeps"]]
        mats = [np.eye(4)]
        lags = [0.0]
        for sw in sweeps:
            m = np.eye(4)
            m[:3, :3] = sw["sweep2lidar_rotation"]
            m[:3, 3] = sw["sweep2lidar_translation"]
            mats.append(m)
            lags.append(1e-6 * ts_us - 1e-6 * sw["timestamp"])
        for pth, m, lag in zip(paths, mats, lags):
            pc = dc.LidarPointCloud.from_file(Path(pth))
            pc.points = np.array(pc.points)      # from_file returns a view of the read buffer
            pc.transform(m)
            pc.remove_close(1.0)
            all_times = np.hstack((all_times, lag * np.ones((1, pc.nbr_points()))))
            all_pts = np.hstack((all_pts, pc.points))
    np.savez_compressed(os.path.join(GOLD, "ref_ingest.npz"), second_points=second_points,
                        devkit_points=all_pts.astype(np.float32), devkit_times=all_times)
    print("second_points", second_points.shape, second_points.dtype, "devkit", all_pts.shape, all_times.shape)


if __name__ == "__main__":
    main()
