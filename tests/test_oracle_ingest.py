"""The ingest oracle (oracle/ingest_oracle.py) against outputs of the reference itself
(tests/golden/ref_ingest.npz, made by oracle/gen_golden_ingest.py executing
second/second/data/nuscenes_dataset.py:196-223 and the devkit's LidarPointCloud) - bit-exact."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import ingest_oracle as io, ref_loader


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_ingest.npz"))


def _mats_lags(sweeps, ts_us):
    mats, lags = [np.eye(4)], [0.0]
    for sw in sweeps:
        m = np.eye(4)
        m[:3, :3] = sw["sweep2lidar_rotation"]
        m[:3, 3] = sw["sweep2lidar_translation"]
        mats.append(m)
        lags.append(1e-6 * ts_us - 1e-6 * sw["timestamp"])
    return mats, lags


def test_second_aggregate_equals_reference_output(g):
    key, sweeps, ts_us = synth.ingest_case()
    got = io.second_aggregate(key, sweeps, ts=ts_us / 1e6)
    assert got.dtype == np.float32 and got.shape == g["second_points"].shape
    assert np.array_equal(got.view(np.uint32), g["second_points"].view(np.uint32))
    # two roundings (after @ R.T and after += T) differ from a single fused rounding somewhere
    fused = []
    for sw in sweeps:
        p = sw["points"][:, :3].astype(np.float64) @ sw["sweep2lidar_rotation"].T + sw["sweep2lidar_translation"]
        fused.append(p.astype(np.float32))
    n_key = key.shape[0]
    assert (np.concatenate(fused) != got[n_key:, :3]).any()


def test_devkit_aggregate_equals_reference_output(g):
    key, sweeps, ts_us = synth.ingest_case()
    mats, lags = _mats_lags(sweeps, ts_us)
    pts, times = io.devkit_aggregate([key] + [s["points"] for s in sweeps], mats, lags, 1.0)
    assert np.array_equal(pts.view(np.uint32), g["devkit_points"].view(np.uint32))
    assert np.array_equal(times, g["devkit_times"])
    assert pts.shape[1] < key.shape[0] + sum(s["points"].shape[0] for s in sweeps)   # remove_close removed rows


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_live_reference_statements(tmp_path):
    key, sweeps, ts_us = synth.ingest_case(n_sweeps=2, n_key=1500, seed=5)
    kp = tmp_path / "key.bin"
    key.tofile(str(kp))
    info = {"lidar_path": str(kp), "timestamp": ts_us, "sweeps": []}
    for i, sw in enumerate(sweeps):
        sp = tmp_path / ("s%d.bin" % i)
        sw["points"].tofile(str(sp))
        info["sweeps"].append({"lidar_path": str(sp), "timestamp": sw["timestamp"],
                               "sweep2lidar_rotation": sw["sweep2lidar_rotation"],
                               "sweep2lidar_translation": sw["sweep2lidar_translation"]})
    ref = ref_loader.run_second_get_sensor_points(info)
    assert np.array_equal(io.second_aggregate(key, sweeps, ts=ts_us / 1e6).view(np.uint32), ref.view(np.uint32))
