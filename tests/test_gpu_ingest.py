"""GPU parity: lv_ingest_sweeps (raw sweeps -> one cloud in the key frame)
vs the reference's own outputs (tests/golden/ref_ingest.npz) and the oracle - bit-exact."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ing():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import ingest
    return ingest


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


def test_second_ingest_equals_reference_output(ing, g):
    key, sweeps, ts_us = synth.ingest_case()
    got = ing.aggregate_sweeps_second(key, sweeps, ts=ts_us / 1e6)
    assert got.dtype == np.float32 and got.shape == g["second_points"].shape
    assert np.array_equal(got.view(np.uint32), g["second_points"].view(np.uint32))
    from oracle import ingest_oracle as io
    five = ing.aggregate_sweeps_second(key, sweeps, ts=ts_us / 1e6, keep_intensity=True)
    assert np.array_equal(five.view(np.uint32), io.second_aggregate_5col(key, sweeps, ts_us / 1e6).view(np.uint32))


def test_devkit_ingest_equals_reference_output(ing, g):
    key, sweeps, ts_us = synth.ingest_case()
    mats, lags = _mats_lags(sweeps, ts_us)
    pts, times = ing.aggregate_sweeps_devkit([key] + [s["points"] for s in sweeps], mats, lags, 1.0)
    assert np.array_equal(pts.view(np.uint32), g["devkit_points"].view(np.uint32))
    np.testing.assert_allclose(times, g["devkit_times"], rtol=1e-7)   # float32 lag vs float64 lag
    # static-shape form: removed rows are NaN rows and the voxelizer ignores them
    rows = ing.aggregate_sweeps_devkit([key] + [s["points"] for s in sweeps], mats, lags, 1.0, compact=False)
    assert int(np.isnan(rows[:, 0]).sum()) == rows.shape[0] - pts.shape[1]


def test_cuda_tensors_tma_and_plain_loads_feed_the_voxelizer(ing):
    """10-sweep cloud on the device: ingest (TMA-staged tiles and plain loads give identical
    bits) -> VoxelGenerator.generate, against oracle ingest -> oracle voxelizer."""
    import torch
    from lyft3d_b200 import _native as nat, voxel_generator as vg
    from oracle import ingest_oracle as io, voxel_oracle as vo
    key, sweeps, ts_us = synth.ingest_case(n_sweeps=10, n_key=20000, seed=3)
    ref_pts = io.second_aggregate(key, sweeps, ts=ts_us / 1e6)
    dkey = torch.from_numpy(key).cuda()
    dsweeps = [dict(sw, points=torch.from_numpy(sw["points"]).cuda()) for sw in sweeps]
    h = nat.get_handle(0)
    outs = []
    for off in (0, 1):
        h.set_option("disable_tma", off)
        try:
            outs.append(ing.aggregate_sweeps_second(dkey, dsweeps, ts=ts_us / 1e6))
        finally:
            h.set_option("disable_tma", 0)
    assert outs[0].is_cuda and bool((outs[0].view(torch.int32) == outs[1].view(torch.int32)).all())
    assert np.array_equal(outs[0].cpu().numpy().view(np.uint32), ref_pts.view(np.uint32))
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, max_voxels=30000)
    res = gen.generate(outs[0], 30000)
    v, c, n = vo.points_to_voxel(ref_pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    assert np.array_equal(res["coordinates"].cpu().numpy(), c)
    assert np.array_equal(res["num_points_per_voxel"].cpu().numpy(), n)
    assert np.array_equal(res["voxels"].cpu().numpy().view(np.uint32), v.view(np.uint32))


def test_empty_and_errors(ing):
    from lyft3d_b200 import _native as nat
    out = ing.ingest_sweeps([np.zeros((0, 5), np.float32)], [None], [0.0], ing.MODE_SECOND)
    assert out.shape == (0, 4)
    with pytest.raises(nat.LyftVoxelError):
        ing.ingest_sweeps([np.zeros((4, 5), np.float32)], [None], [0.0], 7)
