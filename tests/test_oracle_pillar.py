"""Pins oracle/pillar_oracle.py against outputs of the reference's own
pointpillars.py / voxel_encoder.py classes (tests/golden/ref_pillar_decorate.npz,
ref_scatter.npz - produced by oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import pillar_oracle as po

RTOL = 1e-6  # BASELINE.json north_star; second/second/framework/test.py:43
ATOL = 1e-5  # absolute slack on |x|<=50 m coordinates whose fp32 ulp is 3.8e-6


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_pillar_decorate.npz"))


@pytest.mark.parametrize("variant", po.VARIANTS)
@pytest.mark.parametrize("wd", [False, True])
def test_decorate_matches_reference(g, variant, wd):
    ref = g["dec_%s_%d" % (variant, int(wd))]
    out = po.decorate(g["voxels"], g["num_points"], g["coors"], synth.PILLAR_VOXEL_SIZE,
                      synth.PILLAR_RANGE, variant=variant, with_distance=wd)
    assert out.shape == ref.shape and out.dtype == np.float32
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)
    # padded slots are exactly zero
    mask = po.paddings_indicator(g["num_points"], ref.shape[1])
    assert np.all(out[~mask] == 0) and np.all(ref[~mask] == 0)


def test_pfn_and_simple_voxel(g):
    out = po.pfn_layer_eval(g["dec_pfn_0"], g["pfn_weight"], g["pfn_bn_gamma"], g["pfn_bn_beta"],
                            g["pfn_bn_mean"], g["pfn_bn_var"])
    np.testing.assert_allclose(out, g["pfn_out"], rtol=1e-5, atol=1e-5)
    sv = po.simple_voxel_mean(g["voxels"], g["num_points"])
    np.testing.assert_allclose(sv, g["simple_voxel"], rtol=RTOL, atol=ATOL)


def test_scatter_matches_reference(golden_dir):
    s = np.load(os.path.join(golden_dir, "ref_scatter.npz"))
    B, C, ny, nx = s["shape"]
    out = po.scatter(s["feats"], s["coords"], int(B), int(ny), int(nx))
    assert np.array_equal(out, s["canvas"])


def test_merge_batch_coords():
    a = np.array([[0, 1, 2], [0, 3, 4]], np.int32)
    b = np.array([[0, 5, 6]], np.int32)
    m = po.merge_batch_coords([a, b])
    assert m.tolist() == [[0, 0, 1, 2], [0, 0, 3, 4], [1, 0, 5, 6]] and m.dtype == np.int32


def test_eager_torch_restatement_equals_reference_outputs(golden_dir):
    """oracle/pillar_torch_ref.py (the reference's GPU op chain, timed by bench.py as `reference_gpu_eager`) gives
    the outputs of the reference's own classes bit for bit when run on the CPU like they were."""
    import os
    import torch
    from lyft3d_b200 import synth
    from oracle import pillar_torch_ref as tr
    g = np.load(os.path.join(golden_dir, "ref_pillar_decorate.npz"))
    vx, vy = synth.PILLAR_VOXEL_SIZE[0], synth.PILLAR_VOXEL_SIZE[1]
    xo, yo = vx / 2 + synth.PILLAR_RANGE[0], vy / 2 + synth.PILLAR_RANGE[1]
    for wd in (False, True):
        out = tr.decorate(torch.from_numpy(g["voxels"].copy()), torch.from_numpy(g["num_points"]), torch.from_numpy(g["coors"]),
                          vx, vy, xo, yo, with_distance=wd).numpy()
        assert np.array_equal(out.view(np.uint32), g["dec_pfn_%d" % int(wd)].view(np.uint32))
    s = np.load(os.path.join(golden_dir, "ref_scatter.npz"))
    B, C, ny, nx = (int(v) for v in s["shape"])
    canvas = tr.scatter(torch.from_numpy(s["feats"]), torch.from_numpy(s["coords"]), B, ny, nx).numpy()
    assert np.array_equal(canvas.view(np.uint32), s["canvas"].view(np.uint32))
