"""float16 decoration and scatter (apex O2, second/second/pytorch/train.py:34-47): tests/golden/ref_pillar_half.npz
holds the outputs of the reference's OWN classes run on half tensors (oracle/gen_golden_half.py).  CPU: the
restated rounding model equals them bit for bit.  GPU: the kernels equal them - the scatter and every channel that
does not depend on the pillar mean bit for bit, the mean-dependent channels within one half ulp (the float32
accumulation order of the sum over T is the only freedom)."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import pillar_oracle as po

VARIANTS = {"pfn": "PillarFeatureNet", "old": "PillarFeatureNetOld", "radius": "PillarFeatureNetRadius",
            "radius_height": "PillarFeatureNetRadiusHeight"}


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_pillar_half.npz"))


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("wd", [False, True])
def test_rounding_model_equals_reference_half_outputs(g, variant, wd):
    mine = po.decorate_half(g["voxels"], g["num_points"], g["coors"], synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, variant, wd)
    ref = g["dec_%s_%d" % (variant, int(wd))]
    assert mine.dtype == np.float16 and np.array_equal(mine.view(np.uint16), ref.view(np.uint16))


def _cluster_channels(variant):
    k = 4 if variant in ("pfn", "old") else 3
    return [k, k + 1, k + 2]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("wd", [False, True])
def test_gpu_half_decoration_vs_reference(g, variant, wd):
    import torch
    from lyft3d_b200 import pointpillars as pp
    cls = pp.get_vfe_class(VARIANTS[variant])
    net = cls(num_input_features=4, use_norm=True, num_filters=(64,), with_distance=wd,
              voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE).cuda()
    v = torch.from_numpy(g["voxels"]).cuda()
    out = net.decorate(v, torch.from_numpy(g["num_points"]).cuda(), torch.from_numpy(g["coors"]).cuda())
    assert out.dtype == torch.float16
    out = out.cpu().numpy()
    ref = g["dec_%s_%d" % (variant, int(wd))]
    assert out.shape == ref.shape
    cl = _cluster_channels(variant)
    rest = [c for c in range(ref.shape[2]) if c not in cl]
    # values, not bits: the reference leaves -0.0 in padded slots (negative feature times mask 0); -0.0 == 0.0
    assert np.array_equal(out[..., rest], ref[..., rest])
    a, b = out[..., cl].astype(np.float32), ref[..., cl].astype(np.float32)
    ulp = np.spacing(np.maximum(np.abs(b), np.float32(2.0 ** -14)).astype(np.float16)).astype(np.float32)
    worst = float((np.abs(a - b) / ulp).max())
    frac_exact = float((out[..., cl] == ref[..., cl]).mean())
    print("half %s wd=%d: mean-dependent channels worst %.2f half-ulp, %.4f bit-identical" % (variant, wd, worst, frac_exact))
    assert worst <= 1.0 and frac_exact > 0.99


@pytest.mark.gpu
def test_gpu_half_forward_runs_the_pfn_in_half(g):
    """The drop-in VFE under apex O2: half voxels in, half features out (the PFNLayer stays torch's)."""
    import torch
    from lyft3d_b200 import pointpillars as pp
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().half().eval()
    v = torch.from_numpy(g["voxels"]).cuda()
    n, c = torch.from_numpy(g["num_points"]).cuda(), torch.from_numpy(g["coors"]).cuda()
    with torch.no_grad():
        out = net(v, n, c)
        ref = net.pfn_layers[0](torch.from_numpy(g["dec_pfn_0"]).cuda()).squeeze()
    assert out.dtype == torch.float16 and out.shape == (v.shape[0], 64)
    assert float((out.float() - ref.float()).abs().max()) <= 2e-2


@pytest.mark.gpu
def test_gpu_half_scatter_bit_exact(g):
    import torch
    from lyft3d_b200 import pointpillars as pp
    feats = torch.from_numpy(g["scatter_feats"]).cuda()
    coors = torch.from_numpy(g["coors"]).cuda()
    sc = pp.PointPillarsScatter(output_shape=[1, 1, 400, 400, 64], num_input_features=64)
    canvas = sc(feats, coors, 2)
    assert canvas.dtype == torch.float16 and canvas.shape == (2, 64, 400, 400)
    flat = canvas.cpu().numpy().reshape(-1)
    ref = np.zeros(flat.shape, np.float16)
    ref[g["scatter_nz_idx"]] = g["scatter_nz_val"]
    assert np.array_equal(flat.view(np.uint16), ref.view(np.uint16))
    again = sc(feats, coors, 2)                                           # the cell -> pillar map was left empty
    assert bool((again.view(torch.int16) == canvas.view(torch.int16)).all())
    # equals the float32 scatter of the same features, cast
    c32 = sc(feats.float(), coors, 2)
    assert bool((c32.half().view(torch.int16) == canvas.view(torch.int16)).all())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 64, 400, 400), (3, 128, 16, 48), (2, 64, 10, 10), (2, 9, 33, 7), (1, 64, 496, 432)])
def test_gpu_half_scatter_shapes_vs_oracle(shape):
    import torch
    from lyft3d_b200 import pointpillars as pp
    B, C, ny, nx = shape
    rng = np.random.default_rng(sum(shape))
    P = min(B * ny * nx // 3 + 1, 20000)
    cells = rng.choice(B * ny * nx, P, replace=False)
    coors = np.stack([cells // (ny * nx), np.zeros(P, np.int64), (cells % (ny * nx)) // nx, cells % nx], axis=1).astype(np.int32)
    feats = rng.standard_normal((P, C)).astype(np.float16)
    got = pp.scatter_pillars(torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda(), B, ny, nx).cpu().numpy()
    assert np.array_equal(got.view(np.uint16), po.scatter(feats, coors, B, ny, nx).view(np.uint16))
    # gradient: the backward is the gather of the canvas gradient
    f = torch.from_numpy(feats).cuda().requires_grad_(True)
    canvas = pp.scatter_pillars(f, torch.from_numpy(coors).cuda(), B, ny, nx)
    w = torch.randn_like(canvas)
    (canvas * w).sum().backward()
    co = torch.from_numpy(coors).cuda().long()
    assert bool((f.grad == w[co[:, 0], :, co[:, 2], co[:, 3]]).all())
