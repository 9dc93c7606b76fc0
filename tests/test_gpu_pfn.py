"""GPU parity: decoration + PFNLayer fused into one kernel vs the same
module running decoration kernel + the reference's PyTorch PFNLayer, and vs the oracle.

Tolerance: the Linear is an fp32 FMA chain over K = 9..11 in k order and BatchNorm is folded to
one multiply-add, torch uses a cuBLAS GEMM and (x - mean) * invstd * gamma + beta: results agree
to a few fp32 ulps of the pre-activation, i.e. rtol 1e-5 / atol 1e-5 on O(1..100) values."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-5

CLS = {"pfn": "PillarFeatureNet", "old": "PillarFeatureNetOld", "radius": "PillarFeatureNetRadius",
       "radius_height": "PillarFeatureNetRadiusHeight"}


@pytest.fixture(scope="module")
def pp():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import pointpillars
    return pointpillars


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_pillar_decorate.npz"))


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _randomize_bn(net, seed):
    import torch
    gen = torch.Generator().manual_seed(seed)
    bn = net.pfn_layers[0].norm
    with torch.no_grad():
        bn.weight.copy_(torch.rand(bn.weight.shape, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(bn.bias.shape, generator=gen) * 0.3)
        bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=gen) * 2)
        bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=gen) * 4 + 0.1)


@pytest.mark.parametrize("variant", list(CLS))
@pytest.mark.parametrize("wd", [False, True])
@pytest.mark.parametrize("units", [32, 64, 128])
def test_fused_pfn_equals_decorate_then_torch_layer(pp, g, variant, wd, units):
    import torch
    torch.manual_seed(11)
    net = pp.get_vfe_class(CLS[variant])(num_input_features=4, use_norm=True, num_filters=(units,), with_distance=wd,
                                         voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE).cuda().eval()
    _randomize_bn(net, 3)
    v, n, c = _cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])
    with torch.no_grad():
        assert net.can_fuse(v)
        fused = net(v, n, c)
        ref = net.pfn_layers[0](net.decorate(v, n, c)).squeeze()
    assert fused.shape == ref.shape == (v.shape[0], units)
    np.testing.assert_allclose(fused.cpu().numpy(), ref.cpu().numpy(), rtol=RTOL, atol=ATOL)
    # training mode / autograd keeps the reference's PyTorch layer
    net.train()
    assert not net.can_fuse(v)


def test_fused_pfn_without_norm_uses_the_bias(pp, g):
    import torch
    torch.manual_seed(5)
    net = pp.PillarFeatureNet(4, False, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().eval()
    v, n, c = _cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])
    with torch.no_grad():
        fused = net(v, n, c)
        ref = net.pfn_layers[0](net.decorate(v, n, c)).squeeze()
    np.testing.assert_allclose(fused.cpu().numpy(), ref.cpu().numpy(), rtol=RTOL, atol=ATOL)


def test_full_pillars_and_padding_max(pp):
    """A pillar with all T slots live must not see the padding value relu(shift); one with a
    single point must (pointpillars.py:62: the max runs over padded rows too)."""
    import torch
    from oracle import pillar_oracle as po
    rng = np.random.default_rng(8)
    P, T = 64, 60
    num = np.array([T] * 32 + [1] * 32, np.int32)
    v = (rng.normal(size=(P, T, 4)) * 3).astype(np.float32)
    v[np.arange(T)[None, :] >= num[:, None]] = 0
    coors = np.stack([np.zeros(P), np.zeros(P), rng.integers(0, 400, P), rng.integers(0, 400, P)], 1).astype(np.int32)
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().eval()
    _randomize_bn(net, 9)
    bn = net.pfn_layers[0].norm
    with torch.no_grad():
        bn.bias.fill_(5.0)          # relu(shift) is large: it must win for the 1-point pillars only
        bn.running_mean.zero_()
        out = net(_cuda(v), _cuda(num), _cuda(coors)).cpu().numpy()
    dec = po.decorate(v, num, coors, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE)
    ref = po.pfn_layer_eval(dec, net.pfn_layers[0].linear.weight.detach().cpu().numpy(), bn.weight.detach().cpu().numpy(),
                            bn.bias.detach().cpu().numpy(), bn.running_mean.cpu().numpy(), bn.running_var.cpu().numpy())
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
    assert np.all(out[32:] >= 5.0 - 1e-4)


def test_engine_points_to_pillar_features(pp):
    """lv_pillarize_pfn_concat (points -> features, one call) == pillarize -> torch PFNLayer."""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine
    F = 6
    frames = [synth.c5_frame(200 + f) for f in range(F)]
    n = frames[0].shape[0]
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    eng = FrameBatchEngine(0, F, n)
    torch.manual_seed(2)
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().eval()
    _randomize_bn(net, 21)
    eng.pillarize(pts)
    rows = eng.read_total_rows()
    with torch.no_grad():
        ref = net.pfn_layers[0](eng.decorated[:rows]).squeeze().clone()
    coords, num = eng.coords[:rows].clone(), eng.num_points[:rows].clone()
    w, scale, shift = pp.fold_pfn_layer(net.pfn_layers[0])
    eng.features.zero_()
    eng.pillar_features(pts, w, scale, shift)
    assert eng.read_total_rows() == rows
    np.testing.assert_allclose(eng.features[:rows].cpu().numpy(), ref.cpu().numpy(), rtol=RTOL, atol=ATOL)
    assert bool((eng.coords[:rows] == coords).all()) and bool((eng.num_points[:rows] == num).all())
