"""GPU parity of the fused TRAINING form of the pillar feature net: decoration
(pointpillars.py:203-231) + PFNLayer with BatchNorm1d batch statistics (:51-65) + autograd backward as one op
on the CUDA library (lv_pillar_pfn_moments, lv_pillar_pfn, lv_pillar_pfn_backward), against

  * tests/golden/ref_pfn_train.npz: the reference's OWN classes run in training mode on the CPU in float32 and in
    float64 (oracle/gen_golden_pfn_train.py), forward output, dL/dW, dL/dgamma, dL/dbeta, running statistics;
  * the same module with the fused path switched off (decoration kernel + the PyTorch PFNLayer + autograd) on the
    GPU, at a C5-frame size (8.6k pillars, T = 60).

Tolerance: 1e-5 relative to the largest magnitude of each tensor, judged against the float64 run (the reference's
own float32 run differs from its float64 run by up to 3e-7 on the same scale; both errors are printed)."""
import copy
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import gen_golden_pfn_train as gg

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def pp():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import pointpillars
    return pointpillars


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_pfn_train.npz"))


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _relerr(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(np.asarray(got, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("variant,wd,units", gg.CASES)
def test_fused_training_equals_reference_classes(pp, g, variant, wd, units):
    import torch
    tag = "%s_%d_%d" % (variant, int(wd), units)
    net = pp.get_vfe_class(gg.CLS[variant])(num_input_features=4, use_norm=True, num_filters=(units,), with_distance=wd,
                                            voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE).cuda().train()
    layer = net.pfn_layers[0]
    with torch.no_grad():
        layer.linear.weight.copy_(_cuda(g[tag + ".weight"]))
        layer.norm.weight.copy_(_cuda(g[tag + ".gamma"]))
        layer.norm.bias.copy_(_cuda(g[tag + ".beta"]))
    v, n, c = _cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])
    assert net.can_fuse_training(v)
    lib_launches = pp.nat.get_handle(0).launches()
    out = net(v, n, c)
    G = _cuda(g[tag + ".G"].astype(np.float32) / 64)
    (out * G).sum().backward()
    assert pp.nat.get_handle(0).launches() - lib_launches == 5, "moments, statistics, forward, backward, finish kernels expected"
    got = dict(out=out.detach().cpu().numpy(), dW=layer.linear.weight.grad.cpu().numpy(),
               dgamma=layer.norm.weight.grad.cpu().numpy(), dbeta=layer.norm.bias.grad.cpu().numpy(),
               running_mean=layer.norm.running_mean.cpu().numpy(), running_var=layer.norm.running_var.cpu().numpy())
    assert int(layer.norm.num_batches_tracked) == 1
    for k, a in got.items():
        e64, e32 = _relerr(a, g["%s.f64.%s" % (tag, k)]), _relerr(g["%s.f32.%s" % (tag, k)], g["%s.f64.%s" % (tag, k)])
        print("%s %-12s rel err vs the reference in float64: %.2e   (the reference's own float32 run: %.2e)" % (tag, k, e64, e32))
        assert e64 <= REL, (k, e64)


def test_fused_training_equals_torch_layer_at_frame_size(pp):
    """One C5 frame (8.6k pillars, T = 60): fused op vs decoration kernel + PyTorch PFNLayer + autograd on the GPU;
    two optimizer-free steps so that the running statistics are updated twice."""
    import torch
    from oracle import pillar_oracle, voxel_oracle
    v, c, n = voxel_oracle.points_to_voxel(synth.c5_frame(3), synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    co = pillar_oracle.merge_batch_coords([c])
    tv, tn, tc = _cuda(v), _cuda(n), _cuda(co)
    torch.manual_seed(5)
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().train()
    with torch.no_grad():
        net.pfn_layers[0].norm.weight.copy_(torch.rand(64, device="cuda") + 0.5)
        net.pfn_layers[0].norm.bias.copy_(torch.randn(64, device="cuda") * 0.3)
    ref = copy.deepcopy(net).double()
    ref.fuse_training = False
    ref32 = copy.deepcopy(net)          # the PyTorch layer in float32: the error the reference itself has at this size
    ref32.fuse_training = False
    G = torch.randn(v.shape[0], 64, device="cuda")

    def tensors(m, out):
        layer = m.pfn_layers[0]
        return dict(out=out, dW=layer.linear.weight.grad, dgamma=layer.norm.weight.grad, dbeta=layer.norm.bias.grad,
                    running_mean=layer.norm.running_mean, running_var=layer.norm.running_var)
    for step in range(2):
        for m in (net, ref, ref32):
            m.zero_grad()
        out = net(tv, tn, tc)
        (out * G).sum().backward()
        # float64 ground truth: the decoration kernel is float32 -> run it once, then the PyTorch layer in float64
        dec = net.decorate(tv, tn, tc)
        out_ref = ref.pfn_layers[0](dec.double()).squeeze()
        (out_ref * G.double()).sum().backward()
        out32 = ref32.pfn_layers[0](dec).squeeze()
        (out32 * G).sum().backward()
        t64, t32, mine = tensors(ref, out_ref), tensors(ref32, out32), tensors(net, out)
        for name in t64:
            truth = t64[name].detach().cpu().numpy()
            e = _relerr(mine[name].detach().cpu().numpy(), truth)
            e32 = _relerr(t32[name].detach().cpu().numpy(), truth)
            print("step %d %-12s rel err vs float64 PyTorch: %.2e   (PyTorch float32: %.2e)" % (step, name, e, e32))
            # a float32 forward may pick another slot than float64 where two activations nearly tie: the bound is the
            # contract's 1e-5, or twice what the PyTorch float32 layer itself deviates by
            assert e <= max(REL, 2 * e32), (step, name, e, e32)
    assert int(net.pfn_layers[0].norm.num_batches_tracked) == 2


def test_training_falls_back_when_the_fused_form_does_not_apply(pp):
    """Two PFN layers, no norm, or voxels that need a gradient: the PyTorch layers run (and still train)."""
    import torch
    v = torch.rand(40, 20, 4, device="cuda")
    n = torch.randint(1, 21, (40,), device="cuda", dtype=torch.int32)
    c = torch.zeros(40, 4, device="cuda", dtype=torch.int32)
    c[:, 2] = torch.arange(40)
    two = pp.PillarFeatureNet(4, True, (32, 64), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().train()
    nonorm = pp.PillarFeatureNet(4, False, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().train()
    for net in (two, nonorm):
        assert not net.can_fuse_training(v)
        out = net(v, n, c)
        out.sum().backward()
        assert out.shape == (40, 64) and net.pfn_layers[0].linear.weight.grad is not None
    one = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().train()
    assert one.can_fuse_training(v) and not one.can_fuse_training(v.clone().requires_grad_(True))
    assert not one.eval().can_fuse_training(v)
