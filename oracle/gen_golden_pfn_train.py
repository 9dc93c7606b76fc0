"""Generates tests/golden/ref_pfn_train.npz by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden_pfn_train        (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  The reference's own PillarFeatureNet* classes
(second/second/pytorch/models/pointpillars.py, loaded by oracle.ref_loader) in TRAINING mode -
decoration (:203-231) -> PFNLayer with BatchNorm1d batch statistics (:51-65) - on the CPU in
float32 and, as the ground truth the tolerance is judged against, in float64: forward output,
the gradients of linear.weight / norm.weight / norm.bias for a seeded upstream gradient (loss =
sum(out * G), G multiples of 1/64), and the running statistics after the step.  Inputs: 300 pillars
(T = 20) cut from a seeded C5 frame by the oracle voxelizer: batch statistics over 6,000 slots.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import pillar_oracle, ref_loader, voxel_oracle  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")
T, P = 20, 300
CASES = [("pfn", False, 64), ("old", False, 64), ("radius", False, 32), ("radius_height", True, 64), ("pfn", True, 128)]
CLS = {"pfn": "PillarFeatureNet", "old": "PillarFeatureNetOld", "radius": "PillarFeatureNetRadius",
       "radius_height": "PillarFeatureNetRadiusHeight"}


def inputs():
    pts = synth.c5_frame(7)[:30000]
    v, c, n = voxel_oracle.points_to_voxel(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, 30000)
    # the fullest pillars first would bias the statistics: take an evenly spaced subset
    idx = np.linspace(0, v.shape[0] - 1, P).astype(np.int64)
    return v[idx].copy(), n[idx].copy(), pillar_oracle.merge_batch_coords([c[idx]])


def main():
    assert ref_loader.available(), "needs /root/reference"
    import torch
    pp, _ = ref_loader.load_pointpillars()
    voxels, num, coors = inputs()
    out = {"voxels": voxels, "num_points": num, "coors": coors}
    for variant, wd, units in CASES:
        tag = "%s_%d_%d" % (variant, int(wd), units)
        torch.manual_seed(100 + units)
        G = torch.randint(-128, 129, (P, units)).float() / 64
        res = {}
        for dt in (torch.float32, torch.float64):
            torch.manual_seed(7)
            net = getattr(pp, CLS[variant])(num_input_features=4, use_norm=True, num_filters=(units,), with_distance=wd,
                                            voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE)
            bn = net.pfn_layers[0].norm
            with torch.no_grad():
                bn.weight.copy_(torch.rand(units) + 0.5)
                bn.bias.copy_(torch.randn(units) * 0.3)
            net = net.to(dt).train()
            if dt == torch.float32:
                out[tag + ".weight"] = net.pfn_layers[0].linear.weight.detach().numpy().copy()
                out[tag + ".gamma"] = bn.weight.detach().numpy().copy()
                out[tag + ".beta"] = bn.bias.detach().numpy().copy()
            y = net(torch.from_numpy(voxels.copy()).to(dt), torch.from_numpy(num), torch.from_numpy(coors))
            (y * G.to(dt)).sum().backward()
            res[dt] = dict(out=y.detach().numpy(), dW=net.pfn_layers[0].linear.weight.grad.numpy(),
                           dgamma=bn.weight.grad.numpy(), dbeta=bn.bias.grad.numpy(),
                           running_mean=bn.running_mean.numpy().copy(), running_var=bn.running_var.numpy().copy())
        out[tag + ".G"] = (G * 64).numpy().astype(np.int16)    # G = this / 64
        for k, a in res[torch.float32].items():
            out["%s.f32.%s" % (tag, k)] = a.astype(np.float32)
        for k, a in res[torch.float64].items():
            out["%s.f64.%s" % (tag, k)] = a.astype(np.float32 if k == "out" else np.float64)   # (out: rounded once)
            scale = np.abs(a).max()
            print(tag, k, a.shape, "float32 run vs float64 run: max abs diff %.3e (max |value| %.3e)" %
                  (np.abs(a - res[torch.float32][k]).max(), scale))
    path = os.path.join(GOLD, "ref_pfn_train.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
