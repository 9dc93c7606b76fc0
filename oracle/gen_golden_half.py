"""Generates tests/golden/ref_pillar_half.npz by EXECUTING THE REFERENCE'S OWN CLASSES on float16 tensors
(second/second/pytorch/models/pointpillars.py: PillarFeatureNet / ...Old / ...Radius / ...RadiusHeight with the
PFN layers removed, and PointPillarsScatter) - the dtype they see under apex O2
(second/second/pytorch/train.py:34-47, configs/nuscenes/all.pp.mida.config:348).

Run:  python -m oracle.gen_golden_half          (needs /root/reference; CPU only)
TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import pillar_oracle, ref_loader, voxel_oracle  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")


def half_inputs(T=60, n_pillars=400, seed=21):
    """Real pillars of the bundled sweep at the PointPillars config (T = 60), a seeded selection that keeps
    the sparsest, the densest and a random middle, voxels rounded to float16 as train.py:34-47 does."""
    pts = synth.fixture_points_nx4()
    v, c, n = voxel_oracle.points_to_voxel(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, 30000)
    order = np.argsort(n, kind="stable")
    rng = np.random.default_rng(seed)
    pick = np.concatenate([order[:40], order[-80:], rng.choice(order[40:-80], n_pillars - 120, replace=False)])
    pick = rng.permutation(pick)
    coors4 = np.concatenate([np.zeros((len(pick), 1), np.int32), c[pick]], axis=1)
    coors4[len(pick) // 2:, 0] = 1
    return v[pick].astype(np.float16), n[pick], coors4


def main():
    import torch
    assert ref_loader.available(), "needs /root/reference"
    pp, ve = ref_loader.load_pointpillars()
    vh, num, coors = half_inputs()
    out = {"voxels": vh, "num_points": num, "coors": coors}
    kw = dict(num_input_features=4, use_norm=True, num_filters=(64,), voxel_size=synth.PILLAR_VOXEL_SIZE,
              pc_range=synth.PILLAR_RANGE)
    classes = {"pfn": pp.PillarFeatureNet, "old": pp.PillarFeatureNetOld, "radius": pp.PillarFeatureNetRadius,
               "radius_height": pp.PillarFeatureNetRadiusHeight}
    with torch.no_grad():
        for name, cls in classes.items():
            for wd in (False, True):
                net = cls(with_distance=wd, **kw)
                net.pfn_layers = torch.nn.ModuleList([])
                dec = net(torch.from_numpy(vh.copy()), torch.from_numpy(num), torch.from_numpy(coors)).numpy()
                assert dec.dtype == np.float16
                out["dec_%s_%d" % (name, int(wd))] = dec
                mine = pillar_oracle.decorate_half(vh, num, coors, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, name, wd)
                bad = int((mine.view(np.uint16) != dec.view(np.uint16)).sum())
                print("decorate half", name, wd, dec.shape, "restatement mismatches:", bad)
                assert bad == 0
        sc = pp.PointPillarsScatter(output_shape=[1, 1, 400, 400, 64], num_input_features=64)
        rng = np.random.default_rng(5)
        feats = rng.standard_normal((len(num), 64)).astype(np.float16)
        canvas = sc(torch.from_numpy(feats), torch.from_numpy(coors), 2).numpy()
        assert canvas.dtype == np.float16 and canvas.shape == (2, 64, 400, 400)
        nz = np.flatnonzero(canvas.reshape(-1).view(np.uint16))
        out["scatter_feats"] = feats
        out["scatter_nz_idx"] = nz.astype(np.int32)
        out["scatter_nz_val"] = canvas.reshape(-1)[nz]
        assert np.array_equal(pillar_oracle.scatter(feats, coors, 2, 400, 400).view(np.uint16), canvas.view(np.uint16))
    np.savez_compressed(os.path.join(GOLD, "ref_pillar_half.npz"), **out)
    print("wrote ref_pillar_half.npz")


if __name__ == "__main__":
    main()
