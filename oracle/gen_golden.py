"""Generates tests/golden/* by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden          (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  The outputs are small fixtures that travel to the GPU box
(where /root/reference does not exist).  What is produced and from what:

  ref_simplevis_pillar.npz   reference second/second/utils/simplevis.py:64-108
                             ``points_to_bev`` run on the bundled sweep at the
                             pillar config: the per-pillar point-count map
                             (bev_map[-1]) with max_voxels=40000 and, to pin the
                             ``break`` rule (:46-50), with max_voxels=3000.
  ref_pillar_decorate.npz    reference pointpillars.py classes PillarFeatureNet,
                             PillarFeatureNetOld, ...Radius, ...RadiusHeight with the
                             PFN layers removed -> pure decoration outputs, with and
                             without ``with_distance``; plus the full PillarFeatureNet
                             (Linear+BN+ReLU+max, eval mode, torch.manual_seed(0)) and
                             voxel_encoder.py SimpleVoxel / SimpleVoxelRadius.
  ref_scatter.npz            reference PointPillarsScatter.forward (pointpillars.py:444-476).
  c1_bev_known.json          BEV known answers on the bundled sweep (restated numpy
                             oracle; the reference BEV closures cannot be imported,
                             SURVEY.md F5) - must equal the survey-time hashes.
  voxel_oracle_hashes.json   sha256 of the C oracle's outputs at configs C2/C3
                             (PARITY UNPINNED at the spconv boundary, see voxel_oracle.py).
"""
import hashlib
import json
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import bev_oracle, pillar_oracle, ref_loader, voxel_oracle  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def gen_c1():
    pts = synth.fixture_points_4xn()
    bev = bev_oracle.create_voxel_pointcloud(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    u8 = bev_oracle.quantize_u8(bev_oracle.normalize_voxel_intensities(bev))
    known = {
        "n_points": int(pts.shape[1]),
        "in_bounds": int(bev.sum()),
        "nonzero_cells": int((bev > 0).sum()),
        "max_count": int(bev.max()),
        "saturated_cells": int((bev >= 16).sum()),
        "channel_sums": [int(v) for v in bev.sum(axis=(0, 1))],
        "sha256_16_raw_f32": sha16(bev),
        "sha256_16_u8": sha16(u8),
        "sha256_16_norm_f32": sha16(bev_oracle.normalize_voxel_intensities(bev)),
    }
    # survey-time hashes (SURVEY.md 8c / BASELINE.md 4)
    assert known["sha256_16_raw_f32"] == "d8aa630368259baa", known
    assert known["sha256_16_u8"] == "06fc320850db7693", known
    with open(os.path.join(GOLD, "c1_bev_known.json"), "w") as f:
        json.dump(known, f, indent=1)
    print("c1", known)


def gen_simplevis():
    sv = ref_loader.load_simplevis()
    pts = synth.fixture_points_nx4()
    full = sv.points_to_bev(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, max_voxels=40000)
    brk = sv.points_to_bev(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, max_voxels=3000)
    assert full[-1].max() < 65535
    np.savez_compressed(os.path.join(GOLD, "ref_simplevis_pillar.npz"),
                        density_full=full[-1].astype(np.uint16),
                        density_break3000=brk[-1].astype(np.uint16),
                        height_full=full[0].astype(np.float32))
    print("simplevis", full[-1].sum(), (full[-1] > 0).sum(), brk[-1].sum(), (brk[-1] > 0).sum())


def _pillar_inputs():
    """32 real pillars of the bundled sweep (T=20) incl. num=1 and num=T cases."""
    pts = synth.fixture_points_nx4()
    T = 20
    v, c, n = voxel_oracle.points_to_voxel(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, 30000)
    order = np.argsort(n, kind="stable")
    pick = np.concatenate([order[:6], order[-10:], order[len(order) // 2 - 8:len(order) // 2 + 8]])
    rng = np.random.default_rng(7)
    pick = rng.permutation(pick)
    voxels, coors, num = v[pick], c[pick], n[pick]
    coors4 = np.concatenate([np.zeros((len(pick), 1), np.int32), coors], axis=1)
    coors4[len(pick) // 2:, 0] = 1
    return voxels, num, coors4


def gen_pillar():
    import torch
    pp, ve = ref_loader.load_pointpillars()
    voxels, num, coors = _pillar_inputs()
    out = {"voxels": voxels, "num_points": num, "coors": coors}
    kw = dict(num_input_features=4, use_norm=True, num_filters=(64,),
              voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE)
    classes = {"pfn": pp.PillarFeatureNet, "old": pp.PillarFeatureNetOld,
               "radius": pp.PillarFeatureNetRadius, "radius_height": pp.PillarFeatureNetRadiusHeight}
    with torch.no_grad():
        for name, cls in classes.items():
            for wd in (False, True):
                net = cls(with_distance=wd, **kw)
                net.pfn_layers = torch.nn.ModuleList([])
                dec = net(torch.from_numpy(voxels.copy()), torch.from_numpy(num), torch.from_numpy(coors))
                dec = dec.numpy()
                out["dec_%s_%d" % (name, int(wd))] = dec
                mine = pillar_oracle.decorate(voxels, num, coors, synth.PILLAR_VOXEL_SIZE,
                                              synth.PILLAR_RANGE, variant=name, with_distance=wd)
                err = np.abs(mine - dec).max()
                print("decorate", name, wd, dec.shape, "restatement max abs err", err)
                assert err <= 1e-5
        # full PillarFeatureNet with seeded weights, eval mode, non-trivial BN statistics
        torch.manual_seed(0)
        net = pp.PillarFeatureNet(with_distance=False, **kw)
        bn = net.pfn_layers[0].norm
        bn.running_mean.copy_(torch.randn(64) * 0.1)
        bn.running_var.copy_(torch.rand(64) + 0.5)
        bn.weight.copy_(torch.rand(64) + 0.5)
        bn.bias.copy_(torch.randn(64) * 0.1)
        net.eval()
        full = net(torch.from_numpy(voxels.copy()), torch.from_numpy(num), torch.from_numpy(coors)).numpy()
        out["pfn_weight"] = net.pfn_layers[0].linear.weight.numpy().copy()
        out["pfn_bn_gamma"] = bn.weight.numpy().copy()
        out["pfn_bn_beta"] = bn.bias.numpy().copy()
        out["pfn_bn_mean"] = bn.running_mean.numpy().copy()
        out["pfn_bn_var"] = bn.running_var.numpy().copy()
        out["pfn_out"] = full
        mine = pillar_oracle.pfn_layer_eval(out["dec_pfn_0"], out["pfn_weight"], out["pfn_bn_gamma"],
                                            out["pfn_bn_beta"], out["pfn_bn_mean"], out["pfn_bn_var"])
        print("pfn full", full.shape, "restatement max abs err", np.abs(mine - full).max())
        # SimpleVoxel / SimpleVoxelRadius (voxel_encoder.py:207-255)
        sv = ve.SimpleVoxel(num_input_features=4)
        out["simple_voxel"] = sv(torch.from_numpy(voxels), torch.from_numpy(num), None).numpy()
        svr = ve.SimpleVoxelRadius(num_input_features=4)
        out["simple_voxel_radius"] = svr(torch.from_numpy(voxels), torch.from_numpy(num), None).numpy()
    np.savez_compressed(os.path.join(GOLD, "ref_pillar_decorate.npz"), **out)

    # scatter
    rng = np.random.default_rng(11)
    B, C, ny, nx, P = 3, 16, 24, 40, 300
    cells = rng.permutation(B * ny * nx)[:P]
    b, rem = cells // (ny * nx), cells % (ny * nx)
    sc_coords = np.stack([b, np.zeros_like(b), rem // nx, rem % nx], axis=1).astype(np.int32)
    feats = rng.normal(size=(P, C)).astype(np.float32)
    with torch.no_grad():
        sc = pp.PointPillarsScatter(output_shape=[B, 1, ny, nx, C], num_input_features=C)
        canvas = sc(torch.from_numpy(feats), torch.from_numpy(sc_coords), B).numpy()
    mine = pillar_oracle.scatter(feats, sc_coords, B, ny, nx)
    assert np.array_equal(mine, canvas)
    np.savez_compressed(os.path.join(GOLD, "ref_scatter.npz"), feats=feats, coords=sc_coords,
                        canvas=canvas, shape=np.array([B, C, ny, nx]))
    print("scatter", canvas.shape, "restatement equal")


def gen_voxel_hashes():
    res = {}
    cloud11 = synth.multisweep_cloud(11)
    cloud20 = synth.multisweep_cloud(20)
    single = synth.fixture_points_nx4()
    cases = {
        "c2_11": (cloud11, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
        "c2_20": (cloud20, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
        "c3_single": (single, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
        "c3_11": (cloud11, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
        "c3_20": (cloud20, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
        "c3_20_train25000": (cloud20, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 25000),
    }
    for name, (pts, vs, rg, T, mv) in cases.items():
        orc = voxel_oracle.VoxelOracle(vs, rg, T, mv)
        for mode in ("continue", "break"):
            v, c, n = orc.generate(pts, overflow=mode)
            res["%s_%s" % (name, mode)] = {
                "n_points": int(pts.shape[0]), "voxel_num": int(v.shape[0]),
                "kept_points": int(orc.last_kept_points), "stored_points": int(n.sum()),
                "sha_voxels": sha16(v), "sha_coords": sha16(c), "sha_num": sha16(n)}
            print(name, mode, res["%s_%s" % (name, mode)])
    with open(os.path.join(GOLD, "voxel_oracle_hashes.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    gen_c1()
    gen_simplevis()
    gen_pillar()
    gen_voxel_hashes()
