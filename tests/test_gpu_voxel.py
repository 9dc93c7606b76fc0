"""GPU parity: hard voxelizer (C-ABI, sm_100a kernels) vs the CPU oracle - all four
outputs bit-exact including voxel order, in both overflow modes."""
import hashlib
import json
import os

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def vg():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import voxel_generator
    return voxel_generator


@pytest.fixture(scope="module")
def vo():
    from oracle import voxel_oracle
    return voxel_oracle


def check_equal(res, ref):
    v, c, n = ref
    assert res["voxels"].shape == v.shape and res["voxels"].dtype == np.float32
    assert res["coordinates"].dtype == np.int32 and res["num_points_per_voxel"].dtype == np.int32
    assert np.array_equal(res["coordinates"], c)
    assert np.array_equal(res["num_points_per_voxel"], n)
    assert np.array_equal(res["voxels"].view(np.uint32), v.view(np.uint32))   # bit-exact incl. -0.0


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_c3_single_sweep_pillars(vg, vo, fixture_nx4, mode):
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, synth.PILLAR_MAX_POINTS,
                              max_voxels=synth.PILLAR_MAX_VOXELS, overflow=mode)
    res = gen.generate(fixture_nx4, synth.PILLAR_MAX_VOXELS)
    ref = vo.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000, overflow=mode)
    assert res["voxel_num"] == 8569
    check_equal(res, ref)
    res2 = gen.generate(fixture_nx4, synth.PILLAR_MAX_VOXELS)   # dense map was reset
    check_equal(res2, ref)


@pytest.mark.parametrize("mode", ["continue", "break"])
@pytest.mark.parametrize("case", ["c2_11", "c2_20", "c3_11", "c3_20", "c3_20_train25000"])
def test_frozen_configs(vg, vo, golden_dir, cloud11, cloud20, case, mode):
    with open(os.path.join(golden_dir, "voxel_oracle_hashes.json")) as f:
        fz = json.load(f)["%s_%s" % (case, mode)]
    pts = cloud11 if case.endswith("_11") else cloud20
    if case.startswith("c2"):
        vs, rg, T, mv = synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000
    else:
        vs, rg, T, mv = synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, (25000 if "25000" in case else 30000)
    gen = vg.VoxelGeneratorV2(vs, rg, T, max_voxels=mv, overflow=mode)
    res = gen.generate(pts, mv)
    assert res["voxel_num"] == fz["voxel_num"]
    assert int(res["num_points_per_voxel"].sum()) == fz["stored_points"]
    assert sha16(res["coordinates"]) == fz["sha_coords"]
    assert sha16(res["num_points_per_voxel"]) == fz["sha_num"]
    assert sha16(res["voxels"]) == fz["sha_voxels"]
    check_equal(res, vo.points_to_voxel(pts, vs, rg, T, mv, overflow=mode))


def test_generator_properties_and_padded_api(vg, vo, cloud11):
    gen = vg.VoxelGeneratorV2(list(synth.SECOND_VOXEL_SIZE), list(synth.SECOND_RANGE), 5, max_voxels=20000)
    assert gen.grid_size.tolist() == [1056, 1280, 40]
    assert gen.voxel_size.dtype == np.float32 and gen.point_cloud_range.dtype == np.float32
    assert gen.point_cloud_range[[0, 1, 3, 4]].tolist() == [0.0, -32.0, pytest.approx(52.8), 32.0]
    assert gen.max_num_points_per_voxel == 5
    import pickle
    gen = pickle.loads(pickle.dumps(gen))                       # plain picklable object
    res = gen.generate_multi_gpu(cloud11, 70000)                # cap not reached -> padding rows
    v, c, n, k = vo.VoxelOracle(synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 70000).generate(
        cloud11, padded=True)
    assert res["voxel_num"] == k
    assert res["voxels"].shape == (70000, 5, 4)
    assert np.array_equal(res["voxels"], v) and np.array_equal(res["coordinates"], c)
    assert np.array_equal(res["num_points_per_voxel"], n)
    # default max_voxels of the constructor is used when generate() gets none
    r2 = gen.generate(cloud11)
    assert r2["voxels"].shape[0] == 20000


def test_legacy_tuple_apis(vg, vo, fixture_nx4):
    ref = vo.points_to_voxel(fixture_nx4, (0.2, 0.2, 0.4), (0, -40, -3, 70.4, 40, 1), 35, 20000, overflow="break")
    v, c, n = vg.points_to_voxel(fixture_nx4, (0.2, 0.2, 0.4), (0, -40, -3, 70.4, 40, 1), 35, True, 20000)
    assert np.array_equal(v, ref[0]) and np.array_equal(c, ref[1]) and np.array_equal(n, ref[2])
    v, c2, n = vg.points_to_voxel(fixture_nx4, (0.2, 0.2, 0.4), (0, -40, -3, 70.4, 40, 1), 35, False, 20000)
    assert np.array_equal(c2, ref[1][:, ::-1])
    g1 = vg.VoxelGenerator((0.2, 0.2, 0.4), (0, -40, -3, 70.4, 40, 1), 35, 20000)
    t = g1.generate(fixture_nx4, 20000)
    assert isinstance(t, tuple) and np.array_equal(t[0], ref[0])


def test_cuda_tensor_in_out(vg, vo, fixture_nx4):
    import torch
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, max_voxels=30000)
    res = gen.generate(torch.from_numpy(fixture_nx4).cuda(), 30000)
    assert res["voxels"].is_cuda
    ref = vo.points_to_voxel(fixture_nx4, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    assert np.array_equal(res["voxels"].cpu().numpy(), ref[0])
    assert np.array_equal(res["coordinates"].cpu().numpy(), ref[1])
    assert np.array_equal(res["num_points_per_voxel"].cpu().numpy(), ref[2])


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_batched_frames_ragged(vg, vo, mode):
    """C5-style batch: ragged frames incl. an empty one and one with every point out of range;
    small max_voxels so that the cap and both overflow rules are exercised per frame."""
    frames = [synth.c5_frame(0), synth.c5_frame(1)[:30011], np.zeros((0, 4), np.float32),
              synth.c5_frame(2)[:2049], np.full((100, 4), 1e6, np.float32), synth.c5_frame(3)[:1]]
    frames += [synth.c5_frame(10 + f)[: 4000 + 37 * f] for f in range(20)]
    rows = np.concatenate(frames)
    offs = np.zeros(len(frames) + 1, np.int64)
    offs[1:] = np.cumsum([f.shape[0] for f in frames])
    T, V = 9, 2500
    voxels, coords, num, vnum = vg.voxelize_frames(rows, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V,
                                                   overflow=mode, zero_tail=True)
    orc = vo.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V)
    for f, fr in enumerate(frames):
        v, c, n, k = orc.generate(fr, overflow=mode, padded=True)
        assert vnum[f] == k, f
        assert np.array_equal(coords[f], c) and np.array_equal(num[f], n), f
        assert np.array_equal(voxels[f].view(np.uint32), v.view(np.uint32)), f


def test_generic_feature_count_and_three_passes(vg, vo, cloud11):
    # C = 5 features (non-float4 path) and max_voxels > 2^18 -> three radix passes
    rng = np.random.default_rng(9)
    pts5 = np.concatenate([cloud11[:300000], rng.normal(size=(300000, 1)).astype(np.float32)], axis=1)
    vs, rg, T, mv = (0.05, 0.05, 0.1), synth.SECOND_RANGE, 3, 300000
    gen = vg.VoxelGeneratorV2(vs, rg, T, max_voxels=mv)
    res = gen.generate(pts5, mv)
    check_equal(res, vo.points_to_voxel(pts5, vs, rg, T, mv))
    # C = 3
    pts3 = np.ascontiguousarray(cloud11[:100000, :3])
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 4, max_voxels=5000)
    check_equal(gen.generate(pts3, 5000), vo.points_to_voxel(pts3, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 4, 5000))


def test_edges_and_errors(vg, vo):
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, max_voxels=100)
    r = gen.generate(np.zeros((0, 4), np.float32), 100)
    assert r["voxels"].shape == (0, 5, 4) and r["coordinates"].shape == (0, 3)
    pts = np.array([[50.0, 0, 0, 0], [49.999, 0, 0, 0], [-50.0, 0, 0, 0], [-50.001, 0, 0, 0],
                    [0, 0, 10.0, 0], [0, 0, -10.0, 0], [np.nan, 0, 0, 0], [0, np.inf, 0, 0]], np.float32)
    r = gen.generate(pts, 100)
    assert r["coordinates"].tolist() == [[0, 200, 399], [0, 200, 0], [0, 200, 200]]
    # one voxel, more points than max_points: the FIRST T points in index order are kept
    same = np.tile(np.array([[1.01, 1.01, 0.0, 0.0]], np.float32), (1000, 1))
    same[:, 3] = np.arange(1000)
    r = gen.generate(same, 100)
    assert r["num_points_per_voxel"].tolist() == [5] and r["voxels"][0, :, 3].tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(Exception):
        vg.VoxelGeneratorV2((0.001, 0.001, 0.001), synth.PILLAR_RANGE, 5).generate(pts, 10)   # grid > 2^28 cells


def test_order_properties_at_full_size(vg, cloud11):
    """Size-independent properties at full size: coordinates unique, every stored point lies in
    its voxel, slots hold increasing point indices (channel 3 carries the index)."""
    pts = cloud11.copy()
    pts[:, 3] = np.arange(pts.shape[0], dtype=np.float32)     # exact below 2^24
    gen = vg.VoxelGeneratorV2(synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, max_voxels=60000)
    r = gen.generate(pts, 60000)
    v, c, n = r["voxels"], r["coordinates"], r["num_points_per_voxel"]
    assert len(np.unique(c, axis=0)) == c.shape[0] == 60000
    mask = np.arange(5)[None, :] < n[:, None]
    idx = np.where(mask, v[..., 3], np.inf)
    d = np.diff(idx, axis=1)
    assert np.all(d[mask[:, 1:]] > 0)                          # slots in point order
    assert np.all(np.diff(idx[:, 0]) > 0)                      # voxel order == order of first appearance
    lo = np.asarray(synth.SECOND_RANGE[:3], np.float32)
    vs = np.asarray(synth.SECOND_VOXEL_SIZE, np.float32)
    cc = np.floor((v[..., :3] - lo) / vs).astype(np.int32)[..., ::-1]
    assert np.all((cc == c[:, None, :])[mask])


@pytest.mark.parametrize("C", [4, 5])
def test_tma_staged_chunks_equal_plain_loads(vg, vo, cloud11, C):
    """K1 stages full 2048-row chunks by TMA bulk copies; with `disable_tma` the same kernel
    uses plain loads.  Both must match the oracle bit for bit (ragged frames: partial and
    unaligned chunks take the plain path inside the TMA launch too)."""
    from lyft3d_b200 import _native as nat
    rng = np.random.default_rng(5)
    base = cloud11[:40000]
    if C == 5:
        base = np.concatenate([base, rng.normal(size=(base.shape[0], 1)).astype(np.float32)], axis=1)
    sizes = [4096, 2049, 3, 10000, 2048, 6145]
    frames, at = [], 0
    for n in sizes:
        frames.append(np.ascontiguousarray(base[at:at + n]))
        at += n
    rows = np.concatenate(frames)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    T, V = 6, 1500
    h = nat.get_handle(0)
    orc = vo.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V)
    for off in (0, 1):
        h.set_option("disable_tma", off)
        try:
            voxels, coords, num, vnum = vg.voxelize_frames(rows, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE,
                                                           T, V, zero_tail=True)
        finally:
            h.set_option("disable_tma", 0)
        for f, fr in enumerate(frames):
            v, c, n, k = orc.generate(fr, padded=True)
            assert vnum[f] == k and np.array_equal(coords[f], c) and np.array_equal(num[f], n), (off, f)
            assert np.array_equal(voxels[f].view(np.uint32), v.view(np.uint32)), (off, f)


def test_unpad_multigpu_batch_equals_voxelnet_forward(vg, vo):
    """generate_multi_gpu padding (preprocess.py:311-317) -> merge_second_batch_multigpu
    (:60-88) -> the un-padding of VoxelNet.forward (voxelnet.py:346-358)."""
    import torch
    frames = [synth.c5_frame(40 + f)[: 9000 + 1500 * f] for f in range(4)] + [np.zeros((0, 4), np.float32)]
    V, T = 6000, 7
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, max_voxels=V)
    res = [gen.generate_multi_gpu(fr, V) for fr in frames]
    voxels = torch.from_numpy(np.stack([r["voxels"] for r in res])).cuda()
    num = torch.from_numpy(np.stack([r["num_points_per_voxel"] for r in res])).cuda()
    coors = torch.from_numpy(np.stack([np.pad(r["coordinates"], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                                       for i, r in enumerate(res)])).cuda()
    nv = torch.tensor([int(r["voxel_num"]) for r in res])
    v2, n2, c2 = vg.unpad_multigpu_batch(voxels, num, coors, nv)
    ref_v = torch.cat([voxels[i, :k] for i, k in enumerate(nv.tolist())], dim=0)
    ref_n = torch.cat([num[i, :k] for i, k in enumerate(nv.tolist())], dim=0)
    ref_c = torch.cat([coors[i, :k] for i, k in enumerate(nv.tolist())], dim=0)
    assert v2.shape == ref_v.shape and bool((v2.view(torch.int32) == ref_v.view(torch.int32)).all())
    assert bool((n2 == ref_n).all()) and bool((c2 == ref_c).all())


@pytest.mark.parametrize("cout", [4, 3])
def test_fused_mean_vfe_second_config(vg, vo, cloud11, cout):
    """C2 (0.05 m SECOND voxels, T=5, V=60000): voxelize + SimpleVoxel mean in one pipeline vs
    oracle voxelizer -> oracle SimpleVoxel (voxel_encoder.py:219-225)."""
    import torch
    from oracle import pillar_oracle as po
    frames = [cloud11[:200000], cloud11[200000:260000]]
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    offs = np.array([0, 200000, 260000], dtype=np.int64)
    mean, coords, num, vnum = vg.voxelize_mean_frames(pts, offs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE,
                                                      synth.SECOND_MAX_POINTS, synth.SECOND_MAX_VOXELS, cout)
    at = 0
    for b, fr in enumerate(frames):
        v, c, n = vo.points_to_voxel(fr, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, synth.SECOND_MAX_POINTS,
                                     synth.SECOND_MAX_VOXELS)
        k = v.shape[0]
        assert int(vnum[b]) == k
        assert np.array_equal(coords[at:at + k, 1:].cpu().numpy(), c) and bool((coords[at:at + k, 0] == b).all())
        assert np.array_equal(num[at:at + k].cpu().numpy(), n)
        np.testing.assert_allclose(mean[at:at + k].cpu().numpy(), po.simple_voxel_mean(v, n, cout), rtol=1e-6, atol=1e-6)
        at += k
    assert at == mean.shape[0]


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_cluster_frame_kernel_equals_the_five_kernel_prologue(vg, vo, mode):
    """lv_set_option("vox_frame_kernel", 1): K1-K5 of a frame inside one thread-block cluster (dense map in
    distributed shared memory).  Opt-in; must be bit-identical to the default path and to the oracle -
    ragged frames, an empty frame, max_voxels hit (both overflow rules), five features per point."""
    import torch
    from lyft3d_b200 import _native as nat
    h = nat.get_handle(0)
    rng = np.random.default_rng(21)
    sizes = [53146, 0, 1, 31, 20000, 65536, 40001]
    frames = []
    for i, n in enumerate(sizes):
        base = synth.c5_frame(40 + i)
        if n > base.shape[0]:
            base = np.concatenate([base, synth.c5_frame(90 + i)])
        frames.append(base[:n])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    for T, V in ((60, 30000), (7, 3000)):
        outs = []
        for on in (0, 1):
            h.set_option("vox_frame_kernel", on)
            try:
                outs.append(vg.voxelize_frames(pts, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V,
                                               overflow=mode, zero_tail=True))
            finally:
                h.set_option("vox_frame_kernel", 0)
        for a, b in zip(outs[0], outs[1]):
            assert torch.equal(a, b)
        vnum = outs[1][3].cpu().numpy()
        for f in (0, 4, 6):
            v, c, n = vo.points_to_voxel(frames[f], synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode)
            k = int(vnum[f])
            assert k == v.shape[0]
            assert np.array_equal(outs[1][1][f, :k].cpu().numpy(), c)
            assert np.array_equal(outs[1][2][f, :k].cpu().numpy(), n)
            assert np.array_equal(outs[1][0][f, :k].cpu().numpy().view(np.uint32), v.view(np.uint32))
    assert rng is not None
    # five features per point (the non-float4 load path)
    p5 = torch.cat([pts[:53146], torch.rand((53146, 1), device="cuda")], dim=1).contiguous()
    o5 = np.array([0, 53146], dtype=np.int64)
    res = []
    for on in (0, 1):
        h.set_option("vox_frame_kernel", on)
        try:
            res.append(vg.voxelize_frames(p5, o5, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000, overflow=mode))
        finally:
            h.set_option("vox_frame_kernel", 0)
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)


def test_generate_batch_equals_generate_per_cloud(vg):
    """VoxelGeneratorV2.generate_batch: several host clouds in one pass == generate() on each (also with the block filter)."""
    clouds = [synth.c5_frame(70 + i)[: 20000 + 9000 * i] for i in range(4)] + [np.zeros((0, 4), np.float32)]
    for kw in ({}, dict(block_filtering=True, block_factor=1, block_size=8, height_threshold=0.2)):
        vs, rg, T, V = ((0.05, 0.05, 0.2), (-50, -50, -5, 50, 50, 3), 3, 20000) if kw else (synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
        gen = vg.VoxelGeneratorV2(vs, rg, T, max_voxels=V, **kw)
        batch = gen.generate_batch(clouds, V)
        assert len(batch) == len(clouds)
        for c, b in zip(clouds, batch):
            one = gen.generate(c, V)
            assert b["voxel_num"] == one["voxel_num"]
            for key in ("voxels", "coordinates", "num_points_per_voxel"):
                assert b[key].dtype == one[key].dtype and b[key].shape == one[key].shape
                assert np.array_equal(b[key].view(np.uint8), one[key].view(np.uint8)), key
    assert gen.generate_batch([]) == []


def test_gpu_voxelizer_against_the_reference_points_to_bev(vg, fixture_nx4, golden_dir):
    """The GPU voxelizer pinned DIRECTLY to output of the reference's own code: the per-pillar point-count map that the
    reference's in-tree sibling of the voxelizer (second/second/utils/simplevis.py:9-108, executed by oracle/gen_golden.py
    into tests/golden/ref_simplevis_pillar.npz) draws for the bundled sweep - all 8,569 pillars / 47,732 points with
    max_voxels=40000, and the `break` rule of :46-50 with max_voxels=3000 (3,000 pillars / 7,764 points).  The point
    cap is set above the fullest pillar so that num_points_per_voxel is the count; the height map (highest point of
    every pillar above the range floor, in units of the voxel height) follows from the voxels."""
    g = np.load(os.path.join(golden_dir, "ref_simplevis_pillar.npz"))
    T = 1024
    assert int(g["density_full"].max()) < T
    for V, mode, key, nvox, npts in ((40000, "continue", "density_full", 8569, 47732), (3000, "break", "density_break3000", 3000, 7764)):
        voxels, coords, num, vnum = vg.voxelize_frames(fixture_nx4, np.array([0, fixture_nx4.shape[0]], dtype=np.int64),
                                                       synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode,
                                                       zero_tail=False)
        k = int(vnum[0])
        assert k == nvox and int(num[0][:k].sum()) == npts
        dm = np.zeros((400, 400), np.int64)
        dm[coords[0][:k, 1], coords[0][:k, 2]] = num[0][:k]
        assert np.array_equal(dm, g[key].astype(np.int64)), key
        if key == "density_full":
            # simplevis.py:52-57: bev_map[0] = max over the pillar's points of (z - z_lo) / voxel height, where positive
            vz = voxels[0][:k, :, 2]
            live = np.arange(T)[None, :] < num[0][:k, None]
            hn = ((vz - np.float32(synth.PILLAR_RANGE[2])) / np.float32(synth.PILLAR_VOXEL_SIZE[2])).astype(np.float32)
            hmax = np.where(live, hn, -np.inf).max(axis=1)
            hm = np.zeros((400, 400), np.float32)
            hm[coords[0][:k, 1], coords[0][:k, 2]] = np.maximum(hmax, 0).astype(np.float32)
            assert np.array_equal(hm, g["height_full"])


def test_gpu_voxelizer_against_the_reference_points_to_bev_3d_grids(vg, golden_dir):
    """The same pin at three-dimensional grids (tests/golden/ref_simplevis_3d.npz): counts per column, the highest
    point of every (z, y, x) cell, and - through the `break` rule at small caps - which cells exist at all."""
    from oracle import gen_golden_simplevis3d as gg
    from oracle import voxel_oracle as vo
    from test_oracle_voxel import _maps_from_voxels
    g = np.load(os.path.join(golden_dir, "ref_simplevis_3d.npz"))
    for name, first, n, vs, rg, mv in gg.CASES:
        pts = gg.case_points(first, n)
        T = 1024
        assert int(g[name + ".count"].max()) < T
        voxels, coords, num, vnum = vg.voxelize_frames(pts, np.array([0, n], dtype=np.int64), vs, rg, T, mv, overflow="break",
                                                       zero_tail=False)
        k = int(vnum[0])
        count, height = _maps_from_voxels(voxels[0][:k], coords[0][:k], num[0][:k], vs, rg, vo.grid_size(vs, rg))
        assert np.array_equal(count, g[name + ".count"].astype(np.int64)), name
        assert np.array_equal(height, g[name + ".height"]), name
