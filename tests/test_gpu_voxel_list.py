"""GPU parity of the voxelizer's LIST PATH (csrc/lv_voxel_list.cuh: vl_cells / vl_assign / vl_rows) against the
CPU oracle AND against the six-kernel path it replaces for small grids, bit for bit: voxel order, slot order,
both max_voxels rules, caps that bite, lists longer than 64 points, frames without points, padded and
concatenated layouts, fused decoration and PFN outputs."""
import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import _native as nat, voxel_generator as vg
    from oracle import voxel_oracle as vo
    return torch, nat, vg, vo


def _both_paths(nat, fn):
    """fn() under the list path (vox_list_path = 1) and under the default six-kernel path."""
    h = nat.get_handle(0)
    h.set_option("vox_list_path", 1)
    try:
        a = fn()
    finally:
        h.set_option("vox_list_path", 0)
    b = fn()
    return a, b


def _clustered(n, seed, spread=0.12, centres=500, extent=45.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (centres, 2))
    k = rng.integers(0, centres, n)
    xy = c[k] + rng.uniform(-spread, spread, (n, 2))
    return np.concatenate([xy, rng.uniform(-3, 3, (n, 1)), rng.random((n, 1))], axis=1).astype(np.float32)


@pytest.mark.parametrize("mode", ["continue", "break"])
@pytest.mark.parametrize("T,V", [(60, 30000), (60, 700), (35, 3000), (5, 2000), (64, 100), (1, 50)])
def test_padded_layout_vs_oracle_and_six_kernel_path(env, mode, T, V):
    torch, nat, vg, vo = env
    frames = [synth.c5_frame(900)[:30000], _clustered(40000, 1), np.zeros((0, 4), np.float32),
              synth.c5_frame(901)[:777], _clustered(2500, 2, centres=3)]           # 3 centres: lists of ~800 points
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()

    def run():
        v, c, n, vn = vg.voxelize_frames(pts, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode)
        return v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy(), vn.cpu().numpy()

    a, b = _both_paths(nat, run)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)
    for f, fr in enumerate(frames):
        rv, rc, rn = vo.points_to_voxel(fr, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode)
        k = rv.shape[0]
        assert a[3][f] == k
        assert np.array_equal(a[1][f, :k], rc) and np.array_equal(a[2][f, :k], rn)
        assert np.array_equal(a[0][f, :k].view(np.uint32), rv.view(np.uint32))
        assert not a[0][f, k:].any() and not a[2][f, k:].any()                       # zero tail
    # a second call sees a clean map (touched-cell reset by vl_rows, also for voxels beyond max_voxels)
    again = run()
    assert all(np.array_equal(x, y) for x, y in zip(again, a))


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_engine_outputs_equal_the_six_kernel_path(env, mode):
    """Concatenated layout, fused decoration and fused PFN through FrameBatchEngine: identical bits on both paths."""
    torch, nat, vg, vo = env
    from lyft3d_b200 import pointpillars as pp
    from lyft3d_b200.engine import FrameBatchEngine
    F, n = 5, 30000
    frames = [synth.c5_frame(700 + f)[:n] if f % 2 == 0 else _clustered(n, 10 + f, centres=900) for f in range(F)]
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().eval()
    w, sc, sh = pp.fold_pfn_layer(net.pfn_layers[0])
    eng = FrameBatchEngine(0, F, n, max_voxels=2500, overflow=mode)     # the cap bites on the clustered frames

    def run():
        out = {}
        eng.voxelize(pts)
        rows = eng.read_total_rows()
        out["vox"] = (eng.voxels[:rows].clone(), eng.coords[:rows].clone(), eng.num_points[:rows].clone(),
                      eng.voxel_num.clone(), eng.voxel_offsets.clone())
        eng.decorated.fill_(float("nan"))
        eng.pillarize(pts)
        assert eng.read_total_rows() == rows
        out["deco"] = (eng.decorated[:rows].clone(), eng.coords[:rows].clone(), eng.num_points[:rows].clone())
        eng.features.fill_(float("nan"))
        eng.pillar_features(pts, w, sc, sh)
        assert eng.read_total_rows() == rows
        out["pfn"] = (eng.features[:rows].clone(),)
        return out

    a, b = _both_paths(nat, run)
    for key in a:
        for x, y in zip(a[key], b[key]):
            assert x.shape == y.shape and bool((x.view(torch.int32) == y.view(torch.int32)).all() if x.dtype == torch.float32
                                               else (x == y).all()), key
    # and the oracle, frame by frame
    v, c, num, vn, offs = (t.cpu().numpy() for t in a["vox"])
    for f, fr in enumerate(frames):
        rv, rc, rn = vo.points_to_voxel(fr, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 2500, overflow=mode)
        s, e = offs[f], offs[f + 1]
        assert e - s == rv.shape[0] == vn[f]
        assert np.array_equal(c[s:e, 1:], rc) and (c[s:e, 0] == f).all() and np.array_equal(num[s:e], rn)
        assert np.array_equal(v[s:e].view(np.uint32), rv.view(np.uint32))


def test_capacity_smaller_than_the_rows(env):
    """Concatenated output with too little capacity: rows beyond it are not written, the totals still report them."""
    torch, nat, vg, vo = env
    from lyft3d_b200.engine import FrameBatchEngine
    F, n = 3, 20000
    pts = torch.from_numpy(np.concatenate([synth.c5_frame(40 + f)[:n] for f in range(F)])).cuda()
    big = FrameBatchEngine(0, F, n)
    big.pillarize(pts)
    rows = big.read_total_rows()
    small = FrameBatchEngine(0, F, n, voxel_capacity=rows // 2)
    small.decorated.fill_(7.0)
    guard = torch.full((1000,), 7.0, device="cuda")
    small.pillarize(pts)
    torch.cuda.synchronize()
    assert int(small.voxel_offsets[F].item()) == rows
    assert bool((small.decorated[:rows // 2] == big.decorated[:rows // 2]).all())
    assert bool((guard == 7.0).all())
