"""GPU parity of the batched engine entry points: device-side batch assembly
(lv_voxelize_concat), the fused voxelize+decorate (lv_pillarize_concat) and the whole
FrameBatchEngine step, against the oracle chain run frame by frame."""
import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-5   # see tests/test_gpu_pillar.py


@pytest.fixture(scope="module")
def setup():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200.engine import FrameBatchEngine
    from oracle import bev_oracle, pillar_oracle, voxel_oracle
    F, n = 6, 20000
    frames = [synth.c5_frame(f)[:n] for f in range(F)]
    eng = FrameBatchEngine(0, F, n)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    return eng, frames, pts, bev_oracle, pillar_oracle, voxel_oracle


def test_concat_voxelize_equals_merge_second_batch(setup):
    eng, frames, pts, bo, po, vo = setup
    eng.voxelize(pts)
    rows = eng.read_total_rows()
    orc = vo.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    res = [orc.generate(fr) for fr in frames]
    ref_v = np.concatenate([r[0] for r in res])
    ref_c = po.merge_batch_coords([r[1] for r in res])          # preprocess.py:44-50
    ref_n = np.concatenate([r[2] for r in res])
    assert rows == ref_v.shape[0]
    assert eng.voxel_num.cpu().tolist() == [r[0].shape[0] for r in res]
    offs = eng.voxel_offsets.cpu().numpy()
    assert offs.tolist() == np.concatenate([[0], np.cumsum([r[0].shape[0] for r in res])]).tolist()
    assert np.array_equal(eng.voxels[:rows].cpu().numpy().view(np.uint32), ref_v.view(np.uint32))
    assert np.array_equal(eng.coords[:rows].cpu().numpy(), ref_c)
    assert np.array_equal(eng.num_points[:rows].cpu().numpy(), ref_n)


def test_fused_pillarize_equals_voxelize_then_decorate(setup):
    eng, frames, pts, bo, po, vo = setup
    eng.voxelize(pts)
    rows = eng.read_total_rows()
    eng.decorate(rows)
    unfused = eng.decorated[:rows].clone()
    coords = eng.coords[:rows].clone()
    eng.decorated.zero_()
    eng.pillarize(pts)
    assert eng.read_total_rows() == rows
    fused = eng.decorated[:rows]
    assert bool((fused == unfused).all())                       # same warp-level arithmetic: identical bits
    assert bool((eng.coords[:rows] == coords).all())
    orc = vo.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    res = [orc.generate(fr) for fr in frames]
    ref = po.decorate(np.concatenate([r[0] for r in res]), np.concatenate([r[2] for r in res]),
                      po.merge_batch_coords([r[1] for r in res]), synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE)
    ok, worst, max_abs = po.decorate_mismatch(fused.cpu().numpy(), ref, np.concatenate([r[0] for r in res]),
                                              np.concatenate([r[2] for r in res]))
    print("fused decoration vs oracle: f_cluster max abs diff %.3g, %.3f of the summation-order bound" % (max_abs, worst))
    assert ok and max_abs <= 2e-5, (worst, max_abs)


def test_full_step_vs_cpu_path(setup):
    import torch
    eng, frames, pts, bo, po, vo = setup
    torch.manual_seed(3)
    eng.features.copy_(torch.randn_like(eng.features))
    rows = eng.step(pts)
    feats = eng.features[:rows].cpu().numpy()
    orc = vo.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    coords = po.merge_batch_coords([orc.generate(fr)[1] for fr in frames])
    ref_canvas = po.scatter(feats, coords, len(frames), 400, 400)
    assert np.array_equal(eng.canvas.cpu().numpy(), ref_canvas)
    for f, fr in enumerate(frames):
        bev = bo.create_voxel_pointcloud(np.ascontiguousarray(fr.T), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                         synth.BEV_Z_OFFSET)
        norm = bo.normalize_voxel_intensities(bev)
        assert np.array_equal(eng.bev_norm[f].cpu().numpy(), norm)
        assert np.array_equal(eng.bev_u8[f].cpu().numpy(), bo.quantize_u8(norm))


def test_capacity_overflow_is_reported(setup):
    import torch
    from lyft3d_b200 import _native as nat
    from lyft3d_b200.engine import FrameBatchEngine
    eng, frames, pts, *_ = setup
    small = FrameBatchEngine(0, eng.F, eng.n, voxel_capacity=1000)
    small.voxelize(pts)
    with pytest.raises(nat.LyftVoxelError):
        small.read_total_rows()
    torch.cuda.synchronize()


def test_shard_frames():
    from lyft3d_b200.engine import shard_frames
    got = sorted(sum((shard_frames(8192, r, 8) for r in range(8)), []))
    assert got == list(range(8192))
    assert shard_frames(10, 1, 4) == [1, 5, 9]


def test_pipelined_engine_equals_serial_steps():
    """Stream-pipelined steps (no host sync, device-side pillar count, two buffer sets) leave
    exactly the bits of FrameBatchEngine.step."""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine, PipelinedEngine
    F = 5
    batches = []
    for b in range(3):
        frames = [synth.c5_frame(300 + 10 * b + f) for f in range(F)]
        batches.append(torch.from_numpy(np.concatenate(frames)).cuda())
    n = batches[0].shape[0] // F
    eng = FrameBatchEngine(0, F, n)
    torch.manual_seed(7)
    eng.features.copy_(torch.randn_like(eng.features))
    want = []
    for pts in batches:
        rows = eng.step(pts)
        want.append((rows, eng.bev_u8.clone(), eng.canvas.clone(), eng.coords[:rows].clone(),
                     eng.decorated[:rows].clone()))
    pipe = PipelinedEngine(eng)
    for b, pts in enumerate(batches):
        st = pipe.submit(pts)
        pipe.drain()
        torch.cuda.synchronize()
        rows, u8, canvas, coords, dec = want[b]
        assert int(st["voxel_offsets"][F]) == rows
        assert bool((eng.bev_u8 == u8).all()) and bool((eng.canvas == canvas).all())
        assert bool((st["coords"][:rows] == coords).all()) and bool((st["decorated"][:rows] == dec).all())
    # back to back without draining in between: the last step's outputs are still right
    for pts in batches:
        st = pipe.submit(pts)
    pipe.drain()
    torch.cuda.synchronize()
    rows, u8, canvas, coords, dec = want[-1]
    assert bool((eng.canvas == canvas).all()) and bool((eng.bev_u8 == u8).all())
    assert bool((st["coords"][:rows] == coords).all())


def test_sub_batching_gives_identical_results():
    """Workspace budgets split a batch into sub-batches (dense-map limit for the voxelizer, frames in
    flight for the BEV counts); the outputs - including the running row offsets of the
    concatenated layout across sub-batch borders - must not change."""
    import torch
    from lyft3d_b200 import _native as nat
    from lyft3d_b200.engine import FrameBatchEngine
    F = 7
    frames = [synth.c5_frame(700 + f)[: 20000 + 3000 * f] for f in range(F)]
    n = 20000
    pts = torch.from_numpy(np.concatenate([fr[:n] for fr in frames])).cuda()
    eng = FrameBatchEngine(0, F, n)
    h = nat.get_handle(0)
    eng.step(pts)
    want = (eng.total_rows, eng.bev_u8.clone(), eng.bev_norm.clone(), eng.coords[:eng.total_rows].clone(),
            eng.decorated[:eng.total_rows].clone(), eng.voxel_offsets.clone(), eng.canvas.clone())
    try:
        h.set_option("vox_dense_map_limit_bytes", 2 * 400 * 400 * 4)      # two frames per sub-batch
        h.set_option("bev_frames_in_flight", 3)
        rows = eng.step(pts)
    finally:
        h.set_option("vox_dense_map_limit_bytes", 0)
        h.set_option("bev_frames_in_flight", 0)
    assert rows == want[0]
    assert bool((eng.bev_u8 == want[1]).all()) and bool((eng.bev_norm == want[2]).all())
    assert bool((eng.coords[:rows] == want[3]).all()) and bool((eng.decorated[:rows] == want[4]).all())
    assert bool((eng.voxel_offsets == want[5]).all()) and bool((eng.canvas == want[6]).all())


def test_full_size_batch_properties():
    """BASELINE configs[4] at bench size (128 frames x 53,146 points per step): size-independent
    properties checked against an independent torch computation on the same device.
      BEV      per-frame sum of the raw counts == number of points whose truncated fp64 voxel
               coordinates are in bounds; u8 == rint(clip(raw/16)*255)
      pillars  voxel_num[f] == min(#distinct in-range cells of frame f, V); coordinates unique per
               frame; row 0 of a frame is the cell of its first in-range point (first-come order);
               sum(num_points) == in-range points when no pillar overflows T
      canvas   exactly one non-zero column per pillar (features are non-zero)"""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine
    F = 128
    base = np.concatenate([synth.c5_frame(800 + f) for f in range(8)])
    n = base.shape[0] // 8
    host = np.tile(base, (F // 8, 1))
    # make the 128 frames distinct: a per-frame shift of x
    host = host.reshape(F, n, 4).copy()
    host[:, :, 0] += (np.arange(F, dtype=np.float32) * 0.03)[:, None]
    pts = torch.from_numpy(host.reshape(F * n, 4)).cuda()
    eng = FrameBatchEngine(0, F, n)
    eng.features.uniform_(0.5, 1.5)
    raw = torch.empty((F,) + eng.bev_shape, dtype=torch.float32, device="cuda")
    from lyft3d_b200 import bev
    bev.rasterize_frames(pts, eng.offsets, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET, want=("raw",),
                         out={"raw": raw})
    rows = eng.step(pts)
    p3 = pts.view(F, n, 4)[:, :, :3].double()
    vs = torch.tensor(synth.BEV_VOXEL_SIZE, dtype=torch.float64, device="cuda")
    m = 1.0 / vs
    t = torch.tensor([synth.BEV_SHAPE[0] / 2, synth.BEV_SHAPE[1] / 2, synth.BEV_SHAPE[2] / 2], dtype=torch.float64,
                     device="cuda") + torch.tensor([0.0, 0.0, synth.BEV_Z_OFFSET], dtype=torch.float64, device="cuda") / vs
    u = (p3 * m) + t                                   # un-fused multiply then add, float64
    c = torch.trunc(u)
    shp = torch.tensor(synth.BEV_SHAPE, dtype=torch.float64, device="cuda")
    inb = ((c >= 0) & (c < shp)).all(dim=2)
    assert torch.equal(raw.sum(dim=(1, 2, 3)).double(), inb.sum(dim=1).double())
    want_u8 = torch.round(torch.clamp(raw / 16.0, 0, 1) * 255).to(torch.uint8)
    assert torch.equal(eng.bev_u8, want_u8)
    # pillars
    lo = torch.tensor(synth.PILLAR_RANGE[:3], dtype=torch.float32, device="cuda")
    pv = torch.tensor(synth.PILLAR_VOXEL_SIZE, dtype=torch.float32, device="cuda")
    cf = torch.floor((pts.view(F, n, 4)[:, :, :3] - lo) / pv)
    grid = torch.tensor([400.0, 400.0, 1.0], device="cuda")
    ok = ((cf >= 0) & (cf < grid)).all(dim=2)
    cell = (cf[:, :, 1] * 400 + cf[:, :, 0]).long()
    vnum = eng.voxel_num.cpu().numpy()
    off = eng.voxel_offsets.cpu().numpy()
    assert int(off[F]) == rows == int(vnum.sum())
    total_kept = 0
    for f in (0, 1, 63, 127):
        cells_f = cell[f][ok[f]]
        uniq = torch.unique(cells_f)
        assert vnum[f] == min(int(uniq.numel()), eng.V)
        co = eng.coords[off[f]:off[f + 1]]
        assert bool((co[:, 0] == f).all())
        lin = co[:, 2].long() * 400 + co[:, 3].long()
        assert int(torch.unique(lin).numel()) == int(lin.numel())           # unique per frame
        assert int(lin[0]) == int(cells_f[0])                                # first-come order
        assert set(lin.tolist()) == set(uniq.tolist())
        nump = eng.num_points[off[f]:off[f + 1]]
        counts = torch.bincount(cells_f, minlength=160000)[lin]
        assert torch.equal(nump.long(), torch.clamp(counts, max=eng.T))
        total_kept += int(nump.sum())
    nz = (eng.canvas.abs().sum(dim=1) > 0).sum()
    assert int(nz) == rows


def test_host_pipeline_every_step_matches_the_serial_engine():
    """HostPipeline (pinned host points in; PNG files, file sizes and pillar counts written into mapped host
    memory; no host sync inside the loop): the host-side results of EVERY step equal FrameBatchEngine.step on
    the same batch - the per-slot voxel_num / offsets keep step i+1 from overwriting what step i hands out."""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine, HostPipeline
    from oracle import png_oracle
    F = 4
    host_batches, want = [], []
    for b in range(3):
        # batches of different density, so that consecutive steps have different pillar counts
        frames = [synth.c5_frame(500 + 10 * b + f)[:(b + 1) * 15000] for f in range(F)]
        frames = [np.concatenate([fr, np.full((45000 - fr.shape[0], 4), 1e6, np.float32)]) for fr in frames]
        hb = torch.from_numpy(np.concatenate(frames)).pin_memory()
        host_batches.append(hb)
    n = host_batches[0].shape[0] // F
    eng = FrameBatchEngine(0, F, n)
    torch.manual_seed(11)
    eng.features.copy_(torch.randn_like(eng.features))
    for hb in host_batches:
        rows = eng.step(hb.cuda())
        want.append((eng.voxel_num.cpu().numpy().copy(), eng.bev_u8.cpu().numpy().copy(), eng.canvas.clone()))
    assert len({tuple(w[0]) for w in want}) == 3
    pipe = HostPipeline(eng)
    seen = {}

    def consume(step, view):
        sizes = view["sizes"].copy()
        seen[step] = (view["voxel_num"].copy(), [view["png"][f, :sizes[f]].tobytes() for f in range(F)])

    steps = 7
    pipe.run(host_batches, steps, consume=consume)
    assert sorted(seen) == list(range(steps))
    for i in range(steps):
        vnum, u8, _ = want[i % 3]
        assert np.array_equal(seen[i][0], vnum), "step %d: pillar counts of another step" % i
        for f in range(F):
            assert seen[i][1][f] == png_oracle.encode_png(u8[f]), "step %d frame %d" % (i, f)
    assert bool((eng.canvas == want[(steps - 1) % 3][2]).all())
    assert pipe.d2h_bytes * 10 < eng.bev_u8.numel()
    # the dense-image variant (no PNG) hands out the u8 images themselves
    pipe2 = HostPipeline(eng, png=False)
    got = {}
    pipe2.run(host_batches, 4, consume=lambda i, v: got.__setitem__(i, (v["voxel_num"].copy(), v["u8"].copy())))
    for i in range(4):
        assert np.array_equal(got[i][0], want[i % 3][0]) and np.array_equal(got[i][1], want[i % 3][1])


@pytest.mark.parametrize("T", [17, 20, 31, 32, 33])
def test_fused_pillarize_small_max_points(T):
    """max_points between 17 and 33 with DENSE pillars (the two-pillars-per-warp path shares
    the warp's stage and needs T >= 32): fused voxelize+decorate == voxelize then decorate, bit for bit, and
    the fused PFN == the unfused PFN kernel."""
    import torch
    from lyft3d_b200 import pointpillars as pp
    from lyft3d_b200.engine import FrameBatchEngine
    F, n = 3, 30000
    rng = np.random.default_rng(T)
    frames = []
    for f in range(F):
        # clusters of 1..24 points per pillar around a few hundred pillar centres
        centres = rng.uniform(-40, 40, (600, 2))
        k = rng.integers(0, 600, n)
        xy = centres[k] + rng.uniform(-0.12, 0.12, (n, 2))
        frames.append(np.concatenate([xy, rng.uniform(-2, 2, (n, 1)), rng.random((n, 1))], axis=1).astype(np.float32))
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    eng = FrameBatchEngine(0, F, n, max_points=T)
    eng.voxelize(pts)
    rows = eng.read_total_rows()
    assert int(eng.num_points[:rows].max()) == T and int((eng.num_points[:rows] <= 16).sum()) > 0
    eng.decorate(rows)
    unfused = eng.decorated[:rows].clone()
    eng.decorated.fill_(float("nan"))
    eng.pillarize(pts)
    assert eng.read_total_rows() == rows
    assert bool((eng.decorated[:rows] == unfused).all())
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).cuda().eval()
    w, sc, sh = pp.fold_pfn_layer(net.pfn_layers[0])
    eng.pillar_features(pts, w, sc, sh)
    fused = eng.features[:rows].clone()
    ref = pp.pillar_pfn(eng.voxels[:rows], eng.num_points[:rows], eng.coords[:rows], net.vx, net.vy, net.x_offset,
                        net.y_offset, w, sc, sh)
    assert bool((fused == ref).all())


def test_decoration_bound_holds_on_the_frame_that_fails_a_fixed_atol():
    """synth.c5_frame(1031) holds a pillar of 14 points 46.9 m from the sensor whose f_cluster differs from the numpy
    oracle by 1.14e-5 = 3 ulp(46.9 m) - more than a fixed atol of 1e-5, a fraction of what two float32 summation
    orders of 14 such values may differ by (oracle/pillar_oracle.py::decorate_mismatch).  Found by bench.py's post-run
    check on rank 7 of an 8-GPU run; every other output of the frame is bit-exact."""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine
    from oracle import pillar_oracle as po, voxel_oracle as vo
    fr = synth.c5_frame(1031)
    eng = FrameBatchEngine(0, 1, fr.shape[0])
    eng.pillarize(torch.from_numpy(fr).cuda())
    rows = eng.read_total_rows()
    v, c, n = vo.points_to_voxel(fr, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    assert rows == v.shape[0] and np.array_equal(eng.coords[:rows, 1:].cpu().numpy(), c)
    ref = po.decorate(v, n, po.merge_batch_coords([c]), synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE)
    got = eng.decorated[:rows].cpu().numpy()
    ok, worst, max_abs = po.decorate_mismatch(got, ref, v, n)
    print("frame 1031: f_cluster max abs diff %.3g = %.3f of the summation-order bound" % (max_abs, worst))
    assert ok and 1e-5 < max_abs <= 2e-5


def test_pipelined_soak_400_steps_leaves_every_workspace_clean():
    """400 pipelined steps over four different batches without a single host synchronisation, then one more of
    each batch checked bit for bit against the serial engine: the "all empty between calls" invariants of the
    first[] map, the BEV counts and dirty bitmap and the cell->pillar map, and the two buffer sets of the pipeline,
    survive a long run (a stale entry anywhere would show up as a wrong pillar, count or canvas cell)."""
    import torch
    from lyft3d_b200.engine import FrameBatchEngine, PipelinedEngine
    F = 8
    batches = []
    for b in range(4):
        frames = [synth.c5_frame(400 + 8 * b + f)[: 30000 + 2500 * ((b + f) % 5)] for f in range(F)]
        n = 30000
        batches.append(torch.from_numpy(np.concatenate([fr[:n] for fr in frames])).cuda())
    eng = FrameBatchEngine(0, F, n)
    torch.manual_seed(11)
    eng.features.copy_(torch.randn_like(eng.features))
    want = []
    for pts in batches:
        rows = eng.step(pts)
        want.append((rows, eng.bev_u8.clone(), eng.bev_norm.clone(), eng.canvas.clone(), eng.coords[:rows].clone(),
                     eng.num_points[:rows].clone(), eng.decorated[:rows].clone()))
    pipe = PipelinedEngine(eng)
    for i in range(400):
        pipe.submit(batches[(i * 7) % 4])
    for b, pts in enumerate(batches):
        st = pipe.submit(pts)
        pipe.drain()
        torch.cuda.synchronize()
        rows, u8, norm, canvas, coords, num, dec = want[b]
        assert int(st["voxel_offsets"][F]) == rows
        assert bool((eng.bev_u8 == u8).all()) and bool((eng.bev_norm == norm).all()) and bool((eng.canvas == canvas).all())
        assert bool((st["coords"][:rows] == coords).all()) and bool((st["num_points"][:rows] == num).all())
        assert bool((st["decorated"][:rows] == dec).all())
