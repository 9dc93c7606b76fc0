"""GPU: no entry point writes outside the buffers it was given.  Every output is carved out of a
larger allocation whose head and tail are filled with a sentinel; the sentinels must survive.
(compute-sanitizer is closed on this pool, so the bounds are checked this way.)"""
import ctypes

import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu

GUARD = 4096          # elements on each side
SENT_F = 12345.678
SENT_I = -123456789


class Guarded:
    def __init__(self, shape, dtype, device):
        import torch
        n = int(np.prod(shape)) if len(shape) else 1
        self.sent = SENT_F if dtype.is_floating_point else (77 if dtype == torch.uint8 else SENT_I)
        self.full = torch.full((n + 2 * GUARD,), self.sent, dtype=dtype, device=device)
        self.t = self.full[GUARD:GUARD + n].view(shape) if n else self.full[GUARD:GUARD]
        self.n = n

    def intact(self):
        return bool((self.full[:GUARD] == self.sent).all()) and bool((self.full[GUARD + self.n:] == self.sent).all())


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import _native as nat
    from lyft3d_b200.voxel_generator import _make_config
    return torch, nat, nat.load(), nat.get_handle(0), _make_config


def test_pillarize_at_the_capacity_boundary_and_scatter(env):
    torch, nat, lib, h, make_cfg = env
    dev = torch.device("cuda", 0)
    F = 3
    frames = [synth.c5_frame(500 + f)[:30000] for f in range(F)]
    pts = torch.from_numpy(np.concatenate(frames)).to(dev)
    offs = np.arange(F + 1, dtype=np.int64) * 30000
    T, V = 60, 30000
    cfg = make_cfg(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, 4, "continue", False)
    st = nat.current_stream_ptr(dev)
    for cap in (20000, 7001, 1):     # roomy, cut in the middle of a frame (odd: the pair path), one row
        dec = Guarded((cap, T, 9), torch.float32, dev)
        coords = Guarded((cap, 4), torch.int32, dev)
        num = Guarded((cap,), torch.int32, dev)
        vnum = Guarded((F,), torch.int32, dev)
        voff = Guarded((F + 1,), torch.int64, dev)
        nat.check(lib.lv_pillarize_concat(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, cap, 0.25, 0.25,
                                          -49.875, -49.875, 0, 0, dec.t.data_ptr(), coords.t.data_ptr(),
                                          num.t.data_ptr(), vnum.t.data_ptr(), voff.t.data_ptr(), st))
        torch.cuda.synchronize()
        for g in (dec, coords, num, vnum, voff):
            assert g.intact(), cap
        total = int(voff.t[F])
        assert total == int(vnum.t.sum()) and total > 7001
        vox = Guarded((cap, T, 4), torch.float32, dev)
        nat.check(lib.lv_voxelize_concat(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, cap,
                                         vox.t.data_ptr(), coords.t.data_ptr(), num.t.data_ptr(), vnum.t.data_ptr(),
                                         voff.t.data_ptr(), st))
        feats = Guarded((cap, 64), torch.float32, dev)
        w = torch.randn((64, 9), device=dev)
        sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
        nat.check(lib.lv_pillarize_pfn_concat(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, cap, 0.25, 0.25,
                                              -49.875, -49.875, 0, 0, w.data_ptr(), sc.data_ptr(), sh.data_ptr(), 64,
                                              feats.t.data_ptr(), coords.t.data_ptr(), num.t.data_ptr(),
                                              vnum.t.data_ptr(), voff.t.data_ptr(), st))
        canvas = Guarded((F, 64, 400, 400), torch.float32, dev)
        rows = min(total, cap)
        nat.check(lib.lv_pillar_scatter_dev(h.ptr, feats.t.data_ptr(), coords.t.data_ptr(), voff.t[F:].data_ptr(), cap, 64,
                                            F, 400, 400, canvas.t.data_ptr(), st))
        torch.cuda.synchronize()
        for g in (vox, feats, coords, num, vnum, voff, canvas):
            assert g.intact(), cap
        # the canvas holds exactly the kept rows
        assert int((canvas.t.abs().sum(dim=1) > 0).sum()) <= rows


def test_padded_voxelize_bev_and_ingest(env):
    torch, nat, lib, h, make_cfg = env
    dev = torch.device("cuda", 0)
    st = nat.current_stream_ptr(dev)
    frames = [synth.c5_frame(600)[:2049], np.zeros((0, 4), np.float32), synth.c5_frame(601)[:7000]]
    pts = torch.from_numpy(np.concatenate(frames)).to(dev)
    offs = np.zeros(4, np.int64)
    offs[1:] = np.cumsum([f.shape[0] for f in frames])
    for T, V, zero_tail in ((5, 900, True), (60, 2000, False), (3, 17, True)):
        cfg = make_cfg(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, 4, "break", zero_tail)
        vox = Guarded((3, V, T, 4), torch.float32, dev)
        coords = Guarded((3, V, 3), torch.int32, dev)
        num = Guarded((3, V), torch.int32, dev)
        vnum = Guarded((3,), torch.int32, dev)
        nat.check(lib.lv_voxelize(h.ptr, ctypes.byref(cfg), pts.data_ptr(), 3, offs.ctypes.data, vox.t.data_ptr(),
                                  coords.t.data_ptr(), num.t.data_ptr(), vnum.t.data_ptr(), st))
        torch.cuda.synchronize()
        for g in (vox, coords, num, vnum):
            assert g.intact(), (T, V)
    # BEV: odd cell count (scalar finalize) and the flat4 path
    shape3 = (ctypes.c_int32 * 3)
    for shape in ((333, 333, 3), (336, 336, 3)):
        cells = shape[0] * shape[1] * shape[2]
        raw = Guarded((3, cells), torch.float32, dev)
        nrm = Guarded((3, cells), torch.float32, dev)
        u8 = Guarded((3, cells), torch.uint8, dev)
        nat.check(lib.lv_bev_rasterize(h.ptr, pts.data_ptr(), 4, 3, offs.ctypes.data, None, None, 3, shape3(*shape),
                                       (ctypes.c_double * 3)(0.4, 0.4, 1.5), -2.0, 16.0, raw.t.data_ptr(),
                                       nrm.t.data_ptr(), u8.t.data_ptr(), None, None, st))
        torch.cuda.synchronize()
        for g in (raw, nrm, u8):
            assert g.intact(), shape
        assert int(raw.t.sum()) > 0
    # ingest: 1023 + 1024 + 1025 rows straddle the 1024-row TMA tiles
    raw5 = synth.load_fixture_raw()
    sizes = [1023, 1024, 1025]
    rows = torch.from_numpy(np.ascontiguousarray(raw5[:sum(sizes)])).to(dev)
    soffs = np.zeros(4, np.int64)
    soffs[1:] = np.cumsum(sizes)
    tm = np.stack([synth.sweep_transform(s) for s in range(3)]).reshape(3, 16)
    lag = np.array([0.0, 0.05, 0.1], np.float32)
    for cols in (4, 5):
        out = Guarded((sum(sizes), cols), torch.float32, dev)
        nat.check(lib.lv_ingest_sweeps(h.ptr, rows.data_ptr(), 3, soffs.ctypes.data, tm.ctypes.data, None, lag.ctypes.data,
                                       0, -1.0, cols, out.t.data_ptr(), st))
        torch.cuda.synchronize()
        assert out.intact() and bool((out.t != SENT_F).all())


def test_block_filter_and_target_raster_stay_inside(env):
    torch, nat, lib, h, make_cfg = env
    from lyft3d_b200.voxel_generator import _make_filter
    dev = torch.device("cuda", 0)
    st = nat.current_stream_ptr(dev)
    F, T, V = 3, 3, 9000          # V is hit by the larger frames
    frames = [synth.c5_frame(600 + f)[: 15000 + 15000 * f] for f in range(F)]
    pts = torch.from_numpy(np.concatenate(frames)).to(dev)
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
    vs, rg = (0.05, 0.05, 0.2), (-50, -50, -5, 50, 50, 3)
    for zero_tail in (False, True):
        cfg = make_cfg(vs, rg, T, V, 4, "continue", zero_tail)
        flt = _make_filter((1, 8, 0.2, 2.0))
        vox = Guarded((F, V, T, 4), torch.float32, dev)
        coords = Guarded((F, V, 3), torch.int32, dev)
        num = Guarded((F, V), torch.int32, dev)
        vnum = Guarded((F,), torch.int32, dev)
        mask = Guarded((F, V), torch.int32, dev)
        nat.check(lib.lv_voxelize_filtered(h.ptr, ctypes.byref(cfg), ctypes.byref(flt), pts.data_ptr(), F,
                                           offs.ctypes.data, vox.t.data_ptr(), coords.t.data_ptr(), num.t.data_ptr(),
                                           vnum.t.data_ptr(), mask.t.data_ptr(), st))
        torch.cuda.synchronize()
        for g in (vox, coords, num, vnum, mask):
            assert g.intact()
        k = vnum.t.cpu().numpy()
        assert (k > 0).all() and (k < V).all()
        if not zero_tail:      # rows at and beyond the survivors are left untouched
            for f in range(F):
                assert bool((vox.t[f, k[f]:] == SENT_F).all()) and bool((num.t[f, k[f]:] == SENT_I).all())
    # target raster: frames with 0, 1 and many boxes, boxes far outside the image
    corners, cls = synth.box_scene(42, 70, 120.0)
    box_offs = np.array([0, 0, 1, 70], dtype=np.int64)
    d_c = torch.from_numpy(corners).to(dev)
    d_k = torch.from_numpy(cls + 1).to(dev)
    shp = (ctypes.c_int32 * 3)(201, 157, 3)
    vsz = (ctypes.c_double * 3)(0.4, 0.4, 1.5)
    tgt = Guarded((3, 201, 157), torch.uint8, dev)
    nat.check(lib.lv_draw_boxes(h.ptr, d_c.data_ptr(), d_k.data_ptr(), 3, box_offs.ctypes.data, shp, vsz, -2.0,
                                tgt.t.data_ptr(), st))
    torch.cuda.synchronize()
    assert tgt.intact()
    assert not tgt.t[0].any() and tgt.t[2].any() and int(tgt.t.max()) <= 9


def test_pfn_training_kernels_and_hash_table_outputs(env):
    """lv_pillar_pfn_moments / lv_pillar_pfn_backward write exactly C_out + C_out(C_out+1)/2 and units x (2 + C_out)
    doubles; the voxelizer with the open-addressing table writes exactly the rows it was given."""
    torch, nat, lib, h, make_cfg = env
    dev = torch.device("cuda", 0)
    st = nat.current_stream_ptr(dev)
    P, T = 777, 20
    v = torch.rand((P, T, 4), device=dev)
    n = torch.randint(1, T + 1, (P,), device=dev, dtype=torch.int32)
    v = v * (torch.arange(T, device=dev)[None, :, None] < n[:, None, None])
    c = torch.zeros((P, 4), device=dev, dtype=torch.int32)
    c[:, 2] = torch.arange(P, device=dev) % 400
    c[:, 3] = torch.arange(P, device=dev) // 400
    for variant, wd, cout in ((0, 0, 9), (2, 0, 8), (3, 1, 10)):
        mom = Guarded((cout + cout * (cout + 1) // 2,), torch.float64, dev)
        nat.check(lib.lv_pillar_pfn_moments(h.ptr, v.data_ptr(), n.data_ptr(), c.data_ptr(), P, T, 4, 0.25, 0.25, -49.875,
                                            -49.875, variant, wd, mom.t.data_ptr(), st))
        for units in (32, 64, 128):
            w = torch.randn((units, cout), device=dev)
            sc, sh, mu, isd = (torch.rand((units,), device=dev) + 0.5 for _ in range(4))
            g = torch.randn((P, units), device=dev)
            acc = Guarded((units, 2 + cout), torch.float64, dev)
            nat.check(lib.lv_pillar_pfn_backward(h.ptr, v.data_ptr(), n.data_ptr(), c.data_ptr(), P, T, 4, 0.25, 0.25,
                                                 -49.875, -49.875, variant, wd, w.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                                 mu.data_ptr(), isd.data_ptr(), units, g.data_ptr(), acc.t.data_ptr(), st))
            torch.cuda.synchronize()
            assert acc.intact() and bool(torch.isfinite(acc.t).all()), (variant, units)
        torch.cuda.synchronize()
        assert mom.intact() and float(mom.t[0]) != SENT_F
        # the two complete operations
        units = 64
        w = torch.randn((units, cout), device=dev)
        gm, bt = torch.rand((units,), device=dev) + 0.5, torch.randn((units,), device=dev)
        out = Guarded((P, units), torch.float32, dev)
        stats = Guarded((5, units), torch.float32, dev)
        mom2 = Guarded((cout + cout * (cout + 1) // 2,), torch.float64, dev)
        nat.check(lib.lv_pillar_pfn_train_forward(h.ptr, v.data_ptr(), n.data_ptr(), c.data_ptr(), P, T, 4, 0.25, 0.25, -49.875,
                                                  -49.875, variant, wd, w.data_ptr(), gm.data_ptr(), bt.data_ptr(), 1e-3, units,
                                                  out.t.data_ptr(), stats.t.data_ptr(), mom2.t.data_ptr(), st))
        g = torch.randn((P, units), device=dev)
        dw, dg, db = Guarded((units, cout), torch.float32, dev), Guarded((units,), torch.float32, dev), Guarded((units,), torch.float32, dev)
        nat.check(lib.lv_pillar_pfn_train_backward(h.ptr, v.data_ptr(), n.data_ptr(), c.data_ptr(), P, T, 4, 0.25, 0.25, -49.875,
                                                   -49.875, variant, wd, w.data_ptr(), gm.data_ptr(), 1e-3, stats.t.data_ptr(),
                                                   mom2.t.data_ptr(), units, g.data_ptr(), dw.t.data_ptr(), dg.t.data_ptr(),
                                                   db.t.data_ptr(), st))
        torch.cuda.synchronize()
        for g_ in (out, stats, mom2, dw, dg, db):
            assert g_.intact() and bool(torch.isfinite(g_.t).all()), variant
        assert bool(torch.allclose(mom2.t, mom.t, rtol=1e-12, atol=1e-9))    # (float64 atomics: the order of the CTAs is free)
    # four clouds at the 0.05 m SECOND grid: automatic hash table
    F, V, T = 4, 5000, 5
    frames = [synth.c5_frame(600 + f)[:25000] for f in range(F)]
    pts = torch.from_numpy(np.concatenate(frames)).to(dev)
    offs = np.arange(F + 1, dtype=np.int64) * 25000
    cfg = make_cfg(synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, T, V, 4, "continue", True)
    vox = Guarded((F, V, T, 4), torch.float32, dev)
    co = Guarded((F, V, 3), torch.int32, dev)
    num = Guarded((F, V), torch.int32, dev)
    vnum = Guarded((F,), torch.int32, dev)
    nat.check(lib.lv_voxelize(h.ptr, ctypes.byref(cfg), pts.data_ptr(), F, offs.ctypes.data, vox.t.data_ptr(), co.t.data_ptr(),
                              num.t.data_ptr(), vnum.t.data_ptr(), st))
    torch.cuda.synchronize()
    for g_ in (vox, co, num, vnum):
        assert g_.intact()
    assert int(vnum.t.min()) > 0 and int(vnum.t.max()) <= V
