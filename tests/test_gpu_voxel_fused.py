"""The fused voxelizer prologue (vx_fused_kernel: K1-K5 in one launch with per-frame arrival counters, look-back
and polling instead of kernel boundaries; opt-in by lv_set_option("vox_fused_prologue", 1), frames of <= 64
chunks) must be bit-identical to the default five-kernel prologue and to the oracle: ragged frames, one-point frames,
max_voxels hit under both overflow rules, four and five features per point, the raw-voxel, fused-decoration and
mean outputs, many repetitions on the same handle (the counters and descriptors must return to zero)."""
import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import _native as nat
    from lyft3d_b200 import voxel_generator as vg
    from oracle import voxel_oracle as vo
    return torch, nat, vg, vo


def _frames(sizes):
    frames = []
    for i, n in enumerate(sizes):
        base = synth.c5_frame(40 + i)
        while n > base.shape[0]:
            base = np.concatenate([base, synth.c5_frame(90 + i + len(frames))])
        frames.append(np.ascontiguousarray(base[:n]))
    return frames


def _both(nat, fn):
    h = nat.get_handle(0)
    outs = []
    for on in (0, 1):
        h.set_option("vox_fused_prologue", on)
        try:
            outs.append(fn())
        finally:
            h.set_option("vox_fused_prologue", 0)
    return outs


@pytest.mark.parametrize("mode", ["continue", "break"])
@pytest.mark.parametrize("T,V", [(60, 30000), (7, 3000), (5, 700)])
def test_fused_prologue_equals_five_kernels_and_oracle(env, mode, T, V):
    torch, nat, vg, vo = env
    sizes = [53146, 1, 31, 2048, 20000, 65536, 40001, 100000, 2049]
    frames = _frames(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    launches = []

    def run():
        h = nat.get_handle(0)
        l0 = h.launches()
        out = vg.voxelize_frames(pts, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode,
                                 zero_tail=True)
        launches.append(h.launches() - l0)
        return out
    a, b = _both(nat, run)
    assert launches[1] < launches[0], "the fused prologue did not run (launch counts %r)" % (launches,)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    vnum = b[3].cpu().numpy()
    for f in (0, 1, 4, 7, 8):
        v, c, n = vo.points_to_voxel(frames[f], synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, T, V, overflow=mode)
        k = int(vnum[f])
        assert k == v.shape[0]
        assert np.array_equal(b[1][f, :k].cpu().numpy(), c)
        assert np.array_equal(b[2][f, :k].cpu().numpy(), n)
        assert np.array_equal(b[0][f, :k].cpu().numpy().view(np.uint32), v.view(np.uint32))


def test_fused_prologue_five_features_and_fine_grid(env):
    torch, nat, vg, vo = env
    frames = _frames([53146, 30000])
    offs = np.array([0, 53146, 83146], dtype=np.int64)
    p4 = torch.from_numpy(np.concatenate(frames)).cuda()
    p5 = torch.cat([p4, torch.rand((p4.shape[0], 1), device="cuda")], dim=1).contiguous()
    for pts, vs, rg, T, V in ((p5, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
                              (p4, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
                              (p4[1:], synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 2000)):   # unaligned rows: no TMA
        o = offs if pts.shape[0] == p4.shape[0] else np.array([0, 53145, 83145], dtype=np.int64)
        a, b = _both(nat, lambda: vg.voxelize_frames(pts.contiguous(), o, vs, rg, T, V, zero_tail=True))
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    v, c, n = vo.points_to_voxel(frames[0], synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000)
    b = vg.voxelize_frames(p4, offs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000, zero_tail=True)
    k = int(b[3][0])
    assert k == v.shape[0] and np.array_equal(b[1][0, :k].cpu().numpy(), c)
    assert np.array_equal(b[0][0, :k].cpu().numpy().view(np.uint32), v.view(np.uint32))


def test_fused_prologue_batch_engine_outputs_and_repetition(env):
    """128 C5-shaped frames through the fused decoration path, 20 times on one handle: identical every time
    and identical to the five-kernel prologue (the arrival counters and look-back descriptors return to zero)."""
    torch, nat, vg, vo = env
    from lyft3d_b200 import engine as eng_mod
    F = 32
    base = [synth.c5_frame(f) for f in range(4)]
    pts = torch.from_numpy(np.concatenate([base[f % 4] for f in range(F)])).cuda()
    eng = eng_mod.FrameBatchEngine(0, F, base[0].shape[0])
    h = eng.h

    def snap():
        eng.pillarize(pts)
        rows = eng.read_total_rows()
        return (rows, eng.decorated[:rows].clone(), eng.coords[:rows].clone(), eng.num_points[:rows].clone(),
                eng.voxel_num.clone())
    ref = snap()
    h.set_option("vox_fused_prologue", 1)
    try:
        for it in range(20):
            got = snap()
            assert got[0] == ref[0], it
            for x, y in zip(got[1:], ref[1:]):
                assert torch.equal(x, y), it
    finally:
        h.set_option("vox_fused_prologue", 0)
    # mean output (SimpleVoxel) through the same prologue
    offs = np.arange(F + 1, dtype=np.int64) * base[0].shape[0]
    a, b = _both(nat, lambda: vg.voxelize_mean_frames(pts, offs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 20000))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_hash_table_equals_dense_map_and_oracle(env, mode):
    """first[] as an open-addressing table sized by the points (automatic for grids whose dense map exceeds 48 MB
    per frame - the 0.05 m SECOND grid - lv_set_option("vox_hash_map", 1 / -1) forces / forbids it): bit-identical
    to the dense map on the SECOND and the pillar grids, several clouds per launch, and equal to the oracle."""
    torch, nat, vg, vo = env
    h = nat.get_handle(0)
    sizes = [200000, 53146, 1, 120001]
    frames = _frames(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    for vs, rg, T, V in ((synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
                         (synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 2500),
                         (synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)):
        outs = {}
        for opt in (-1, 1, 0):
            h.set_option("vox_hash_map", opt)
            try:
                outs[opt] = vg.voxelize_frames(pts, offs, vs, rg, T, V, overflow=mode, zero_tail=True)
                again = vg.voxelize_frames(pts, offs, vs, rg, T, V, overflow=mode, zero_tail=True)   # table left empty?
            finally:
                h.set_option("vox_hash_map", 0)
            for x, y in zip(outs[opt], again):
                assert torch.equal(x, y), opt
        for opt in (1, 0):
            for x, y in zip(outs[-1], outs[opt]):
                assert torch.equal(x, y), opt
        vnum = outs[1][3].cpu().numpy()
        for f in (0, 2, 3):
            v, c, n = vo.points_to_voxel(frames[f], vs, rg, T, V, overflow=mode)
            k = int(vnum[f])
            assert k == v.shape[0]
            assert np.array_equal(outs[1][1][f, :k].cpu().numpy(), c)
            assert np.array_equal(outs[1][2][f, :k].cpu().numpy(), n)
            assert np.array_equal(outs[1][0][f, :k].cpu().numpy().view(np.uint32), v.view(np.uint32))


def test_two_level_bin_scan_equals_single_cta_scan(env):
    """K4 in two levels (one CTA per (frame, bin) row, then one per frame; automatic when a frame's [bin][chunk] table
    exceeds 16 k counters, i.e. for 1 M-point clouds) == the single-CTA scan: forced on and off on small ragged frames,
    an empty frame, both overflow rules, and on a 584 k-point cloud where it is the default."""
    torch, nat, vg, vo = env
    h = nat.get_handle(0)
    sizes = [53146, 0, 1, 2049, 70000]
    frames = _frames(sizes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    big = torch.from_numpy(synth.multisweep_cloud(11)).cuda()
    boffs = np.array([0, big.shape[0]], dtype=np.int64)
    for mode in ("continue", "break"):
        for p_, o_, vs, rg, T, V in ((pts, offs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
                                     (pts, offs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 3000),
                                     (big, boffs, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000),
                                     (big, boffs, synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000)):
            outs = []
            for opt in (-1, 1, 0):
                h.set_option("vox_two_level_scan", opt)
                try:
                    outs.append(vg.voxelize_frames(p_, o_, vs, rg, T, V, overflow=mode, zero_tail=True))
                finally:
                    h.set_option("vox_two_level_scan", 0)
            for o in outs[1:]:
                for x, y in zip(outs[0], o):
                    assert torch.equal(x, y), (mode, T)
