"""GPU: size-independent properties of the hot path at BASELINE's full sizes (SURVEY.md section 4:
"property tests (permutation/duplication/boundary points)") - no oracle involved, so they run at
sizes the CPU restatement would not finish in seconds."""
import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import bev, voxel_generator
    return torch, bev, voxel_generator


def test_bev_is_a_permutation_invariant_additive_histogram(mods, cloud20):
    torch, bev, _ = mods
    pts = torch.from_numpy(cloud20).cuda()                         # 1,062,920 points
    n = pts.shape[0]
    shape, vs, zo = synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE, synth.BEV_Z_OFFSET
    offs = np.array([0, n], dtype=np.int64)
    raw = bev.rasterize_frames(pts, offs, shape, vs, zo, want=("raw",))["raw"][0]
    g = torch.Generator(device="cuda").manual_seed(3)
    perm = torch.randperm(n, generator=g, device="cuda")
    raw_p = bev.rasterize_frames(pts[perm].contiguous(), offs, shape, vs, zo, want=("raw",))["raw"][0]
    assert torch.equal(raw, raw_p)                                  # order of the points is irrelevant
    # duplication doubles every count; a split cloud adds up (linearity of the histogram)
    both = torch.cat([pts, pts])
    raw2 = bev.rasterize_frames(both, np.array([0, 2 * n], dtype=np.int64), shape, vs, zo, want=("raw",))["raw"][0]
    assert torch.equal(raw2, 2 * raw)
    halves = bev.rasterize_frames(pts, np.array([0, n // 3, n], dtype=np.int64), shape, vs, zo, want=("raw",))["raw"]
    assert torch.equal(halves[0] + halves[1], raw)
    # checksum: the grid sums to the number of in-bounds points, computed independently in fp64
    m = torch.tensor([1.0 / v for v in vs], dtype=torch.float64, device="cuda")
    t = torch.tensor([shape[0] / 2, shape[1] / 2, shape[2] / 2 + zo / vs[2]], dtype=torch.float64, device="cuda")
    c = torch.trunc(pts[:, :3].double() * m + t)
    lim = torch.tensor(shape, dtype=torch.float64, device="cuda")
    inb = ((c >= 0) & (c < lim)).all(dim=1).sum().item()
    assert int(raw.sum().item()) == inb
    # normalised image: idempotent clip, u8 = rint(norm * 255)
    res = bev.rasterize_frames(pts, offs, shape, vs, zo, want=("norm", "u8"))
    norm, u8 = res["norm"][0], res["u8"][0]
    assert torch.equal(norm, torch.clamp(raw / 16.0, 0, 1))
    assert torch.equal(u8, torch.round(norm * 255).to(torch.uint8))


def test_voxelizer_permutation_keeps_the_voxel_set(mods, cloud11):
    """Below the caps the SET of voxels and their point counts do not depend on the point order; only
    the first-come order of the voxel list does.  Sortedness: re-sorting both results by cell gives the
    same lists."""
    torch, _, vg = mods
    pts = torch.from_numpy(cloud11).cuda()
    n = pts.shape[0]
    offs = np.array([0, n], dtype=np.int64)
    V = 120000                                                       # far above the occupied pillars of this cloud
    vs, rg = synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE
    g = torch.Generator(device="cuda").manual_seed(9)
    perm = torch.randperm(n, generator=g, device="cuda")
    outs = []
    for x in (pts, pts[perm].contiguous()):
        vox, co, num, vn = vg.voxelize_frames(x, offs, vs, rg, 64, V, zero_tail=False)
        k = int(vn[0])
        key = (co[0, :k, 0].long() * 400 + co[0, :k, 1].long()) * 400 + co[0, :k, 2].long()
        order = torch.argsort(key)
        outs.append((key[order], num[0, :k][order], k))
    assert outs[0][2] == outs[1][2]
    assert torch.equal(outs[0][0], outs[1][0])                       # same voxel set
    assert len(torch.unique(outs[0][0])) == outs[0][2]               # coordinates unique
    # counts are capped at max_points (64) identically: min(true count, 64) is order independent
    assert torch.equal(outs[0][1], outs[1][1])


def test_scatter_is_the_inverse_of_a_gather(mods):
    torch, _, _ = mods
    from lyft3d_b200 import pointpillars as pp
    g = torch.Generator(device="cuda").manual_seed(1)
    B, C, ny, nx, P = 3, 64, 400, 400, 30000
    cells = torch.stack([torch.randperm(ny * nx, generator=g, device="cuda")[:P] for _ in range(B)])
    coords = torch.zeros((B * P, 4), dtype=torch.int32, device="cuda")
    coords[:, 0] = torch.arange(B, device="cuda").repeat_interleave(P)
    coords[:, 2] = (cells.flatten() // nx).int()
    coords[:, 3] = (cells.flatten() % nx).int()
    feats = torch.randn((B * P, C), generator=g, device="cuda")
    canvas = pp.scatter_pillars(feats, coords, B, ny, nx)
    back = canvas[coords[:, 0].long(), :, coords[:, 2].long(), coords[:, 3].long()]
    assert torch.equal(back, feats)                                  # round trip
    assert int((canvas != 0).sum().item()) == int((feats != 0).sum().item())   # zeros elsewhere
