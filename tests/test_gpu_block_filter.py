"""GPU parity of the block-filtering voxelizer (lv_voxelize_filtered / lv_voxel_block_filter) against
oracle/voxel_oracle.py::points_to_voxel_filtered - bit-exact voxels, coordinates, counts and order.
The oracle restates spconv 1.x's rule from memory of its published source: PARITY UNPINNED at that
boundary."""
import numpy as np
import pytest

from lyft3d_b200 import synth

pytestmark = pytest.mark.gpu

FHD_VS = (0.05, 0.05, 0.2)
FHD_RANGE = (-50, -50, -5, 50, 50, 3)


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


def check(res, ref):
    v, c, n = ref[:3]
    assert res["voxel_num"] == v.shape[0]
    assert np.array_equal(res["coordinates"], c)
    assert np.array_equal(res["num_points_per_voxel"], n)
    assert np.array_equal(res["voxels"].view(np.uint32), v.view(np.uint32))


@pytest.mark.parametrize("bf,bs,lo,hi", [(1, 8, 0.2, 2.0), (8, 3, 0.1, 2.0), (4, 1, 0.3, None), (2, 5, 0.0, 1.0)])
def test_fhd_config_single_sweep(vg, vo, bf, bs, lo, hi):
    pts = synth.c5_frame(3)
    gen = vg.VoxelGeneratorV2(FHD_VS, FHD_RANGE, 3, max_voxels=40000, block_filtering=True, block_factor=bf,
                              block_size=bs, height_threshold=lo, height_high_threshold=hi)
    ref = vo.points_to_voxel_filtered(pts, FHD_VS, FHD_RANGE, 3, 40000, bf, bs, lo, hi)
    check(gen.generate(pts, 40000), ref)
    check(gen.generate(pts, 40000), ref)        # workspace state is reusable
    padded = gen.generate_multi_gpu(pts, 40000)
    k = padded["voxel_num"]
    assert k == ref[0].shape[0] and padded["voxels"].shape[0] == 40000
    assert np.array_equal(padded["coordinates"][:k], ref[1]) and not padded["coordinates"][k:].any()
    assert not padded["voxels"][k:].any() and not padded["num_points_per_voxel"][k:].any()


def test_multisweep_cloud_hits_max_voxels(vg, vo, cloud11):
    # the cap is reached before the filter runs: spconv filters the first max_voxels voxels
    gen = vg.VoxelGeneratorV2(FHD_VS, FHD_RANGE, 5, max_voxels=20000, block_filtering=True, block_factor=1,
                              block_size=8, height_threshold=0.2)
    ref = vo.points_to_voxel_filtered(cloud11, FHD_VS, FHD_RANGE, 5, 20000, 1, 8, 0.2, 2.0)
    assert ref[3].shape[0] == 20000
    check(gen.generate(cloud11, 20000), ref)


def test_pillar_grid_and_five_features(vg, vo):
    rng = np.random.default_rng(11)
    pts = np.concatenate([synth.c5_frame(5), rng.uniform(-1, 1, (53146, 1)).astype(np.float32)], axis=1)
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 7, max_voxels=30000,
                              block_filtering=True, block_factor=2, block_size=3, height_threshold=0.4,
                              height_high_threshold=3.0)
    ref = vo.points_to_voxel_filtered(pts, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 7, 30000, 2, 3, 0.4, 3.0)
    check(gen.generate(pts, 30000), ref)


def test_batched_frames_on_device_with_mask(vg, vo):
    import ctypes
    import torch
    from lyft3d_b200 import _native as nat
    frames = [synth.c5_frame(i)[: 20000 + 3000 * i] for i in range(5)] + [np.zeros((0, 4), np.float32)]
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    F, V, T = len(frames), 12000, 4
    voxels, coords, num, vnum = vg.voxelize_frames(pts, offs, FHD_VS, FHD_RANGE, T, V, zero_tail=True,
                                                   block_filter=(1, 8, 0.2, 2.0))
    for f, fr in enumerate(frames):
        ref = vo.points_to_voxel_filtered(fr, FHD_VS, FHD_RANGE, T, V, 1, 8, 0.2, 2.0)
        k = int(vnum[f])
        assert k == ref[0].shape[0]
        assert np.array_equal(coords[f, :k].cpu().numpy(), ref[1])
        assert np.array_equal(num[f, :k].cpu().numpy(), ref[2])
        assert np.array_equal(voxels[f, :k].cpu().numpy().view(np.uint32), ref[0].view(np.uint32))
        assert not voxels[f, k:].any() and not coords[f, k:].any() and not num[f, k:].any()
    # the stand-alone filter on ordinary voxelizer output, with the keep mask
    v0, c0, n0, k0 = vg.voxelize_frames(pts, offs, FHD_VS, FHD_RANGE, T, V, zero_tail=False)
    cfg = vg._make_config(FHD_VS, FHD_RANGE, T, V, 4, "continue", False)
    flt = vg._make_filter((1, 8, 0.2, 2.0))
    ov, oc, on = torch.empty_like(v0), torch.empty_like(c0), torch.empty_like(n0)
    ok, mask = torch.empty_like(k0), torch.empty_like(n0)
    h = nat.get_handle(0)
    nat.check(nat.load().lv_voxel_block_filter(h.ptr, ctypes.byref(cfg), ctypes.byref(flt), F, v0.data_ptr(),
                                               c0.data_ptr(), n0.data_ptr(), k0.data_ptr(), ov.data_ptr(),
                                               oc.data_ptr(), on.data_ptr(), ok.data_ptr(), mask.data_ptr(),
                                               nat.current_stream_ptr(0)))
    assert torch.equal(ok, vnum)
    for f, fr in enumerate(frames):
        ref = vo.points_to_voxel_filtered(fr, FHD_VS, FHD_RANGE, T, V, 1, 8, 0.2, 2.0)
        ku = int(k0[f])
        assert np.array_equal(mask[f, :ku].cpu().numpy().astype(bool), ref[3])
        assert not mask[f, ku:].any()


def test_errors(vg):
    with pytest.raises(AssertionError):
        vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 5, block_filtering=True, block_factor=7)
    with pytest.raises(ValueError):
        vg.VoxelGeneratorV2(FHD_VS, FHD_RANGE, 5, block_filtering=True, overflow="break")
