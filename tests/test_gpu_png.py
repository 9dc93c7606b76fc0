"""GPU parity of the device PNG encoder (lv_png_encode): byte-identical to oracle/png_oracle.py and
decode-exact under cv2 (the reference writes with cv2.imwrite, generating_train_bev.py:215,224, and reads
with cv2.imread, dataset.py:83-90)."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import bev_oracle as bo, png_oracle as po

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def bev():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import bev as b
    return b


def _check(files, imgs):
    for png, img in zip(files, imgs):
        assert png == po.encode_png(img), "bytes differ from the oracle's stream"
        dec = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec.shape == img.shape and np.array_equal(dec, img)


def test_bev_frames_batch(bev):
    import torch
    F = 6
    frames = [synth.c5_frame(40 + f) for f in range(F)]
    rows = torch.from_numpy(np.concatenate(frames)).cuda()
    offs = np.arange(F + 1, dtype=np.int64) * frames[0].shape[0]
    u8 = bev.rasterize_frames(rows, offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET, want=("u8",))["u8"]
    out, sizes = bev.encode_png_frames(u8)
    sizes = sizes.cpu().numpy()
    out = out.cpu().numpy()
    imgs = u8.cpu().numpy()
    _check([out[f, :sizes[f]].tobytes() for f in range(F)], imgs)
    assert int(sizes.sum()) * 10 < imgs.size            # what crosses PCIe instead of the dense images
    # the oracle chain end to end: points -> reference BEV closures' restatement -> u8 -> PNG
    for f in (0, F - 1):
        raw = bo.create_voxel_pointcloud(np.ascontiguousarray(frames[f].T), synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
        assert out[f, :sizes[f]].tobytes() == po.encode_png(bo.quantize_u8(bo.normalize_voxel_intensities(raw)))


@pytest.mark.parametrize("shape,kind", [((336, 336, 3), "dense"), ((1024, 1024, 3), "sparse"), ((1024, 1024), "grey"),
                                        ((7, 5, 3), "dense"), ((1, 1), "dense"), ((9, 1365, 3), "runs"), ((13, 4095), "runs"),
                                        ((64, 600), "lengths"), ((336, 336), "grey")])
def test_shapes_and_contents(bev, shape, kind):
    rng = np.random.default_rng(sum(shape) * 31 + len(kind))
    if kind == "dense":
        img = rng.integers(0, 256, shape, dtype=np.uint8)
    elif kind == "sparse":
        img = np.where(rng.random(shape) > 0.98, rng.integers(1, 256, shape), 0).astype(np.uint8)
    elif kind == "grey":
        img = (rng.random(shape) > 0.7).astype(np.uint8) * rng.integers(1, 10, shape).astype(np.uint8)
    elif kind == "runs":
        img = np.repeat(rng.integers(0, 256, shape[:1] + (1,) + shape[2:], dtype=np.uint8), shape[1], axis=1)
        img[:, ::97] = 255
    else:
        img = np.zeros(shape, np.uint8)
        for r in range(shape[0]):
            img[r, :min(r * 9 + (r % 4), shape[1])] = 5 + r
    _check([bev.imencode_png(img)], [img])


def test_cuda_tensor_single_and_target_images(bev):
    import torch
    corners, cls = synth.box_scene(3, 50)
    tgt = bev.rasterize_targets(corners, cls + 1, np.array([0, 50], dtype=np.int64), synth.BEV_SHAPE,
                                synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)[0]          # (336,336) uint8 class ids
    _check([bev.imencode_png(torch.from_numpy(tgt).cuda())], [tgt])


def test_imwrite_reads_back_with_cv2(bev, tmp_path):
    img = np.where(np.random.default_rng(1).random((120, 90, 3)) > 0.9, 200, 0).astype(np.uint8)
    path = str(tmp_path / "x_input.png")
    assert bev.imwrite(path, img) is True
    assert np.array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED), img)
    ref = str(tmp_path / "ref.png")
    cv2.imwrite(ref, img)
    assert np.array_equal(cv2.imread(ref, cv2.IMREAD_UNCHANGED), cv2.imread(path, cv2.IMREAD_UNCHANGED))


def test_too_small_slot_reports_the_needed_size(bev):
    import torch
    img = np.random.default_rng(2).integers(0, 256, (2, 64, 64, 3), dtype=np.uint8)
    t = torch.from_numpy(img).cuda()
    out = torch.zeros((2, 1024), dtype=torch.uint8, device="cuda")
    guard = out.clone()
    _, sizes = bev.encode_png_frames(t, out=out[:, :512].contiguous(), stride=512)
    need = [len(po.encode_png(img[f])) for f in range(2)]
    assert sizes.cpu().tolist() == [-n for n in need]
    assert bool((out == guard).all())


def test_files_written_straight_into_mapped_host_memory(bev):
    """`out` and `sizes` in mapped pinned host memory: the kernels write the files over PCIe themselves."""
    import torch
    from lyft3d_b200.engine import MappedBuffer
    rng = np.random.default_rng(5)
    img = np.where(rng.random((4, 200, 160, 3)) > 0.95, rng.integers(1, 256, (4, 200, 160, 3)), 0).astype(np.uint8)
    stride = bev.png_slot_bytes(200, 160, 3)
    mo, ms = MappedBuffer((4, stride), torch.uint8), MappedBuffer((4,), torch.int32)
    mo.tensor.zero_()
    bev.encode_png_frames(torch.from_numpy(img).cuda(), out=mo.tensor, sizes=ms.tensor, stride=stride)
    torch.cuda.synchronize()
    sizes = ms.tensor.numpy()
    _check([mo.tensor[f, :sizes[f]].numpy().tobytes() for f in range(4)], img)
    assert not mo.tensor[0, sizes[0]:].any()             # nothing beyond the file is touched


def test_write_training_pngs_equals_the_reference_loop_body(tmp_path):
    """bev.write_training_pngs = the per-sample body of prepare_training_data_for_scene
    (generating_train_bev.py:208-224) for a list of samples: the files it writes decode (cv2.imread, what
    dataset.py:83-90 does) to exactly what the oracle chain from_file -> transform -> create_voxel_pointcloud ->
    normalize -> round*255 and draw_boxes produce."""
    import cv2
    from lyft3d_b200 import bev
    from oracle import bev_oracle as bo
    from oracle import draw_oracle
    raw5 = synth.load_fixture_raw()
    sweeps = [raw5[: 30000 + 7000 * i] for i in range(3)] + [raw5[:0]]
    tms = [synth.sweep_transform(3 * i + 1) for i in range(4)]
    scenes = [synth.box_scene(7100 + i, 25 + 10 * i) for i in range(3)] + [(np.zeros((0, 3, 4)), np.zeros((0,), np.int32))]
    corners = np.concatenate([c for c, _ in scenes])
    colors = np.concatenate([k + 1 for _, k in scenes]).astype(np.int32)
    box_offs = np.concatenate([[0], np.cumsum([c.shape[0] for c, _ in scenes])]).astype(np.int64)
    tokens = ["tok%d" % i for i in range(4)]
    paths = bev.write_training_pngs(sweeps, tms, corners, colors, box_offs, tokens, str(tmp_path), synth.BEV_SHAPE,
                                    synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    assert [os.path.basename(a) for a, _ in paths] == ["tok%d_input.png" % i for i in range(4)]
    for i, (a, b) in enumerate(paths):
        cloud = bo.sensor_to_car(np.ascontiguousarray(sweeps[i][:, :4].T), tms[i])
        ref = bo.quantize_u8(bo.normalize_voxel_intensities(
            bo.create_voxel_pointcloud(cloud, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)))
        got = cv2.imread(a, cv2.IMREAD_UNCHANGED)
        assert got.shape == ref.shape and np.array_equal(got, ref), i
        tgt = np.zeros(synth.BEV_SHAPE, dtype=np.float32)
        draw_oracle.draw_boxes(tgt, synth.BEV_VOXEL_SIZE, list(scenes[i][0]), list(scenes[i][1] + 1), synth.BEV_Z_OFFSET)
        got_t = cv2.imread(b, cv2.IMREAD_UNCHANGED)
        assert got_t.shape == (336, 336) and np.array_equal(got_t, tgt[:, :, 0].astype(np.uint8)), i
