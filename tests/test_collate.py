"""Batch assembly drop-ins (lyft3d_b200.collate) against the reference's OWN
merge_second_batch / merge_second_batch_multigpu / example_convert_to_torch
(second/second/data/preprocess.py:21-88, second/second/pytorch/train.py:34-62):
tests/golden/ref_collate.npz holds their outputs (oracle/gen_golden_collate.py executed the
reference's function bodies); when /root/reference exists they are re-run live.
CPU part: examples that already hold voxels; the deferred (raw points) layout and its pickling."""
import os
import pickle

import numpy as np
import pytest

from lyft3d_b200 import collate
from oracle import gen_golden_collate as gg
from oracle import ref_loader


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_collate.npz"))


def _check_merged(tag, merged, gold):
    assert sorted(merged.keys()) == list(gold["%s.keys" % tag])
    n = 0
    for name in gold.files:
        if not name.startswith(tag + ".merged."):
            continue
        key = name[len(tag) + 8:]
        got = merged["calib"][key[6:]] if key.startswith("calib.") else merged[key]
        ref = gold[name]
        assert got.dtype == ref.dtype and got.shape == ref.shape, key
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), key
        n += 1
    assert n >= 9
    return n


@pytest.mark.parametrize("tag,multigpu", [("single", False), ("multi", True)])
def test_merge_equals_reference_output(gold, tag, multigpu):
    examples = gg.make_examples(multigpu)
    fn = collate.merge_second_batch_multigpu if multigpu else collate.merge_second_batch
    merged = fn(examples)
    _check_merged(tag, merged, gold)
    assert merged["metadata"] == [{"token": "sample%d" % i} for i in range(3)]
    if not multigpu:
        assert merged["metrics"] == [e["metrics"] for e in examples]
        assert list(merged["gt_names"]) == ["car"] * 9


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
@pytest.mark.parametrize("multigpu", [False, True])
def test_merge_equals_reference_live(multigpu):
    fn = ref_loader.load_collate_functions()
    examples = gg.make_examples(multigpu)
    ref = fn["merge_second_batch_multigpu" if multigpu else "merge_second_batch"](examples)
    got = (collate.merge_second_batch_multigpu if multigpu else collate.merge_second_batch)(examples)
    assert list(got.keys()) == list(ref.keys())
    for k, v in ref.items():
        if isinstance(v, np.ndarray):
            assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k
        elif k == "calib":
            assert all(np.array_equal(got[k][k1], v1) for k1, v1 in v.items())
        else:
            assert got[k] == v, k


def test_convert_to_torch_cpu_equals_reference_output(gold):
    """Voxel-holding examples pass through example_convert_to_torch as in the reference (CPU device)."""
    import torch
    for tag, multigpu in (("single", False), ("multi", True)):
        merged = (collate.merge_second_batch_multigpu if multigpu else collate.merge_second_batch)(gg.make_examples(multigpu))
        conv = collate.example_convert_to_torch(merged, torch.float32, torch.device("cpu"))
        n = 0
        for name in gold.files:
            if not name.startswith(tag + ".torch."):
                continue
            key = name[len(tag) + 7:]
            ref = gold[name]
            if ref.dtype.kind == "U":      # "== merged": stored once under merged.*
                ref = gold["%s.merged.%s" % (tag, key)]
            got = conv[key]
            assert str(got.dtype) == str(gold["%s.torch_dtype.%s" % (tag, key)]), key
            assert got.shape == ref.shape and np.array_equal(got.numpy(), ref), key
            n += 1
        assert n >= 8
        assert conv["metadata"] == merged["metadata"]


def test_deferred_generator_and_layout():
    gen = collate.DeferredVoxelGenerator(gg.VOXEL_SIZE, gg.PC_RANGE, gg.MAX_POINTS, max_voxels=gg.MAX_VOXELS)
    assert np.array_equal(gen.grid_size, [200, 200, 1]) and gen.max_num_points_per_voxel == gg.MAX_POINTS
    gen2 = pickle.loads(pickle.dumps(gen))               # captured in functools.partial for the workers
    assert np.array_equal(gen2.voxel_size, gen.voxel_size)
    examples = []
    for i in range(3):
        pts = gg.frame_points(i)
        res = gen2.generate(pts, gg.MAX_VOXELS)
        assert isinstance(res["voxels"], collate.RawPoints) and res["voxels"].shape == pts.shape
        # what prep_pointcloud builds from it (preprocess.py:305-325)
        ex = {"voxels": res["voxels"], "num_points": res["num_points_per_voxel"], "coordinates": res["coordinates"],
              "num_voxels": np.array([res["voxels"].shape[0]], dtype=np.int64), "anchors": np.zeros((5, 7), np.float32),
              "metadata": {"token": i}}
        examples.append(pickle.loads(pickle.dumps(ex)))  # the worker queue
        assert isinstance(examples[-1]["voxels"], collate.RawPoints)
    for fn, nd in ((collate.merge_second_batch, 2), (collate.merge_second_batch_multigpu, 3)):
        b = fn(examples)
        assert isinstance(b["voxels"], collate.RawPoints) and b["voxels"].shape == (sum(gg.FRAME_SIZES), 4)
        assert np.array_equal(b[collate.POINT_OFFSETS_KEY], np.concatenate([[0], np.cumsum(gg.FRAME_SIZES)]))
        assert b["coordinates"].ndim == nd and b["anchors"].shape == (3, 5, 7)
        for i in range(3):
            lo, hi = b[collate.POINT_OFFSETS_KEY][i:i + 2]
            assert np.array_equal(np.asarray(b["voxels"][lo:hi]), gg.frame_points(i))
    with pytest.raises(ValueError):
        collate.example_convert_to_torch(collate.merge_second_batch(examples), device="cpu")   # no generator given


@pytest.mark.gpu
@pytest.mark.parametrize("tag,multigpu", [("single", False), ("multi", True)])
def test_gpu_deferred_batch_equals_reference_output(gold, tag, multigpu):
    """Raw points through the collate, voxelized once per batch on the GPU == the reference's batch."""
    import torch
    gen = collate.DeferredVoxelGenerator(gg.VOXEL_SIZE, gg.PC_RANGE, gg.MAX_POINTS, max_voxels=gg.MAX_VOXELS)
    ref_examples = gg.make_examples(multigpu)
    examples = []
    for i, rex in enumerate(ref_examples):
        res = (gen.generate_multi_gpu if multigpu else gen.generate)(gg.frame_points(i), gg.MAX_VOXELS)
        ex = dict(rex)
        ex.update(voxels=res["voxels"], num_points=res["num_points_per_voxel"], coordinates=res["coordinates"],
                  num_voxels=np.array([res["voxels"].shape[0]], dtype=np.int64))
        examples.append(ex)
    merged = (collate.merge_second_batch_multigpu if multigpu else collate.merge_second_batch)(examples)
    conv = collate.example_convert_to_torch(merged, torch.float32, torch.device("cuda:0"), voxel_generator=gen,
                                            max_voxels=gg.MAX_VOXELS)
    assert collate.POINT_OFFSETS_KEY not in conv
    n = 0
    for name in gold.files:
        if not name.startswith(tag + ".torch."):
            continue
        key = name[len(tag) + 7:]
        ref = gold[name]
        if ref.dtype.kind == "U":
            ref = gold["%s.merged.%s" % (tag, key)]
        got = conv[key]
        assert str(got.dtype) == str(gold["%s.torch_dtype.%s" % (tag, key)]), key
        if key != "num_voxels":
            assert got.is_cuda, key
        assert tuple(got.shape) == ref.shape, key
        assert np.array_equal(got.cpu().numpy().view(np.uint8), ref.view(np.uint8)), key
        n += 1
    assert n >= 8
    assert int(conv["num_voxels"].sum()) == int(gold["single.merged.voxels"].shape[0])


@pytest.mark.gpu
def test_gpu_deferred_batch_half_and_block_filter():
    """dtype=half (apex O2, train.py:220-228) and a block-filtering generator take the deferred path too."""
    import torch
    from lyft3d_b200 import voxel_generator as vg
    frames = [gg.frame_points(i) for i in range(3)]
    for kw in ({}, dict(block_filtering=True, block_factor=1, block_size=8, height_threshold=0.2)):
        vs, rg, T, V = ((0.05, 0.05, 0.2), (-50, -50, -5, 50, 50, 3), 3, 4000) if kw else (gg.VOXEL_SIZE, gg.PC_RANGE, gg.MAX_POINTS, gg.MAX_VOXELS)
        gen = collate.DeferredVoxelGenerator(vs, rg, T, max_voxels=V, **kw)
        plain = vg.VoxelGeneratorV2(vs, rg, T, max_voxels=V, **kw)
        ex = [{"voxels": gen.generate(p)["voxels"], "num_points": np.zeros((0,), np.int32),
               "coordinates": np.zeros((0, 3), np.int32), "num_voxels": np.array([p.shape[0]])} for p in frames]
        conv = collate.example_convert_to_torch(collate.merge_second_batch(ex), torch.float16, torch.device("cuda:0"),
                                                voxel_generator=gen)
        refs = [plain.generate(p) for p in frames]
        ref = collate.merge_second_batch([{"voxels": r["voxels"], "num_points": r["num_points_per_voxel"],
                                           "coordinates": r["coordinates"],
                                           "num_voxels": np.array([r["voxels"].shape[0]], dtype=np.int64)} for r in refs])
        assert conv["voxels"].dtype == torch.float16
        assert np.array_equal(conv["voxels"].cpu().numpy(), ref["voxels"].astype(np.float16))
        assert np.array_equal(conv["coordinates"].cpu().numpy(), ref["coordinates"])
        assert np.array_equal(conv["num_points"].cpu().numpy(), ref["num_points"])
        assert np.array_equal(conv["num_voxels"].numpy(), ref["num_voxels"])
