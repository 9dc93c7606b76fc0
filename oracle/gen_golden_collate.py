"""Generates tests/golden/ref_collate.npz by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden_collate        (needs /root/reference; CPU only)

TEST INFRASTRUCTURE.  The examples are what ``prep_pointcloud`` returns
(second/second/data/preprocess.py:299-325,346,365-395): "voxels" / "num_points" /
"coordinates" / "num_voxels" from the voxelizer (here: the oracle restatement,
oracle/voxel_oracle.py, on seeded frames of different sizes), plus stand-ins of every other
kind of entry the collate treats differently (anchors, labels, reg_targets, gt_boxes,
gt_names, metadata, metrics, calib).  They are collated by the reference's OWN
``merge_second_batch`` and ``merge_second_batch_multigpu`` and converted by its own
``example_convert_to_torch`` on the CPU device (oracle.ref_loader.load_collate_functions:
the function bodies are read from the reference files and executed unchanged).
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import ref_loader, voxel_oracle  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")

# a small pillar configuration (the C3 grid at 0.5 m, T = 12) keeps the fixture small
VOXEL_SIZE = (0.5, 0.5, 20.0)
PC_RANGE = (-50.0, -50.0, -10.0, 50.0, 50.0, 10.0)
MAX_POINTS = 12
MAX_VOXELS = 1500          # frame 2 overflows it
FRAME_SIZES = (2500, 4000, 9000)


def frame_points(i):
    return np.ascontiguousarray(synth.c5_frame(40 + i)[:FRAME_SIZES[i]])


def make_examples(multigpu):
    """One example per frame, as prep_pointcloud builds it (voxels from the oracle)."""
    out = []
    for i in range(len(FRAME_SIZES)):
        pts = frame_points(i)
        rng = np.random.default_rng(900 + i)
        v, c, n = voxel_oracle.points_to_voxel(pts, VOXEL_SIZE, PC_RANGE, MAX_POINTS, MAX_VOXELS)
        k = v.shape[0]
        if multigpu:   # generate_multi_gpu: padded to max_voxels (preprocess.py:311-317)
            pv = np.zeros((MAX_VOXELS,) + v.shape[1:], np.float32); pv[:k] = v
            pc = np.zeros((MAX_VOXELS, 3), np.int32); pc[:k] = c
            pn = np.zeros((MAX_VOXELS,), np.int32); pn[:k] = n
            v, c, n = pv, pc, pn
        n_gt = 2 + i
        ex = {"voxels": v, "num_points": n, "coordinates": c, "num_voxels": np.array([k], dtype=np.int64),
              "metrics": {"voxel_gene_time": 0.0, "prep_time": float(i)},
              "calib": {"rect": np.eye(4, dtype=np.float32) * (i + 1), "Trv2c": rng.random((4, 4)).astype(np.float32)},
              "anchors": rng.random((50, 7)).astype(np.float32),
              "labels": rng.integers(-1, 3, size=(50,)).astype(np.int32),
              "reg_targets": rng.random((50, 7)).astype(np.float32),
              "importance": rng.random((50,)).astype(np.float32),
              "gt_names": np.array(["car"] * n_gt), "gt_boxes": rng.random((n_gt, 7)).astype(np.float32),
              "metadata": {"token": "sample%d" % i}}
        out.append(ex)
    return out


def main():
    assert ref_loader.available(), "needs /root/reference"
    import torch
    fn = ref_loader.load_collate_functions()
    gold = {}
    for tag, multigpu in (("single", False), ("multi", True)):
        merged = fn["merge_second_batch_multigpu" if multigpu else "merge_second_batch"](make_examples(multigpu))
        conv = fn["example_convert_to_torch"](merged, torch.float32, torch.device("cpu"))
        for k, v in merged.items():
            if isinstance(v, np.ndarray) and v.dtype.kind in "fiu":
                gold["%s.merged.%s" % (tag, k)] = v
            elif k == "calib":
                for k1, v1 in v.items():
                    gold["%s.merged.calib.%s" % (tag, k1)] = v1
        for k, v in conv.items():
            if isinstance(v, torch.Tensor):
                a = v.numpy()
                m = merged.get(k)
                same = isinstance(m, np.ndarray) and m.shape == a.shape and m.dtype == a.dtype and np.array_equal(m, a)
                # large arrays the conversion leaves unchanged are stored once (under merged.*)
                gold["%s.torch.%s" % (tag, k)] = np.array("== merged") if same and a.nbytes > 65536 else a
                gold["%s.torch_dtype.%s" % (tag, k)] = np.array(str(v.dtype))
        gold["%s.keys" % tag] = np.array(sorted(merged.keys()))
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "ref_collate.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes;", len(gold), "arrays")


if __name__ == "__main__":
    main()
