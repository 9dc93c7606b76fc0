"""Batch assembly drop-ins that ship RAW POINTS through the DataLoader.

Reference boundary (all three are replaced name for name):

    merge_second_batch(batch_list)            second/second/data/preprocess.py:21-55
    merge_second_batch_multigpu(batch_list)   second/second/data/preprocess.py:57-88
    example_convert_to_torch(example, dtype, device)   second/second/pytorch/train.py:34-62

In the reference every DataLoader worker voxelizes its sample on a CPU core
(``prep_pointcloud`` -> ``voxel_generator.generate``, preprocess.py:299-317), pickles the
``(V, T, C)`` array through the worker queue, ``merge_second_batch`` concatenates the
samples and ``example_convert_to_torch`` copies the batch to the GPU.  Here the worker is
handed a ``DeferredVoxelGenerator``: its ``generate`` does no work, the "voxels" entry of
the example is the sample's raw ``(N, C)`` points (a ``RawPoints`` array - 16 bytes per
point instead of T*C*4 bytes per voxel), ``merge_second_batch`` lays the samples of the
batch back to back with their offsets, and ``example_convert_to_torch`` copies the points
once and runs the batched voxelizer ONCE for the whole batch on the training process's
GPU (``lv_voxelize_concat`` - the device-side ``merge_second_batch`` - or ``lv_voxelize``
for the padded multi-GPU layout).  What comes out - "voxels", "num_points", "coordinates"
(with the batch column), "num_voxels" - has exactly the reference's shapes, dtypes and
values.  Every other key of an example is treated as the reference treats it.

Examples that already hold voxels (a plain ``VoxelGeneratorV2`` in the workers) pass
through unchanged, so the three functions can replace the reference's unconditionally.

Precondition of the deferred path: nothing between ``generate`` and the collate may read
the voxel coordinates on the host.  ``prep_pointcloud`` does so only for the anchor mask
(``anchor_area_threshold >= 0``, preprocess.py:348-359), which every Lyft / nuScenes config
of the reference disables (``anchor_area_threshold: -1``).
"""
from collections import defaultdict

import numpy as np

from .voxel_generator import VoxelGeneratorV2, voxelize_concat_frames, voxelize_frames

POINT_OFFSETS_KEY = "point_offsets"


class RawPoints(np.ndarray):
    """Marker type: an ``(N, C)`` float32 array of raw points standing in for the voxels of a
    sample (or, after the collate, of a batch).  Survives pickling through the worker queue."""


class DeferredVoxelGenerator(VoxelGeneratorV2):
    """``VoxelGeneratorV2`` for DataLoader workers that defers the work to the collate.

    Same constructor, same properties (``voxel_size``, ``point_cloud_range``, ``grid_size``,
    ``max_num_points_per_voxel`` - read by ``prep_pointcloud`` and ``VoxelNet.__init__``).
    ``generate`` / ``generate_multi_gpu`` return the result dict of the reference with
    "voxels" = the points as ``RawPoints`` and empty "coordinates" / "num_points_per_voxel".
    """

    def _defer(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        if pts.ndim != 2:
            raise ValueError("points must be (N, C)")
        return {"voxels": pts.view(RawPoints), "coordinates": np.zeros((0, 3), dtype=np.int32),
                "num_points_per_voxel": np.zeros((0,), dtype=np.int32), "voxel_num": 0}

    def generate(self, points, max_voxels=None):
        return self._defer(points)

    def generate_multi_gpu(self, points, max_voxels=None):
        return self._defer(points)


def _is_deferred(elems):
    return len(elems) > 0 and all(isinstance(e, RawPoints) for e in elems)


def _merge_calib(elems):
    out = {}
    for elem in elems:
        for k1, v1 in elem.items():
            out.setdefault(k1, []).append(v1)
    return {k1: np.stack(v1, axis=0) for k1, v1 in out.items()}


def _pad_batch_column(elems):
    # preprocess.py:44-50: the sample index in front of (z, y, x)
    return [np.pad(c, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, c in enumerate(elems)]


def _merge(batch_list, multigpu):
    merged = defaultdict(list)
    for example in batch_list:
        for k, v in example.items():
            merged[k].append(v)
    deferred = "voxels" in merged and _is_deferred(merged["voxels"])
    ret = {}
    for key, elems in merged.items():
        if deferred and key == "voxels":
            sizes = [e.shape[0] for e in elems]
            ret[key] = np.concatenate([np.asarray(e) for e in elems], axis=0).view(RawPoints)
            ret[POINT_OFFSETS_KEY] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        elif deferred and key == "num_points":      # placeholders; their rank tells the layout apart
            ret[key] = np.zeros((len(elems), 0) if multigpu else (0,), dtype=np.int32)
        elif deferred and key == "coordinates":
            ret[key] = np.zeros((len(elems), 0, 4) if multigpu else (0, 4), dtype=np.int32)
        elif key == "metadata":
            ret[key] = elems
        elif key == "calib":
            ret[key] = _merge_calib(elems)
        elif key == "coordinates":
            coors = _pad_batch_column(elems)
            ret[key] = np.stack(coors, axis=0) if multigpu else np.concatenate(coors, axis=0)
        elif multigpu:
            if key in ("gt_names", "gt_classes", "gt_boxes"):
                continue                                   # preprocess.py:83-84
            ret[key] = np.stack(elems, axis=0)
        elif key in ("voxels", "num_points", "num_gt", "voxel_labels", "gt_names", "gt_classes", "gt_boxes"):
            ret[key] = np.concatenate(elems, axis=0)
        elif key == "metrics":
            ret[key] = elems
        else:
            ret[key] = np.stack(elems, axis=0)
    return ret


def merge_second_batch(batch_list):
    """``collate_fn`` of the single-GPU DataLoader (preprocess.py:21-55)."""
    return _merge(batch_list, multigpu=False)


def merge_second_batch_multigpu(batch_list):
    """``collate_fn`` of the multi-GPU DataLoader (preprocess.py:57-88): padded per-sample layout."""
    return _merge(batch_list, multigpu=True)


def _voxelize_batch(example, device, voxel_generator, max_voxels, multigpu):
    """points of the whole batch -> the reference's voxel entries, on `device`."""
    import torch
    if voxel_generator is None:
        raise ValueError("the batch carries raw points: example_convert_to_torch needs the voxel_generator "
                         "(the object prep_pointcloud was given) to voxelize them")
    gen = voxel_generator
    mv = int(gen._max_voxels if max_voxels is None else max_voxels)
    offs = np.ascontiguousarray(example[POINT_OFFSETS_KEY], dtype=np.int64)
    B = offs.shape[0] - 1
    pts = torch.from_numpy(np.asarray(example["voxels"])).to(device)
    if pts.shape[0] != offs[-1]:
        raise ValueError("point_offsets do not match the points of the batch")
    with torch.cuda.device(pts.device):
        if multigpu or gen._block_filter is not None:
            voxels, coords3, num, vnum = voxelize_frames(pts, offs, gen._voxel_size, gen._point_cloud_range,
                                                         gen._max_num_points, mv, overflow=gen._overflow,
                                                         zero_tail=True, block_filter=gen._block_filter)
            coords = torch.nn.functional.pad(coords3, (1, 0))
            coords[:, :, 0] = torch.arange(B, dtype=torch.int32, device=pts.device)[:, None]
            counts = vnum.cpu().numpy().astype(np.int64)
            if not multigpu:
                from .voxel_generator import unpad_multigpu_batch
                voxels, num, coords = unpad_multigpu_batch(voxels, num, coords, vnum)
        else:
            voxels, coords, num, vnum = voxelize_concat_frames(pts, offs, gen._voxel_size, gen._point_cloud_range,
                                                               gen._max_num_points, mv, overflow=gen._overflow)
            counts = vnum.cpu().numpy().astype(np.int64)
    return voxels, num, coords, counts.reshape(B, 1)


def example_convert_to_torch(example, dtype=None, device=None, voxel_generator=None, max_voxels=None) -> dict:
    """train.py:34-62 with one addition: a batch that carries raw points (``DeferredVoxelGenerator``
    in the workers) is voxelized here, once, on `device`.  `voxel_generator` / `max_voxels` are the
    objects ``prep_pointcloud`` was given (dataset_builder.py:75-79; ``max_number_of_voxels``).  A batch
    collated by ``merge_second_batch_multigpu`` comes out in the padded layout ``VoxelNet.forward`` un-pads
    (voxelnet.py:346-358): "voxels" (B, V, T, C), "num_points" (B, V), "coordinates" (B, V, 4)."""
    import torch
    dtype = dtype or torch.float32
    device = device or torch.device("cuda:0")
    example = dict(example)
    if isinstance(example.get("voxels"), RawPoints):
        multigpu = np.ndim(example["coordinates"]) == 3     # collated by merge_second_batch_multigpu
        voxels, num, coords, counts = _voxelize_batch(example, device, voxel_generator, max_voxels, multigpu)
        example.pop(POINT_OFFSETS_KEY)
        example["voxels"], example["num_points"], example["coordinates"] = voxels, num, coords
        example["num_voxels"] = counts
    example_torch = {}
    float_names = ["voxels", "anchors", "reg_targets", "reg_weights", "bev_map", "importance"]
    for k, v in example.items():
        if k in float_names:
            if isinstance(v, torch.Tensor):
                example_torch[k] = v.to(device=device, dtype=dtype)
            else:
                example_torch[k] = torch.tensor(v, dtype=torch.float32, device=device).to(dtype)
        elif k in ["coordinates", "labels", "num_points"]:
            if isinstance(v, torch.Tensor):
                example_torch[k] = v.to(device=device, dtype=torch.int32)
            else:
                example_torch[k] = torch.tensor(v, dtype=torch.int32, device=device)
        elif k in ["anchors_mask"]:
            example_torch[k] = torch.tensor(v, dtype=torch.uint8, device=device)
        elif k == "calib":
            example_torch[k] = {k1: torch.tensor(v1, dtype=dtype, device=device).to(dtype) for k1, v1 in v.items()}
        elif k == "num_voxels":
            example_torch[k] = torch.tensor(v)
        else:
            example_torch[k] = v
    return example_torch


__all__ = ["RawPoints", "DeferredVoxelGenerator", "merge_second_batch", "merge_second_batch_multigpu",
           "example_convert_to_torch", "POINT_OFFSETS_KEY"]
