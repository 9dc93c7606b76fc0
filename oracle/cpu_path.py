"""The reference's CPU path for one frame, assembled from the oracle restatements.
TEST INFRASTRUCTURE / bench.py CPU-baseline leg only (see oracle/__init__.py).

Per frame, what the reference does on the host:
  BEV     create_voxel_pointcloud -> normalize_voxel_intensities -> round*255 u8
          (numpy, generating-dataset/generating_train_bev.py:210-213)
  pillar  VoxelGeneratorV2.generate (single-thread native loop -> the C restatement),
          PillarFeatureNet decoration and PointPillarsScatter (numpy restatement of the
          torch ops of second/second/pytorch/models/pointpillars.py:203-231,444-476)
"""
import os
import time

import numpy as np

from . import bev_oracle, pillar_oracle, voxel_oracle

_STATE = {}


def _voxel_oracle(cfg):
    key = (tuple(cfg["voxel_size"]), tuple(cfg["pc_range"]), cfg["max_points"], cfg["max_voxels"])
    if key not in _STATE:
        _STATE[key] = voxel_oracle.VoxelOracle(cfg["voxel_size"], cfg["pc_range"], cfg["max_points"],
                                               cfg["max_voxels"])
    return _STATE[key]


def process_frame(points_nx4, cfg, features=None):
    """Runs both paths on one (N,4) float32 frame; returns (u8 image, canvas, voxel_num)."""
    bev = bev_oracle.create_voxel_pointcloud(points_nx4.T, cfg["bev_shape"], cfg["bev_voxel_size"],
                                             cfg["bev_z_offset"])
    norm = bev_oracle.normalize_voxel_intensities(bev)
    u8 = bev_oracle.quantize_u8(norm)
    orc = _voxel_oracle(cfg)
    v, c, n = orc.generate(points_nx4, overflow=cfg.get("overflow", "continue"))
    coors = pillar_oracle.merge_batch_coords([c])
    dec = pillar_oracle.decorate(v, n, coors, cfg["voxel_size"], cfg["pc_range"])
    if features is None:
        # stand-in for the PFN output (excluded from the scope): first 64
        # decorated values of every pillar - same bytes moved as real features
        feats = np.ascontiguousarray(dec.reshape(dec.shape[0], -1)[:, :cfg["channels"]])
    else:
        feats = features[:v.shape[0]]
    canvas = pillar_oracle.scatter(feats, coors, 1, cfg["canvas"][0], cfg["canvas"][1])
    return u8, canvas, v.shape[0]


def _worker(args):
    frame_ids, cfg = args
    from lyft3d_b200 import synth
    base = synth.fixture_points_nx4()
    frames = [synth.c5_frame(int(f), base) for f in frame_ids]
    t0 = time.perf_counter()
    npts = 0
    for fr in frames:
        process_frame(fr, cfg)
        npts += fr.shape[0]
    return npts, time.perf_counter() - t0


def run_pool(frame_ids, cfg, workers):
    """Frames over a process pool (the reference's own parallelism: Pool.imap_unordered over
    scenes / DataLoader workers).  Returns (points, wall seconds of the
    slowest worker) - frame generation is outside the timed region."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")  # generating_train_bev.py:10
    if workers <= 1:
        n, t = _worker((list(frame_ids), cfg))
        return n, t
    import multiprocessing as mp
    chunks = [list(frame_ids[w::workers]) for w in range(workers)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_worker, [(c, cfg) for c in chunks])
    return sum(r[0] for r in res), max(r[1] for r in res)
