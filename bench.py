#!/usr/bin/env python
"""bench.py - points/sec voxelized (BEV + pillar) on N B200s, with the roofline of the
dominant kernel and the reference's CPU path timed beside it.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[4]): synthetic Lyft-shaped sweeps of
53,146 points (the bundled sweep rotated about z + N(0,0.02) jitter).  One step = one
pass of BOTH paths over `frames_per_step` frames per GPU:
    BEV     336x336x3 @ 0.4x0.4x1.5 m -> normalised f32 grid + u8 image
    pillar  0.25 m pillars, max_points 60, max_voxels 30000 -> voxelize (device-side batch
            assembly) -> PillarFeatureNet decoration -> PointPillarsScatter into (F,64,400,400)
Frames are independent: every rank works on its own frames, no collective on the data
path (weak scaling: per-GPU work is fixed).  `value` counts each point once per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/sec voxelized (BEV + pillar)"
UNIT = "points/s"


def workload_cfg():
    from lyft3d_b200 import synth
    return dict(bev_shape=synth.BEV_SHAPE, bev_voxel_size=synth.BEV_VOXEL_SIZE, bev_z_offset=synth.BEV_Z_OFFSET,
                voxel_size=synth.PILLAR_VOXEL_SIZE, pc_range=synth.PILLAR_RANGE,
                max_points=synth.PILLAR_MAX_POINTS, max_voxels=synth.PILLAR_MAX_VOXELS,
                canvas=synth.PILLAR_CANVAS, channels=synth.PILLAR_FEATURES)


def config_json(frames_per_step, n_gpus, extra=None):
    """The workload both arms of the bench run (BASELINE.json configs[4]); identical for `--impl native` and
    `--impl reference` at the same --gpus / --frames-per-step (what differs between the arms - the execution, the
    bounded sample of the CPU arm - is reported beside it, not inside it)."""
    c = {"workload": "C5 throughput sweep: synthetic 53,146-point Lyft sweeps, BEV 336x336x3 (norm f32 + u8) "
                     "+ pillar path (voxelize 0.25 m/T=60/V=30000 -> decorate -> scatter 64x400x400)",
         "pillar_path": "fused voxelize+decorate" if not (extra or {}).get("unfused") else "separate voxelize, decorate",
         "frames_per_step_per_gpu": frames_per_step, "points_per_frame": 53146,
         "parallelism": "frames sharded f mod G, dp%d, no collective" % n_gpus,
         "frames": "synth.c5_frame(f): the seeded definition of configs[4] the parity tests use, frame f on rank f mod G",
         "l2": "inputs (pool of distinct frames) and outputs per step exceed the 126 MB L2"}
    return c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _gen_frames(ids):
    from lyft3d_b200 import synth
    return synth.c5_frames(ids)[0]


def make_frames_host(frame_ids, workers=1):
    """The seeded C5 frames `frame_ids` (synth.c5_frame: the definition the parity tests use - bundled sweep
    rotated about z by 2*pi*u_f + N(0, 0.02) jitter, default_rng(5000 + f)) as one (len*N, 4) float32 host array.
    With workers > 1 the frames are generated over a fork pool - call this BEFORE the first CUDA call."""
    import numpy as np
    frame_ids = list(frame_ids)
    if workers <= 1 or len(frame_ids) < 4 * workers:
        return _gen_frames(frame_ids)
    import multiprocessing as mp
    chunks = [frame_ids[i:i + 64] for i in range(0, len(frame_ids), 64)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_gen_frames, chunks)
    return np.concatenate(parts, axis=0)


def source_sha():
    """sha256 over the kernel sources whose DRAM traffic profiles/r02_traffic.json records."""
    import hashlib
    h = hashlib.sha256()
    for f in ("lv_pillar.cu", "lv_voxel.cu", "lv_bev.cu", "lv_decorate.cuh", "lv_common.cuh"):
        with open(os.path.join(ROOT, "lyft-3d-object-detection_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def time_other_configs(dev, peak, reps=10):
    """BASELINE.json configs[0..3] as single-cloud latencies (configs[1..3]: CUDA events, median of
    `reps`, inputs resident on the device; the headline workload is configs[4]), each with its roofline
    fraction = algorithmic bytes / time / the measured HBM peak.  Host-side oracle timings
    (`cpu_port_ms`) are filled in by the CPU leg, which runs in its own process."""
    import numpy as np
    import torch
    from lyft3d_b200 import bev, synth, voxel_generator as vg

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    def roof(alg_bytes, ms):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        return {"algorithmic_MB": round(alg_bytes / 1e6, 2), "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}

    out = {}
    # C1: the reference's own CPU-runnable case through the drop-in call, numpy in -> numpy out
    # (H2D, kernels, D2H and the synchronisation inside the call), wall clock
    import time as _time
    p4 = synth.fixture_points_4xn()
    for _ in range(3):
        bev.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    t0 = _time.perf_counter()
    for _ in range(reps):
        bev.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    c1_ms = (_time.perf_counter() - t0) / reps * 1e3
    out["C1 create_voxel_pointcloud on the bundled sweep, drop-in call with host buffers"] = {
        "points": int(p4.shape[1]), "ms": round(c1_ms, 4), "cpu_port_ms": None,
        "note": "wall clock of a synchronous call (H2D + 2 kernels + D2H of the 1.35 MB grid): latency, not a roofline case"}
    # C1 resident on the device: the kernels alone
    rows1 = torch.from_numpy(synth.fixture_points_nx4()).to(dev)
    offs1 = np.array([0, rows1.shape[0]], dtype=np.int64)
    res1 = {"raw": torch.empty((1,) + synth.BEV_SHAPE, dtype=torch.float32, device=dev)}
    ms = timed(lambda: bev.rasterize_frames(rows1, offs1, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET,
                                            want=("raw",), out=res1))
    cells1 = int(np.prod(synth.BEV_SHAPE))
    out["C1 the same, device-resident (1 frame per launch)"] = dict(
        {"points": int(rows1.shape[0]), "ms": round(ms, 4)}, **roof(16 * rows1.shape[0] + 4 * cells1, ms))
    cloud = torch.from_numpy(synth.multisweep_cloud(20)).to(dev)       # 1,062,920 points, 10+ sweeps
    n = int(cloud.shape[0])
    offs = np.array([0, n], dtype=np.int64)
    for name, vs, rg, T, V in (("C2 SECOND 0.05 m voxels, T=5, V=60000", synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
                               ("C3 pillars 0.25 m, T=60, V=30000", synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)):
        vn = vg.voxelize_frames(cloud, offs, vs, rg, T, V, zero_tail=False)[3]
        v_act = int(vn[0].item())
        ms = timed(lambda: vg.voxelize_frames(cloud, offs, vs, rg, T, V, zero_tail=False))
        out[name] = dict({"points": n, "voxels": v_act, "ms": round(ms, 4), "points_per_s": round(n / (ms * 1e-3))},
                         **roof(16 * n + v_act * (T * 4 * 4 + 16), ms))
    # C2 batched: six 1.06 M-point clouds per call (copies of the cloud turned about z by k x 60 degrees).  The dense
    # first[] map of this grid is 216 MB per cloud; the open-addressing table is 16 MB, so six clouds share a launch
    nb = 6
    parts = [cloud]
    for k in range(1, nb):
        a = 2.0 * np.pi * k / nb
        rot = torch.tensor([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]], dtype=torch.float32, device=dev)
        q = cloud.clone()
        q[:, :2] = cloud[:, :2] @ rot
        parts.append(q)
    batch = torch.cat(parts).contiguous()
    boffs = np.arange(nb + 1, dtype=np.int64) * n
    vs, rg, T, V = synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000
    vn = vg.voxelize_frames(batch, boffs, vs, rg, T, V, zero_tail=False)[3]
    v_tot = int(vn.sum().item())
    ms = timed(lambda: vg.voxelize_frames(batch, boffs, vs, rg, T, V, zero_tail=False))
    out["C2 batched: %d clouds of 1.06 M points per call" % nb] = dict(
        {"points": nb * n, "voxels": v_tot, "ms": round(ms, 4), "ms_per_cloud": round(ms / nb, 4),
         "points_per_s": round(nb * n / (ms * 1e-3))}, **roof(16 * nb * n + v_tot * (T * 4 * 4 + 16), ms))
    # C3 batched: the same six clouds at the pillar config (dense 640 KB map per cloud): what a batch of ten-sweep
    # samples costs per call
    vs, rg, T, V = synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000
    vn = vg.voxelize_frames(batch, boffs, vs, rg, T, V, zero_tail=False)[3]
    v_tot = int(vn.sum().item())
    ms = timed(lambda: vg.voxelize_frames(batch, boffs, vs, rg, T, V, zero_tail=False))
    out["C3 batched: %d clouds of 1.06 M points per call" % nb] = dict(
        {"points": nb * n, "voxels": v_tot, "ms": round(ms, 4), "ms_per_cloud": round(ms / nb, 4),
         "points_per_s": round(nb * n / (ms * 1e-3))}, **roof(16 * nb * n + v_tot * (T * 4 * 4 + 16), ms))
    # C4 batched: six frames of 20 sweeps each (per-sweep 4x4) -> 1024^2 x 3 u8 + CHW with the map channels
    per = n // 20
    seg_offs6 = np.arange(20 * nb + 1, dtype=np.int64) * per
    seg_frame6 = np.repeat(np.arange(nb, dtype=np.int32), 20)
    seg_tm6 = np.concatenate([np.stack([synth.sweep_transform(s) for s in range(20)])] * nb)
    maps6 = torch.from_numpy(synth.map_raster(seed=4000)[None]).to(dev).expand(nb, -1, -1, -1).contiguous()
    res6 = {"u8": torch.empty((nb,) + synth.BEV1024_SHAPE, dtype=torch.uint8, device=dev),
            "chw": torch.empty((nb, 6, 1024, 1024), dtype=torch.float32, device=dev)}
    b6 = batch[:20 * nb * per]
    ms = timed(lambda: bev.rasterize_frames(b6, seg_offs6, synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE,
                                            synth.BEV_Z_OFFSET, seg_frame=seg_frame6, seg_tm=seg_tm6,
                                            n_frames=nb, want=("u8", "chw"), map_u8=maps6, out=res6))
    hw6 = 1024 * 1024
    out["C4 batched: %d frames of 20 sweeps per call" % nb] = dict(
        {"points": int(b6.shape[0]), "ms": round(ms, 4), "ms_per_frame": round(ms / nb, 4),
         "points_per_s": round(b6.shape[0] / (ms * 1e-3))}, **roof(16 * b6.shape[0] + nb * (3 * hw6 + 3 * hw6 + 24 * hw6), ms))
    del batch, parts, res6, maps6
    # C4: 1024^2 x 3 BEV of the same cloud, 20 sweeps each with its own sensor->car 4x4, u8 + CHW/map
    per = n // 20
    seg_offs = np.arange(21, dtype=np.int64) * per
    seg_tm = np.stack([synth.sweep_transform(s) for s in range(20)])
    maps = torch.from_numpy(synth.map_raster(seed=4000)[None]).to(dev)
    res = {"u8": torch.empty((1,) + synth.BEV1024_SHAPE, dtype=torch.uint8, device=dev),
           "chw": torch.empty((1, 6, 1024, 1024), dtype=torch.float32, device=dev)}
    ms = timed(lambda: bev.rasterize_frames(cloud, seg_offs, synth.BEV1024_SHAPE, synth.BEV1024_VOXEL_SIZE,
                                            synth.BEV_Z_OFFSET, seg_frame=np.zeros(20, np.int32), seg_tm=seg_tm,
                                            n_frames=1, want=("u8", "chw"), map_u8=maps, out=res))
    hw = 1024 * 1024
    out["C4 BEV 1024x1024x3, 20 sweeps with 4x4, u8 + (6,1024,1024) CHW with map"] = dict(
        {"points": n, "ms": round(ms, 4), "points_per_s": round(n / (ms * 1e-3))}, **roof(16 * n + 3 * hw + 3 * hw + 24 * hw, ms))
    # the drop-in call of a DataLoader worker (preprocess.py:299-317): host numpy in, host numpy out, wall clock
    sweep = synth.fixture_points_nx4()
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, max_voxels=30000)
    for _ in range(3):
        gen.generate(sweep, 30000)
    t0 = _time.perf_counter()
    for _ in range(reps):
        gen.generate(sweep, 30000)
    gen_ms = (_time.perf_counter() - t0) / reps * 1e3
    out["C3 VoxelGeneratorV2.generate on the bundled sweep, drop-in call with host buffers"] = {
        "points": int(sweep.shape[0]), "ms": round(gen_ms, 4), "cpu_port_ms": None}
    # block-filtering voxelizer (all.fhd.config) on the same cloud; target raster and PNG of 128 frames
    fhd_vs, fhd_rg = (0.05, 0.05, 0.2), (-50, -50, -5, 50, 50, 3)
    ms = timed(lambda: vg.voxelize_frames(cloud, offs, fhd_vs, fhd_rg, 5, 60000, zero_tail=False,
                                          block_filter=(1, 8, 0.2, 2.0)))
    out["n4 block-filtering voxelizer 0.05 m, T=5, V=60000, block 1x8, 0.2 m"] = {
        "points": n, "ms": round(ms, 4), "points_per_s": round(n / (ms * 1e-3))}
    scenes = [synth.box_scene(7000 + f, 60) for f in range(128)]
    d_c = torch.from_numpy(np.concatenate([c for c, _ in scenes])).to(dev)
    d_k = torch.from_numpy(np.concatenate([k + 1 for _, k in scenes]).astype(np.int32)).to(dev)
    box_offs = np.arange(129, dtype=np.int64) * 60
    ms = timed(lambda: bev.rasterize_targets(d_c, d_k, box_offs, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE,
                                             synth.BEV_Z_OFFSET))
    out["n4 target raster (draw_boxes) 128 frames x 60 boxes, 336x336"] = {
        "boxes": 128 * 60, "ms": round(ms, 4), "frames_per_s": round(128 / (ms * 1e-3))}
    return out


def time_png(eng, reps=10):
    """n4: lv_png_encode of the engine's 128 BEV u8 images (generating_train_bev.py:215) into device memory."""
    import numpy as np
    import torch
    from lyft3d_b200 import bev
    out, sizes = bev.encode_png_frames(eng.bev_u8)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        bev.encode_png_frames(eng.bev_u8, out=out, sizes=sizes)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms = float(np.median(ms))
    total = int(sizes.sum().item())
    return {"frames": eng.F, "ms": round(ms, 4), "png_bytes_per_frame": round(total / eng.F),
            "dense_bytes_per_frame": int(eng.bev_u8[0].numel()), "GBps_read": round(eng.bev_u8.numel() / (ms * 1e-3) / 1e9, 1)}


def time_reference_gpu_eager(eng, pts, reps=5):
    """The reference's OWN GPU path for decoration and scatter - chains of eager torch ops
    (pointpillars.py:203-231 and :444-476, restated op for op in oracle/pillar_torch_ref.py and pinned against
    the reference's outputs) - on the same B200 and the same batch, beside lv_pillar_decorate / lv_pillar_scatter.
    This, not a GPU-over-numpy ratio, is what a maintainer of the reference would see."""
    import numpy as np
    import torch
    from oracle import pillar_torch_ref as tr

    def timed(fn, reps=reps):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    eng.voxelize(pts)
    rows = eng.read_total_rows()
    v, num, co = eng.voxels[:rows], eng.num_points[:rows], eng.coords[:rows]
    vx, vy = eng.cfg.voxel_size[0], eng.cfg.voxel_size[1]
    xo, yo = vx / 2 + eng.cfg.coors_range[0], vy / 2 + eng.cfg.coors_range[1]
    res = {"pillars": int(rows), "batch": eng.F}
    res["decorate_ms_reference_eager"] = round(timed(lambda: tr.decorate(v, num, co, vx, vy, xo, yo)), 4)
    res["decorate_ms_ours"] = round(timed(lambda: eng.decorate(rows)), 4)
    feats = eng.features[:rows]
    res["scatter_ms_reference_eager"] = round(timed(lambda: tr.scatter(feats, co, eng.F, eng.ny, eng.nx)), 4)
    res["scatter_ms_ours"] = round(timed(lambda: eng.scatter(rows)), 4)
    ref = tr.scatter(feats, co, eng.F, eng.ny, eng.nx)
    res["scatter_equal"] = bool((ref == eng.canvas).all())
    del ref
    dec = tr.decorate(v, num, co, vx, vy, xo, yo)
    res["decorate_max_abs_diff"] = float((dec - eng.decorated[:rows]).abs().max())
    del dec
    res["speedup_decorate"] = round(res["decorate_ms_reference_eager"] / res["decorate_ms_ours"], 2)
    res["speedup_scatter"] = round(res["scatter_ms_reference_eager"] / res["scatter_ms_ours"], 2)
    torch.cuda.empty_cache()
    # a TRAINING step of the pillar feature net (pointpillars.py:203-231 + :51-65: decoration, Linear, BatchNorm1d with
    # batch statistics, ReLU, max; forward + backward for linear.weight, norm.weight, norm.bias) on 16 frames of the
    # batch: the reference's eager chain (it keeps the (P,60,9) and (P,60,64) tensors for autograd) beside the fused
    # op (lv_pillar_pfn_train_forward + lv_pillar_pfn_train_backward: five kernels)
    from lyft3d_b200 import pointpillars as pp, synth
    nsub = int(eng.voxel_offsets[min(16, eng.F)].item())
    vs_, ns_, cs_ = v[:nsub], num[:nsub], co[:nsub]
    torch.manual_seed(0)
    net = pp.PillarFeatureNet(4, True, (64,), False, synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE).to(v.device).train()
    g = torch.randn(nsub, 64, device=v.device)

    def step_fused():
        net.zero_grad(set_to_none=True)
        (net(vs_, ns_, cs_) * g).sum().backward()

    def step_eager():
        net.zero_grad(set_to_none=True)
        dec_ = tr.decorate(vs_, ns_, cs_, vx, vy, xo, yo)
        (net.pfn_layers[0](dec_).squeeze() * g).sum().backward()
    res["pfn_training_step"] = {"pillars": nsub, "frames": min(16, eng.F),
                                "ms_ours_fused": round(timed(step_fused), 4)}
    try:
        res["pfn_training_step"]["ms_reference_eager"] = round(timed(step_eager, reps=3), 4)
        res["pfn_training_step"]["speedup"] = round(res["pfn_training_step"]["ms_reference_eager"] /
                                                    res["pfn_training_step"]["ms_ours_fused"], 2)
    except torch.cuda.OutOfMemoryError as e:      # (P,60,64) activations and their gradients: several GB each
        res["pfn_training_step"]["ms_reference_eager"] = None
        res["pfn_training_step"]["note"] = "the eager chain ran out of memory: %s" % str(e)[:80]
    del net, g
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------- legs that run the oracle: their own process
def cpu_leg(spec):
    """`bench.py --leg cpu`: everything that executes oracle/ on the host cores, in its OWN process, so that the
    native arm never maps the oracle's library.  Prints one JSON object."""
    import numpy as np
    from lyft3d_b200 import synth
    from oracle import bev_oracle, cpu_path, voxel_oracle
    out = {}
    cfg = workload_cfg()
    n_frames = int(spec.get("cpu_frames", 0))
    if n_frames:
        cpu_path.run_pool([0], cfg, 1)  # warm-up (page-in, dense map allocation)
        pts, sec = cpu_path.run_pool(list(range(n_frames)), cfg, 1)
        out["cpu_baseline"] = {"value": pts / sec, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "%d frames of the same workload, single process (%.1f s)" % (n_frames, sec)}
    if spec.get("single"):
        p4 = synth.fixture_points_4xn()
        t0 = time.perf_counter()
        for _ in range(3):
            bev_oracle.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
        out["c1_cpu_port_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
        sweep = synth.fixture_points_nx4()
        ora = voxel_oracle.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
        ora.generate(sweep)
        t0 = time.perf_counter()
        for _ in range(3):
            ora.generate(sweep)
        out["c3_generate_cpu_port_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
    print(json.dumps(out))
    return 0


def verify_leg(spec):
    """`bench.py --leg verify`: the oracle's answer for ONE frame of the timed workload, compared with what the
    GPU produced for it (dumped by the native arm into spec["npz"]).  Prints {"ok": bool, ...}."""
    import hashlib
    import numpy as np
    from lyft3d_b200 import synth
    from oracle import bev_oracle, pillar_oracle, voxel_oracle
    g = np.load(spec["npz"])
    fr = synth.c5_frame(int(spec["frame_id"]))
    cfg = workload_cfg()
    raw = bev_oracle.create_voxel_pointcloud(np.ascontiguousarray(fr.T), cfg["bev_shape"], cfg["bev_voxel_size"], cfg["bev_z_offset"])
    u8 = bev_oracle.quantize_u8(bev_oracle.normalize_voxel_intensities(raw))
    v, c, n = voxel_oracle.points_to_voxel(fr, cfg["voxel_size"], cfg["pc_range"], cfg["max_points"], cfg["max_voxels"])
    coors = pillar_oracle.merge_batch_coords([c])
    dec = pillar_oracle.decorate(v, n, coors, cfg["voxel_size"], cfg["pc_range"])
    res = {"frame_id": int(spec["frame_id"]),
           "points_equal": bool(np.array_equal(g["points"].view(np.uint32), fr.view(np.uint32))),
           "bev_u8_equal": bool(np.array_equal(g["bev_u8"], u8)),
           "bev_u8_sha16": hashlib.sha256(u8.tobytes()).hexdigest()[:16],
           "voxel_num": int(v.shape[0]), "voxel_num_equal": int(g["voxel_num"]) == int(v.shape[0]),
           "coords_equal": bool(np.array_equal(g["coords"][:, 1:], c)),
           "num_points_equal": bool(np.array_equal(g["num_points"], n)),
           }
    # decoration: copy / single-subtraction channels equal, f_cluster inside the bound that any two float32 summation
    # orders satisfy (oracle/pillar_oracle.py::decorate_mismatch); the achieved error is reported beside it
    dec_ok, worst, max_abs = pillar_oracle.decorate_mismatch(g["decorated"], dec, v, n)
    res["decorated_within_summation_order_bound"] = bool(dec_ok)
    res["decorated_worst_error_over_bound"] = round(worst, 4)
    res["decorated_max_abs_diff"] = max_abs
    res["ok"] = all(v for k, v in res.items() if k.endswith("equal") or k.startswith("decorated_within"))
    print(json.dumps(res))
    return 0


def run_leg(name, spec, timeout=900):
    """Runs `bench.py --leg name` in a subprocess and returns its JSON (or {"error": ...})."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", name, "--leg-spec", json.dumps(spec)],
                             capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                             env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
        if out.returncode != 0 or not lines:
            return {"error": (out.stderr or out.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001 - a failed baseline leg must not lose the measured line
        return {"error": repr(e)}


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workers = os.cpu_count() or 1
    from oracle import cpu_path
    cfg = workload_cfg()
    per_step = 2 * workers
    cpu_path.run_pool(list(range(workers)), cfg, workers)
    for _ in range(args.warmup):
        cpu_path.run_pool(list(range(per_step)), cfg, workers)
    tot_pts, tot_sec = 0, 0.0
    for s in range(args.steps):
        ids = list(range(s * per_step, (s + 1) * per_step))
        p, t = cpu_path.run_pool(ids, cfg, workers)
        tot_pts += p
        tot_sec += t
    value = tot_pts / tot_sec
    sample = "%d steps x %d frames of the C5 workload over a %d-process pool" % (args.steps, per_step, workers)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_sec / max(args.steps, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 coords, u32 counts",
            "data": "synthetic", "config": config_json(args.frames_per_step, args.gpus),
            "execution": "the oracle chain of oracle/cpu_path.py over a process pool on the host cores; every step is a "
                         "bounded sample of the workload: " + sample,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--pool-frames", type=int, default=256)
    ap.add_argument("--cpu-frames", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the single-cloud timings of configs[0..3]")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--serial", action="store_true", help="headline from the one-stream loop (no stream pipelining)")
    ap.add_argument("--strong", action="store_true",
                    help="also time the WHOLE 8192-frame set of BASELINE configs[4], sharded f mod G (strong scaling)")
    ap.add_argument("--strong-frames", type=int, default=8192)
    ap.add_argument("--no-verify", action="store_true", help="skip the post-run check of one frame per rank against the oracle")
    ap.add_argument("--no-eager-ref", action="store_true", help="skip the eager-torch timing of the reference's GPU ops")
    ap.add_argument("--leg", default=None, choices=["cpu", "verify"], help=argparse.SUPPRESS)
    ap.add_argument("--leg-spec", default="{}", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.leg == "cpu":
        return cpu_leg(json.loads(args.leg_spec))
    if args.leg == "verify":
        return verify_leg(json.loads(args.leg_spec))
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from lyft3d_b200 import _native as nat
    from lyft3d_b200.engine import FrameBatchEngine, HostPipeline, PipelinedEngine, bind_to_gpu_numa_node, shard_frames
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    F = args.frames_per_step

    # ---- the workload, generated on the host BEFORE the first CUDA call (fork pool): rank r owns the frames
    # f = r (mod G) of BASELINE configs[4] (engine.shard_frames); the weak-scaling steps cycle a pool of its first
    # `pool_frames` frames, the strong-scaling pass walks all of its shard of the 8192-frame set
    gen_workers = max(1, min(8, (os.cpu_count() or 1) // max(world, 1)))
    my_ids = shard_frames(args.strong_frames, rank, world)
    n_pool = max(args.pool_frames, F) // F * F
    strong_ids = my_ids[:len(my_ids) // F * F] if args.strong else []
    need_ids = my_ids[:max(n_pool, len(strong_ids))]
    if len(need_ids) < n_pool:       # tiny --strong-frames: wrap around
        need_ids = (need_ids * (n_pool // max(len(need_ids), 1) + 1))[:n_pool]
    host_all = make_frames_host(need_ids, gen_workers)
    n = host_all.shape[0] // len(need_ids)

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    if not os.path.exists(nat.LIB_PATH):
        if local == 0:
            ge.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None if args.no_numa_bind else bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    all_pts = torch.from_numpy(host_all).to(dev)
    pool_pts = all_pts[:n_pool * n]
    n_batches = n_pool // F
    eng = FrameBatchEngine(local, F, n)
    torch.manual_seed(1234 + rank)
    eng.features.copy_(torch.randn((eng.cap, eng.channels), device=dev))
    cells = eng.bev_shape[0] * eng.bev_shape[1] * eng.bev_shape[2]

    def batch(s):
        b = s % n_batches
        return pool_pts[b * F * n:(b + 1) * F * n]

    stages = ["bev", "pillarize", "scatter"]

    def run_steps(k, record):
        evs, rows_seen = [], []
        for s in range(k):
            pts = batch(s)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] if record else None
            if record:
                e[0].record()
            eng.bev(pts)
            if record:
                e[1].record()
            eng.pillarize(pts)
            rows = eng.read_total_rows()
            if record:
                e[2].record()
            eng.scatter(rows)
            if record:
                e[3].record()
                evs.append(e)
            rows_seen.append(rows)
        return evs, rows_seen

    # ---- pass 1: stages back to back on one stream, CUDA events around every stage ---------
    # (the per-stage times behind `stages` and `roofline`)
    run_steps(args.warmup, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    evs, rows_seen = run_steps(args.steps, True)
    t_end.record()
    torch.cuda.synchronize()
    serial_ms_total = t_start.elapsed_time(t_end)

    # ---- pass 2 (the headline): the same steps through PipelinedEngine - BEV and the voxelizer of
    # step i+1 run on high-priority streams next to the canvas stream of step i, no host sync
    pipe_eng = PipelinedEngine(eng) if not args.serial else None

    def launches_now():
        return eng.h.launches() + (pipe_eng.launches() if pipe_eng is not None else 0)

    if pipe_eng is not None:
        for s in range(args.warmup):
            pipe_eng.submit(batch(s))
        pipe_eng.drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = launches_now()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start.record()
    if pipe_eng is not None:
        for s in range(args.steps):
            pipe_eng.submit(batch(s))
        pipe_eng.drain()
    else:
        run_steps(args.steps, False)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = launches_now() - launches0
    ms_total = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None

    stage_ms = {st: 0.0 for st in stages}
    for e in evs:
        for i, st in enumerate(stages):
            stage_ms[st] += e[i].elapsed_time(e[i + 1])
    mean_rows = sum(rows_seen) / max(len(rows_seen), 1)

    # ---- strong scaling: the WHOLE 8192-frame set of configs[4], rank r works through its shard f = r mod G
    strong_ms = float("nan")
    if args.strong and strong_ids:
        sp = all_pts[:len(strong_ids) * n]
        k_steps = len(strong_ids) // F
        runner = pipe_eng

        def strong_pass():
            for s in range(k_steps):
                pts = sp[s * F * n:(s + 1) * F * n]
                if runner is not None:
                    runner.submit(pts)
                else:
                    eng.bev(pts)
                    eng.pillarize(pts)
                    eng.scatter_dev()
            if runner is not None:
                runner.drain()
        strong_pass()                      # warm-up pass over the set
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        strong_pass()
        b.record()
        torch.cuda.synchronize()
        strong_ms = a.elapsed_time(b)

    # ---- post-run check: frame 0 of the LAST timed batch of this rank, against the oracle (own process) and,
    # for the canvas, against an independent torch gather
    verify = None
    if not args.no_verify:
        last = (args.steps - 1) % n_batches
        pts = batch(args.steps - 1)
        rows = eng.step(pts)
        torch.cuda.synchronize()
        offs = eng.voxel_offsets.cpu().numpy()
        r0, r1 = int(offs[0]), int(offs[1])
        co = eng.coords[r0:r1].long()
        cv = eng.canvas[0]
        canvas_ok = bool((cv[:, co[:, 2], co[:, 3]].t() == eng.features[r0:r1]).all()) and \
            int((cv != 0).sum()) == int((eng.features[r0:r1] != 0).sum())
        import tempfile
        tmp = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
        tmp.close()
        np.savez(tmp.name, points=pts[:n].cpu().numpy(), bev_u8=eng.bev_u8[0].cpu().numpy(), voxel_num=int(eng.voxel_num[0].item()),
                 coords=eng.coords[r0:r1].cpu().numpy(), num_points=eng.num_points[r0:r1].cpu().numpy(),
                 decorated=eng.decorated[r0:r1].cpu().numpy())
        verify = run_leg("verify", {"npz": tmp.name, "frame_id": int(need_ids[last * F])})
        os.unlink(tmp.name)
        verify["canvas_equals_gather_of_features"] = canvas_ok
        verify["ok"] = bool(verify.get("ok")) and canvas_ok

    # ---- reported separately (not part of `value`): the pillar path with the
    # PFNLayer fused behind the decoration - points -> (rows, 64) features -> canvas
    from lyft3d_b200 import pointpillars as pp
    pfn_net = pp.PillarFeatureNet(4, True, (eng.channels,), False, eng.cfg.voxel_size[:], eng.cfg.coors_range[:]).to(dev).eval()
    pw, pscale, pshift = pp.fold_pfn_layer(pfn_net.pfn_layers[0])
    pfn_ms = None
    if eng.channels == 64:
        keep = eng.features.clone()
        for s in range(3):
            eng.pillar_features(batch(s), pw, pscale, pshift)
            eng.scatter(eng.read_total_rows())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            eng.pillar_features(batch(s), pw, pscale, pshift)
            eng.scatter(eng.read_total_rows())
        e1.record()
        torch.cuda.synchronize()
        pfn_ms = e0.elapsed_time(e1) / args.steps
        eng.features.copy_(keep)
        del keep

    # ---- end to end: pinned host points in; PNG files, file sizes and voxel_num back on the host ------------
    pipe = HostPipeline(eng)
    host_batches = []
    for b in range(min(n_batches, 2)):
        hb = torch.empty((F * n, 4), dtype=torch.float32, pin_memory=True)
        hb.copy_(batch(b))
        host_batches.append(hb)
    e2e_sec = float("nan")
    host_seen = {"png_bytes": 0, "pillars": 0, "steps": 0}

    def consume(step, view):   # the host reads every step's result: file sizes and pillar counts
        host_seen["png_bytes"] += int(view["sizes"].sum())
        host_seen["pillars"] += int(view["voxel_num"].sum())
        host_seen["steps"] += 1

    if not args.no_e2e:
        pipe.run(host_batches, 3)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        pipe.run(host_batches, args.steps, consume=consume)
        e2e_sec = time.perf_counter() - t0

    t = torch.tensor([ms_total, e2e_sec, serial_ms_total, strong_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_sec, serial_ms_total, strong_ms = (float(v) for v in t)
    ok_all = torch.tensor([1 if (verify is None or verify.get("ok")) else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    pts_per_step_all = F * n * world
    value = pts_per_step_all * args.steps / (ms_total * 1e-3)
    e2e_value = pts_per_step_all * args.steps / e2e_sec

    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = {}
        if os.path.exists(pk):
            with open(pk) as f:
                peaks = json.load(f)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # algorithmic bytes per step and stage
        C, T = 4, eng.T
        alg = {
            "bev": F * (16 * n + 4 * cells + 1 * cells),
            "pillarize": F * 16 * n + mean_rows * (T * (C + 5) * 4 + 16 + 4),
            "scatter": mean_rows * (eng.channels * 4 + 16) + F * eng.channels * eng.ny * eng.nx * 4,
        }
        stage_info = {}
        for st in stages:
            ms = stage_ms[st] / args.steps
            gbs = alg[st] / (ms * 1e-3) / 1e9
            stage_info[st] = {"ms_per_step": round(ms, 4), "algorithmic_MB_per_step": round(alg[st] / 1e6, 2),
                              "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}
        dom = max(stages, key=lambda s: stage_ms[s])
        kernel_names = {"bev": "bev_hist_kernel + bev_finalize_flat4_kernel",
                        "scatter": "pillar_canvas_q_kernel (+ pillar_index_kernel, 1.5% of the stage)",
                        "pillarize": "vx_cells/assign/keys/scan_hist/scatter + vx_bins_kernel<decorate>"}
        # DRAM traffic of the dominant kernel per launch from the committed ncu --set full capture
        # (profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at this workload), valid only
        # for the kernel sources it was captured from
        traffic, traffic_note = None, "no capture for these sources"
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("source_sha16") != source_sha():
                traffic_note = "profiles/r02_traffic.json was captured from other kernel sources (sha %s, now %s)" % (
                    tj.get("source_sha16"), source_sha())
            elif tj.get("frames_per_step") == F and dom in tj.get("stages", {}):
                traffic = tj["stages"][dom]["dram_bytes_per_launch"]
                traffic_note = "ncu --set full, %s" % tj.get("captured", "")
        roof = {"bound": "hbm", "kernel": kernel_names[dom], "stage": dom,
                "achieved": stage_info[dom]["GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(stage_info[dom]["GBps"] / peak, 4), "traffic": traffic, "traffic_source": traffic_note,
                "algorithmic_bytes_per_launch": int(alg[dom]),
                "whole_step_frac": round(sum(alg[st] for st in stages) / (ms_total / args.steps * 1e-3) / 1e9 / peak, 4)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 points, f64 BEV affine, u32 counts",
                "data": "synthetic",
                "config": config_json(F, world),
                "frame_pool": "a pool of %d distinct frames per GPU is cycled (weak scaling); --strong walks the whole set" % n_pool,
                "mean_pillars_per_frame": round(mean_rows / F, 1),
                "execution": "stages of neighbouring steps overlap on 3 streams (PipelinedEngine), no host sync"
                             if pipe_eng is not None else "one stream, stages back to back",
                "roofline": roof, "stages": stage_info,
                "one_stream_ms_per_step": round(serial_ms_total / args.steps, 4),
                "pillar_path_with_fused_pfn": None if pfn_ms is None else {
                    "ms_per_step": round(pfn_ms, 4),
                    "note": "points -> voxelize+decorate+PFNLayer(eval) in one pipeline -> scatter; the unfused "
                            "chain it replaces is pillarize + a PyTorch PFNLayer + scatter (not part of `value`)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                        "d2h_bytes_per_step": pipe.d2h_bytes, "rank0_numa_node": numa_node,
                        "host_read": dict(host_seen),
                        "note": "HostPipeline: pinned host points -> device -> both paths -> the BEV images as PNG "
                                "FILES (generating_train_bev.py:215; lv_png_encode writes them into mapped pinned "
                                "host memory) + file sizes + voxel_num read on the host every step, no host sync inside "
                                "the loop; the canvas stays on the device (its consumer is the RPN, voxelnet.py:336)"},
                "verified_against_oracle": {"all_ranks_ok": bool(int(ok_all.item())), "rank0": verify}}
        if args.strong and strong_ids:
            tot_frames = len(strong_ids) * world
            line["strong_scaling"] = {
                "frames": tot_frames, "steps_per_gpu": len(strong_ids) // F, "ms_whole_set": round(strong_ms, 3),
                "points_per_s": tot_frames * n / (strong_ms * 1e-3),
                "note": "the whole %d-frame set of BASELINE configs[4], frame f on GPU f mod %d, device-resident, "
                        "timed with CUDA events, max over ranks" % (tot_frames, world)}
        if world == 1 and not args.no_other_configs:
            line["other_configs"] = time_other_configs(dev, peak)
            line["other_configs"]["n4 PNG encode (lv_png_encode) of the 128 BEV images"] = time_png(eng)
        if world == 1 and not args.no_eager_ref:
            line["reference_gpu_eager"] = time_reference_gpu_eager(eng, batch(0))
        if world == 1 and not args.no_cpu_baseline:
            leg = run_leg("cpu", {"cpu_frames": args.cpu_frames, "single": not args.no_other_configs})
            if "cpu_baseline" in leg:
                line["cpu_baseline"] = leg["cpu_baseline"]
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "leg failed: %s" % leg.get("error")}
            oc = line.get("other_configs", {})
            for key, k2 in (("C1 create_voxel_pointcloud on the bundled sweep, drop-in call with host buffers", "c1_cpu_port_ms"),
                            ("C3 VoxelGeneratorV2.generate on the bundled sweep, drop-in call with host buffers", "c3_generate_cpu_port_ms")):
                if key in oc and k2 in leg:
                    oc[key]["cpu_port_ms"] = leg[k2]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
