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
    c = {"workload": "C5 throughput sweep: synthetic 53,146-point Lyft sweeps, BEV 336x336x3 (norm f32 + u8) "
                     "+ pillar path (voxelize 0.25 m/T=60/V=30000 -> decorate -> scatter 64x400x400)",
         "pillar_path": "fused voxelize+decorate" if not (extra or {}).get("unfused") else "separate voxelize, decorate",
         "frames_per_step_per_gpu": frames_per_step, "points_per_frame": 53146,
         "parallelism": "frames sharded f mod G, dp%d, no collective" % n_gpus,
         "l2": "inputs (pool of distinct frames) and outputs per step exceed the 126 MB L2"}
    if extra:
        c.update(extra)
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


def make_pool_frames(n_frames, device, seed):
    """Pool of distinct C5-shaped frames generated on the device (rotation about z + jitter
    of the bundled sweep; the seeded numpy definition synth.c5_frame is what the parity
    tests use).  Returns (n_frames*N, 4) float32."""
    import math
    import torch
    from lyft3d_b200 import synth
    base = torch.from_numpy(synth.fixture_points_nx4()).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(5000 + seed)
    n = base.shape[0]
    out = torch.empty((n_frames, n, 4), dtype=torch.float32, device=device)
    ang = torch.rand(n_frames, generator=g, device=device) * (2 * math.pi)
    c, s = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    noise = torch.randn((n_frames, n, 3), generator=g, device=device) * 0.02
    out[:, :, 0] = c * base[None, :, 0] - s * base[None, :, 1] + noise[:, :, 0]
    out[:, :, 1] = s * base[None, :, 0] + c * base[None, :, 1] + noise[:, :, 1]
    out[:, :, 2] = base[None, :, 2] + noise[:, :, 2]
    out[:, :, 3] = base[None, :, 3]
    return out.reshape(n_frames * n, 4), n


def time_other_configs(dev, reps=10):
    """BASELINE.json configs[0..3] as single-cloud latencies (configs[1..3]: CUDA events, median of
    `reps`, inputs resident on the device; the headline workload is configs[4])."""
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

    out = {}
    # C1: the reference's own CPU-runnable case through the drop-in call, numpy in -> numpy out
    # (H2D, kernels, D2H and the synchronisation inside the call), wall clock; the CPU port beside it
    import time as _time
    from oracle import bev_oracle
    p4 = synth.fixture_points_4xn()
    for _ in range(3):
        bev.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    t0 = _time.perf_counter()
    for _ in range(reps):
        bev.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    c1_ms = (_time.perf_counter() - t0) / reps * 1e3
    t0 = _time.perf_counter()
    for _ in range(3):
        bev_oracle.create_voxel_pointcloud(p4, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    cpu_ms = (_time.perf_counter() - t0) / 3 * 1e3
    out["C1 create_voxel_pointcloud on the bundled sweep, drop-in call with host buffers"] = {
        "points": int(p4.shape[1]), "ms": round(c1_ms, 4), "cpu_port_ms": round(cpu_ms, 3)}
    cloud = torch.from_numpy(synth.multisweep_cloud(20)).to(dev)       # 1,062,920 points, 10+ sweeps
    n = int(cloud.shape[0])
    offs = np.array([0, n], dtype=np.int64)
    for name, vs, rg, T, V in (("C2 SECOND 0.05 m voxels, T=5, V=60000", synth.SECOND_VOXEL_SIZE, synth.SECOND_RANGE, 5, 60000),
                               ("C3 pillars 0.25 m, T=60, V=30000", synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)):
        ms = timed(lambda: vg.voxelize_frames(cloud, offs, vs, rg, T, V, zero_tail=False))
        out[name] = {"points": n, "ms": round(ms, 4), "points_per_s": round(n / (ms * 1e-3))}
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
    out["C4 BEV 1024x1024x3, 20 sweeps with 4x4, u8 + (6,1024,1024) CHW with map"] = {
        "points": n, "ms": round(ms, 4), "points_per_s": round(n / (ms * 1e-3))}
    # the drop-in call of a DataLoader worker (preprocess.py:299-317): host numpy in, host numpy out, wall clock
    from oracle import voxel_oracle
    sweep = synth.fixture_points_nx4()
    gen = vg.VoxelGeneratorV2(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, max_voxels=30000)
    for _ in range(3):
        gen.generate(sweep, 30000)
    t0 = _time.perf_counter()
    for _ in range(reps):
        gen.generate(sweep, 30000)
    gen_ms = (_time.perf_counter() - t0) / reps * 1e3
    ora = voxel_oracle.VoxelOracle(synth.PILLAR_VOXEL_SIZE, synth.PILLAR_RANGE, 60, 30000)
    ora.generate(sweep)
    t0 = _time.perf_counter()
    for _ in range(3):
        ora.generate(sweep)
    ora_ms = (_time.perf_counter() - t0) / 3 * 1e3
    out["C3 VoxelGeneratorV2.generate on the bundled sweep, drop-in call with host buffers"] = {
        "points": int(sweep.shape[0]), "ms": round(gen_ms, 4), "cpu_port_ms": round(ora_ms, 3)}
    # block-filtering voxelizer (all.fhd.config) on the same cloud; target raster of 128 frames
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


def cpu_baseline(n_frames, workers):
    from oracle import cpu_path
    cfg = workload_cfg()
    cpu_path.run_pool([0], cfg, 1)  # warm-up (page-in, dense map allocation)
    pts, sec = cpu_path.run_pool(list(range(n_frames)), cfg, workers)
    return pts / sec, pts, sec


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
            "data": "synthetic", "config": config_json(per_step, args.gpus, {"sample": sample}),
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
    ap.add_argument("--unfused", action="store_true", help="voxelize and decorate as separate stages")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the single-cloud timings of configs[1..3]")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--serial", action="store_true", help="headline from the one-stream loop (no stream pipelining)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from lyft3d_b200 import _native as nat
    from lyft3d_b200.engine import FrameBatchEngine
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    if not os.path.exists(nat.LIB_PATH):
        if local == 0:
            ge.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from lyft3d_b200.engine import bind_to_gpu_numa_node
    numa_node = None if args.no_numa_bind else bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    F = args.frames_per_step
    pool_pts, n = make_pool_frames(max(args.pool_frames, F), dev, seed=rank)
    n_batches = max(args.pool_frames, F) // F
    eng = FrameBatchEngine(local, F, n)
    feats_src = torch.randn((eng.cap, eng.channels), device=dev)
    eng.features.copy_(feats_src)
    cells = eng.bev_shape[0] * eng.bev_shape[1] * eng.bev_shape[2]

    def batch(s):
        b = s % n_batches
        return pool_pts[b * F * n:(b + 1) * F * n]

    fused = not args.unfused
    stages = ["bev", "pillarize", "scatter"] if fused else ["bev", "voxelize", "decorate", "scatter"]

    def run_steps(k, record):
        evs = []
        rows_seen = []
        for s in range(k):
            pts = batch(s)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] if record else None
            i = 0
            if record:
                e[i].record()
            eng.bev(pts)
            i += 1
            if record:
                e[i].record()
            if fused:
                eng.pillarize(pts)
                rows = eng.read_total_rows()
            else:
                eng.voxelize(pts)
                rows = eng.read_total_rows()
                i += 1
                if record:
                    e[i].record()
                eng.decorate(rows)
            i += 1
            if record:
                e[i].record()
            eng.scatter(rows)
            i += 1
            if record:
                e[i].record()
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
    from lyft3d_b200.engine import PipelinedEngine
    pipe_eng = PipelinedEngine(eng) if fused and not args.serial else None
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
    launches0 = eng.h.launches() + (pipe_eng.launches() if pipe_eng is not None else 0)
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
    launches = eng.h.launches() + (pipe_eng.launches() if pipe_eng is not None else 0) - launches0
    ms_total = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None

    stage_ms = {st: 0.0 for st in stages}
    for e in evs:
        for i, st in enumerate(stages):
            stage_ms[st] += e[i].elapsed_time(e[i + 1])
    mean_rows = sum(rows_seen) / max(len(rows_seen), 1)

    # ---- reported separately (not part of `value`): the pillar path with the
    # PFNLayer fused behind the decoration - points -> (rows, 64) features -> canvas
    from lyft3d_b200 import pointpillars as pp
    pfn_net = pp.PillarFeatureNet(4, True, (eng.channels,), False, eng.cfg.voxel_size[:], eng.cfg.coors_range[:]).to(dev).eval()
    pw, pscale, pshift = pp.fold_pfn_layer(pfn_net.pfn_layers[0])
    pfn_ms = None
    if eng.channels == 64:
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

    # ---- end to end: pinned host points in, BEV u8 + voxel_num back on the host ------------
    from lyft3d_b200.engine import HostPipeline
    pipe = HostPipeline(eng, fused=fused)
    host_batches = []
    for b in range(min(n_batches, 2)):
        hb = torch.empty((F * n, 4), dtype=torch.float32, pin_memory=True)
        hb.copy_(batch(b))
        host_batches.append(hb)
    e2e_sec = float("nan")
    if not args.no_e2e:
        pipe.run(host_batches, 3)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        pipe.run(host_batches, args.steps)
        e2e_sec = time.perf_counter() - t0

    t = torch.tensor([ms_total, e2e_sec, serial_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_sec, serial_ms_total = float(t[0]), float(t[1]), float(t[2])
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
            "voxelize": F * 16 * n + mean_rows * (T * C * 4 + 16 + 4),
            "decorate": mean_rows * (T * C * 4 + 20 + T * (C + 5) * 4),
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
                        "voxelize": "vx_cells/assign/keys/scan_hist/scatter + vx_bins_kernel<voxels>",
                        "decorate": "pillar_decorate_fast_kernel",
                        "scatter": "pillar_canvas_q_kernel (+ pillar_index_kernel, 1.5% of the stage)",
                        "pillarize": "vx_cells/assign/keys/scan_hist/scatter + vx_bins_kernel<decorate>"}
        # DRAM traffic of the dominant kernel per launch from the committed ncu --set full capture
        # (profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum at this workload)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("frames_per_step") == F and dom in tj.get("stages", {}):
                traffic = tj["stages"][dom]["dram_bytes_per_launch"]
        roof = {"bound": "hbm", "kernel": kernel_names[dom], "stage": dom,
                "achieved": stage_info[dom]["GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(stage_info[dom]["GBps"] / peak, 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": int(alg[dom]),
                "whole_step_frac": round(sum(alg[st] for st in stages) / (ms_total / args.steps * 1e-3) / 1e9 / peak, 4)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 points, f64 BEV affine, u32 counts",
                "data": "synthetic",
                "config": config_json(F, world, {"mean_pillars_per_frame": round(mean_rows / F, 1), "unfused": not fused,
                                                 "execution": "stages of neighbouring steps overlap on 3 streams "
                                                              "(PipelinedEngine), no host sync" if pipe_eng is not None
                                                 else "one stream, stages back to back"}),
                "roofline": roof, "stages": stage_info,
                "one_stream_ms_per_step": round(serial_ms_total / args.steps, 4),
                "pillar_path_with_fused_pfn": None if pfn_ms is None else {
                    "ms_per_step": round(pfn_ms, 4),
                    "note": "points -> voxelize+decorate+PFNLayer(eval) in one pipeline -> scatter; the unfused "
                            "chain it replaces is pillarize + a PyTorch PFNLayer + scatter (not part of `value`)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                        "d2h_bytes_per_step": pipe.d2h_bytes, "rank0_numa_node": numa_node,
                        "note": "HostPipeline: pinned host points -> device -> both paths -> the BEV images as PNG "
                                "FILES (generating_train_bev.py:215; lv_png_encode writes them into mapped pinned "
                                "host memory) + file sizes + voxel_num on the host every step, no host sync inside "
                                "the loop; the canvas stays on the device (its consumer is the RPN, voxelnet.py:336)"}}
        if world == 1 and not args.no_other_configs:
            line["other_configs"] = time_other_configs(dev)
        if world == 1 and not args.no_cpu_baseline:
            v, pts, sec = cpu_baseline(args.cpu_frames, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d frames of the same workload, single process (%.1f s)" %
                                              (args.cpu_frames, sec)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
