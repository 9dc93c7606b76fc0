"""A/B: pillarize 128 frames in one call vs two half-batches on two streams (two handles): does the latency-bound
prologue of one half hide under the store-bound bins kernel of the other?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lyft3d_b200 import _native as nat
from lyft3d_b200.engine import FrameBatchEngine

dev = torch.device("cuda", 0)
F = 128
pool, n = bench.make_pool_frames(256, dev, 0)
full = FrameBatchEngine(0, F, n)
halves = [FrameBatchEngine(0, F // 2, n, handle=nat.Handle(0), voxel_capacity=full.cap // 2) for _ in range(2)]
streams = [torch.cuda.Stream(dev, priority=p) for p in (0, -1)]

def t(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

k = [0]
def one():
    k[0] ^= 1
    full.pillarize(pool[k[0] * F * n:(k[0] + 1) * F * n])

def two(delay_second=False):
    k[0] ^= 1
    base = k[0] * F * n
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    dones = []
    for i, (e, s) in enumerate(zip(halves, streams)):
        with torch.cuda.stream(s):
            s.wait_event(ev)
            e.pillarize(pool[base + i * (F // 2) * n: base + (i + 1) * (F // 2) * n])
            d = torch.cuda.Event(); d.record(s); dones.append(d)
    for d in dones: cur.wait_event(d)

def seq_halves():
    k[0] ^= 1
    base = k[0] * F * n
    for i, e in enumerate(halves):
        e.pillarize(pool[base + i * (F // 2) * n: base + (i + 1) * (F // 2) * n])

print("one call, 128 frames      %.4f ms" % t(one))
print("two halves, one stream    %.4f ms" % t(seq_halves))
print("two halves, two streams   %.4f ms" % t(two))
