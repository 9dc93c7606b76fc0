#!/usr/bin/env python
"""Aggregate pinned-host -> device copy bandwidth of one box with N ranks copying at once (torchrun), next to
what one rank reaches alone: the ceiling of bench.py's end-to-end leg, whose every step ships 109 MB of points
per rank.  python -m torch.distributed.run --nproc-per-node N tools/h2d_aggregate.py"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 109 * (1 << 20)
    host = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for h in host:
        h.fill_(rank)
    dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
    back = torch.empty(43 * (1 << 20), dtype=torch.uint8, pin_memory=True)
    dsrc = torch.empty(43 * (1 << 20), dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    for with_d2h in (False, True):
        for _ in range(3):
            dev[0].copy_(host[0], non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 50
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(s_in):
                dev[i & 1].copy_(host[i & 1], non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(s_out):
                    back.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gbs = world * reps * n / float(t[0]) / 1e9
            print("ranks %d  %s: aggregate H2D %.1f GB/s (%.1f per rank)" %
                  (world, "H2D 109 MB + D2H 43 MB per step" if with_d2h else "H2D only", gbs, gbs / world))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
