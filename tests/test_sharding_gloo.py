"""World-size-2 gloo test (CPU) of the host-side multi-GPU logic: frames are sharded
f mod G with no data-path collective; the only collective is the MAX-reduce of the
timing scalars that bench.py performs after the timed region."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lyft3d_b200.engine import shard_frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_frames(n_frames, rank, world)
    # per-rank "work": the oracle BEV count of each owned frame would go here; the host
    # logic under test is the partition and the timing reduction
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([10.0 + rank, 1.0 + 2 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, t.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8192, 11])
def test_frame_sharding_world2(n_frames):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, t = q.get()
    allf = sorted(gathered[0] + gathered[1])
    assert allf == list(range(n_frames))                    # every frame exactly once
    assert set(gathered[0]).isdisjoint(gathered[1])
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1    # balanced
    assert t == [11.0, 3.0]                                 # max over ranks, as bench.py reports
