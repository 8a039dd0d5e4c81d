"""N>1 host logic on CPU with gloo, world_size 2: the view -> rank assignment bench.py uses covers every
view exactly once per step, and the max-over-ranks timing reduction behaves (no data-path collective
exists on the raster path — views are independent)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, steps, n_poses, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        views = [(i * world + rank) % n_poses for i in range(steps)]
        mine = torch.zeros(steps * world, dtype=torch.int64)
        for i, v in enumerate(views):
            mine[i * world + rank] = v + 1
        dist.all_reduce(mine)  # gather by sum: every slot written by exactly one rank
        ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            out.put((mine.tolist(), ms.item()))
    finally:
        dist.destroy_process_group()


def test_view_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, steps, n_poses = 2, 6, 8
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, n_poses, q)) for r in range(world)]
    for p in procs:
        p.start()
    slots, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(s > 0 for s in slots)                       # every (step, rank) slot rendered once
    assert [s - 1 for s in slots] == [k % n_poses for k in range(steps * world)]  # consecutive views, no overlap
    assert ms == 15.0                                      # max over ranks
