"""N>1 host logic on CPU with gloo, world_size 2: the view -> rank assignment bench.py uses (step i, rank r ->
pose (i + r) mod 8) renders `world` distinct views per step and cycles every rank through all poses, and the
max-over-ranks timing reduction behaves (no data-path collective
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
        views = [(i + rank) % n_poses for i in range(steps)]  # bench.py run_b200.view_of
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
    world, steps, n_poses = 2, 8, 8
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, n_poses, q)) for r in range(world)]
    for p in procs:
        p.start()
    slots, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(s > 0 for s in slots)                       # every (step, rank) slot rendered once
    v = [s - 1 for s in slots]
    for i in range(steps):                                 # per step: `world` distinct, consecutive views
        assert v[i * world:(i + 1) * world] == [(i + r) % n_poses for r in range(world)]
    for r in range(world):                                 # over n_poses steps every rank renders every pose once
        assert sorted(v[r::world]) == list(range(n_poses))
    assert ms == 15.0                                      # max over ranks


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussian_splatting_b200.view_parallel import GradientBucket, all_reduce_statistics, views_of_rank

        torch.manual_seed(0)  # same parameters on every rank (replicated Gaussians)
        params = [torch.randn(7, 3, requires_grad=True), torch.randn(7, 1, requires_grad=True),
                  torch.randn(7, 3, 15, requires_grad=True)]
        # (1) attach mode: .grad are views of one flat buffer, autograd accumulates in place
        bucket = GradientBucket(params).attach()
        bucket.zero()
        w = float(views_of_rank(0, rank, world, 8) + 1)  # a per-rank "view": loss = w * sum(p^2)
        loss = sum((p * p).sum() for p in params) * w
        loss.backward()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
        bucket.all_reduce(average=True)
        mean_w = sum(views_of_rank(0, r, world, 8) + 1 for r in range(world)) / world
        err_attach = max(float((p.grad - 2 * mean_w * p.detach()).abs().max()) for p in params)
        # (2) adopt mode: gradients already live in one allocation (what the fused backward leaves behind)
        flat = torch.empty(sum(p.numel() for p in params))
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view(p.shape)
            p.grad.copy_(2 * w * p.detach())
            off += p.numel()
        b2 = GradientBucket.adopt(flat, params)
        assert b2.zero_copy
        b2.all_reduce(average=False)
        sum_w = sum(views_of_rank(0, r, world, 8) + 1 for r in range(world))
        err_adopt = max(float((p.grad - 2 * sum_w * p.detach()).abs().max()) for p in params)
        # (3) gradients somewhere else -> flatten copy
        for p in params:
            p.grad = (2 * w * p.detach()).clone()
        b3 = GradientBucket.adopt(flat, params)
        assert not b3.zero_copy
        b3.all_reduce(average=False)
        err_copy = max(float((p.grad - 2 * sum_w * p.detach()).abs().max()) for p in params)
        stats = torch.full((7, 2), float(rank + 1))
        all_reduce_statistics([stats])
        if rank == 0:
            out.put((err_attach, err_adopt, err_copy, float(stats[0, 0])))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_all_reduce_two_ranks():
    """view_parallel.GradientBucket: one flat all-reduce of the parameter gradients (attach / adopt / copy)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err_attach, err_adopt, err_copy, stat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err_attach < 1e-5 and err_adopt < 1e-5 and err_copy < 1e-5
    assert stat == 3.0  # 1 + 2


def _adc_worker(rank, world, port, out):
    """Densification with several ranks: per-view statistics differ per rank, are summed right before a pass
    (all_reduce_statistics), and every rank must then take IDENTICAL decisions (same plan, same random split samples)
    — that is what keeps the replicas of tools/e2e/train_view_parallel.py bit-identical."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussian_splatting_b200.densify import DensificationStats, DensifyConfig, plan_adaptive_density_control
        from gaussian_splatting_b200.structs import Gaussians
        from gaussian_splatting_b200.view_parallel import all_reduce_statistics

        n = 3000
        g0 = torch.Generator().manual_seed(0)                      # replicated parameters
        P = dict(xyz=torch.randn(n, 3, generator=g0), quaternion=torch.randn(n, 4, generator=g0),
                 scale=torch.randn(n, 3, generator=g0) * 1.2 - 4.0, opacity=torch.randn(n, 1, generator=g0) * 2.0 - 1.0,
                 rgb=torch.randn(n, 3, generator=g0), sh=torch.randn(n, 3, 3, generator=g0) * 0.1)
        g = Gaussians(P["xyz"], P["rgb"], P["opacity"], P["scale"], P["quaternion"], P["sh"])
        gr = torch.Generator().manual_seed(100 + rank)             # this rank's own views
        stats = DensificationStats.__new__(DensificationStats)
        stats.grad_accum_count = torch.randint(0, 3, (n,), generator=gr, dtype=torch.int32)
        stats.uv_grad_accum = torch.rand(n, 2, generator=gr) * 1e-3 * (stats.grad_accum_count > 0).unsqueeze(1)
        stats.xyz_grad_accum = torch.rand(n, 3, generator=gr) * 1e-3
        local_count = stats.grad_accum_count.clone()
        all_reduce_statistics([stats.uv_grad_accum, stats.xyz_grad_accum, stats.grad_accum_count])
        torch.manual_seed(7)                                       # same random streams on every rank
        plan = plan_adaptive_density_control(g, stats, DensifyConfig(), 1200)
        sig = torch.cat([plan.src.double(), plan.clone_row.double(), plan.split_row.double(),
                         plan.xyz_add.reshape(-1).double(), plan.xyz_sub.reshape(-1).double()])
        lo, hi = sig.clone(), sig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        total = local_count.clone()
        dist.all_reduce(total)
        if rank == 0:
            out.put((bool(torch.equal(lo, hi)), bool(torch.equal(total, stats.grad_accum_count)), plan.info))
    finally:
        dist.destroy_process_group()


def test_densification_plan_is_identical_on_every_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_adc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    same_plan, counts_summed, info = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_plan and counts_summed
    assert info["deleted"] > 0 and info["cloned"] > 0 and info["split"] > 0, info
