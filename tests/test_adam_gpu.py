"""Optimizer step on the flat parameter buffer (csrc/gsr_adam.cu through the C ABI) — needs a B200.

Checked against (1) the CPU oracle, (2) torch.optim.Adam on the GPU with the reference's parameter groups
(splat_py/optimizer_manager.py:13-44) — same values expected, bit for bit where torch's kernels fuse the same way —
and (3) end to end: rasterize -> backward -> FlatAdam.step(state.grad_flat) equals the torch optimizer path."""
import numpy as np
import pytest
import torch

from gaussian_splatting_b200 import synth
from gaussian_splatting_b200.flat_adam import FIELDS, REFERENCE_LR_MULTIPLIERS, FlatAdam, flatten_gaussians, section_ends
from gaussian_splatting_b200.rasterize import rasterize
from oracle import adam_oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_flat_layout_matches_backward():
    ends = section_ends(1001, 15)  # N not a multiple of 4: sections are padded to 16 bytes
    assert ends == [3004, 7008, 10012, 11016, 14020, 59068] and all(e % 4 == 0 for e in ends)
    assert section_ends(8, 0) == [24, 56, 80, 88, 112]


@pytest.mark.parametrize("n", [1001, 40000])
def test_flat_adam_vs_oracle_and_torch(n):
    rng = np.random.default_rng(1)
    ends = section_ends(n, 15)
    lrs = [0.002 * REFERENCE_LR_MULTIPLIERS[f] for f in FIELDS]
    total = ends[-1]
    p0 = rng.standard_normal(total).astype(np.float32)
    flat = torch.tensor(p0, device=dev())
    opt = FlatAdam(flat, ends, lrs)
    starts = [0] + ends[:-1]
    lr_el = np.zeros(total)
    for s, e, lr in zip(starts, ends, lrs):
        lr_el[s:e] = lr
    # torch reference: one parameter group per section, as the reference's OptimizerManager builds them
    tparams = [torch.tensor(p0[s:e], device=dev(), requires_grad=True) for s, e in zip(starts, ends)]
    topt = torch.optim.Adam([dict(params=[t], lr=lr) for t, lr in zip(tparams, lrs)])
    p, m, v = p0.copy(), np.zeros(total, np.float32), np.zeros(total, np.float32)
    for step in range(1, 6):
        g = (rng.standard_normal(total) * 10.0 ** rng.uniform(-7, 0, total)).astype(np.float32)
        g[rng.random(total) < 0.25] = 0.0
        gt = torch.tensor(g, device=dev())
        opt.step(gt)
        for t, s, e in zip(tparams, starts, ends):
            t.grad = gt[s:e].clone()
        topt.step()
        p, m, v = adam_oracle.adam_step(p, g, m, v, lr_el, step)
    mine = flat.cpu().numpy()
    ref = torch.cat([t.detach() for t in tparams]).cpu().numpy()
    assert np.abs(mine - p).max() <= 1e-6 * np.abs(p).max()
    np.testing.assert_allclose(opt.m.cpu().numpy(), m, rtol=2e-7, atol=1e-35)
    np.testing.assert_allclose(opt.v.cpu().numpy(), v, rtol=2e-7, atol=1e-35)
    diff = int((mine.view(np.int32) != ref.view(np.int32)).sum())
    assert np.abs(mine - ref).max() <= 1e-6 * np.abs(ref).max(), np.abs(mine - ref).max()
    print(f"n={n}: {diff}/{total} parameters differ bitwise from torch.optim.Adam (max abs {np.abs(mine - ref).max():.3e})")


def test_training_steps_match_torch_optimizer():
    """rasterize -> L1 loss -> backward -> optimizer, five steps: FlatAdam on the flat buffers vs torch.optim.Adam
    on separate tensors, same scene and views."""
    res, n = "tiny", 3000

    def make():
        return synth.make_gaussians(n, res, sh_degree=3, seed=0, device=dev(), requires_grad=True,
                                    sigma_px=(2.0, 0.5, 0.5, 8.0))

    cam = synth.make_camera(res, device=dev())
    bg = torch.full((3,), 0.5, device=dev())
    target = torch.rand(cam.height, cam.width, 3, device=dev(), generator=torch.Generator(device=dev()).manual_seed(3))
    ga, gb = make(), make()
    flat, ends, names = flatten_gaussians(ga)
    lrs = [0.002 * REFERENCE_LR_MULTIPLIERS[f] for f in names]
    fa = FlatAdam(flat, ends, lrs)
    tb = torch.optim.Adam([dict(params=[getattr(gb, f)], lr=lr) for f, lr in zip(names, lrs)])
    for it in range(5):
        T = synth.make_pose(it % 3, 3, device=dev())
        for g in (ga, gb):
            for f in names:
                getattr(g, f).grad = None
        img, _, _, st = rasterize(ga, T, cam, 0.3, 500.0, 100, 3.0, True, bg, return_state=True)
        (img - target).abs().mean().backward()
        fa.step(st.grad_flat)
        img2, _, _ = rasterize(gb, T, cam, 0.3, 500.0, 100, 3.0, True, bg)
        (img2 - target).abs().mean().backward()
        tb.step()
    for f in names:
        a, b = getattr(ga, f).detach(), getattr(gb, f).detach()
        # gradients carry the rasterizer's atomics noise (1e-7 relative); Adam's normalisation turns a sign flip of
        # a ~0 gradient into a full +-lr step, so compare with an absolute tolerance of a few steps
        lr = lrs[names.index(f)]
        frac = float(((a - b).abs() > 1e-6 + 1e-5 * b.abs()).float().mean())
        assert float((a - b).abs().max()) <= 2 * 5 * lr + 1e-6 and frac < 0.01, (f, float((a - b).abs().max()), frac)


def test_densification_statistics_match_reference_ops():
    """gsr_densify_accumulate vs the reference's update (splat_py/trainer.py:376-385) written with torch ops."""
    from gaussian_splatting_b200.densify import DensificationStats

    res, n = "small", 20000
    cam = synth.make_camera(res, device=dev())
    bg = torch.full((3,), 0.5, device=dev())
    G = synth.make_upstream_grad(res, device=dev())
    g = synth.make_gaussians(n, res, sh_degree=1, seed=2, device=dev(), requires_grad=True)
    stats = DensificationStats(n, dev())
    ref_uv, ref_xyz = torch.zeros(n, 2, device=dev()), torch.zeros(n, 3, device=dev())
    ref_cnt = torch.zeros(n, dtype=torch.int32, device=dev())
    for it in range(3):
        T = synth.make_pose(it, 3, device=dev())
        for f in ("xyz", "rgb", "opacity", "scale", "quaternion", "sh"):
            getattr(g, f).grad = None
        image, mask, uv, st = rasterize(g, T, cam, 0.3, 500.0, 100, 3.0, True, bg, return_state=True)
        uv.retain_grad()
        image.backward(G)
        # reference update, on a copy of uv.grad (both scale it in place)
        uv_grad = uv.grad.detach().clone()
        uv_grad[:, 0] = uv_grad[:, 0] * cam.K[0, 0]
        uv_grad[:, 1] = uv_grad[:, 1] * cam.K[1, 1]
        ref_uv[~mask] += torch.abs(uv_grad)
        ref_xyz += torch.abs(g.xyz.grad.detach())
        ref_cnt += (~mask).int()
        stats.accumulate(st, uv, g.xyz, cam.K)
        assert torch.equal(uv.grad, uv_grad)  # scaled in place, like the reference
    assert torch.equal(stats.uv_grad_accum, ref_uv) and torch.equal(stats.xyz_grad_accum, ref_xyz)
    assert torch.equal(stats.grad_accum_count, ref_cnt)


def _sharded_split(n, world):
    per = ((n // 4 + world - 1) // world) * 4
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_adam_kernel_emulated_ranks_on_one_gpu(world):
    """gsr_adam_step_sharded (reduce-scatter + Adam + all-gather in one kernel, csrc/gsr_adam.cu) with `world`
    EMULATED ranks on one device: every "rank" owns its own parameter and gradient buffer (the peer pointers then
    simply point at the same device), each rank's kernel is launched in turn.  Must equal: average the gradients
    over ranks (rank order), then FlatAdam on the average — and leave every replica bit-identical.  This is the
    single-GPU guard of the kernel the 2-GPU test below runs over NVLink."""
    n_g = 5001
    ends = section_ends(n_g, 15)
    lrs = [0.002 * REFERENCE_LR_MULTIPLIERS[f] for f in FIELDS]
    total = ends[-1]
    gen = torch.Generator(device=dev()).manual_seed(11)
    p0 = torch.randn(total, device=dev(), generator=gen)
    params = [p0.clone() for _ in range(world)]
    grads = [torch.empty(total, device=dev()) for _ in range(world)]
    split = _sharded_split(total, world)
    ms = [torch.zeros(hi - lo, device=dev()) for lo, hi in split]
    vs = [torch.zeros(hi - lo, device=dev()) for lo, hi in split]
    ref_p = p0.clone()
    ref = FlatAdam(ref_p, ends, lrs)
    ext = __import__("gaussian_splatting_b200").native()
    for step in range(1, 5):
        for r in range(world):
            grads[r].copy_(torch.randn(total, device=dev(), generator=gen) * 10.0 ** (-3.0 * torch.rand(total, device=dev(), generator=gen)))
        mean = grads[0].clone()
        for r in range(1, world):
            mean += grads[r]          # rank order, like the kernel
        mean *= 1.0 / world   # the kernel scales by the float reciprocal
        ref.step(mean)
        gp, pp = [int(t.data_ptr()) for t in grads], [int(t.data_ptr()) for t in params]
        for r, (lo, hi) in enumerate(split):
            ext.adam_step_sharded(lo, hi, gp, pp, r, ms[r], vs[r], ends, lrs, 0.9, 0.999, 1e-8, step)
        torch.cuda.synchronize()
        for r in range(1, world):
            assert torch.equal(params[r], params[0]), f"replica {r} differs after step {step}"
        err = float((params[0] - ref_p).abs().max())
        assert err <= 1e-6 * float(ref_p.abs().max()), (step, err)
    m_all, v_all = torch.cat(ms), torch.cat(vs)
    assert float((m_all - ref.m).abs().max()) <= 1e-6 * float(ref.m.abs().max())
    assert float((v_all - ref.v).abs().max()) <= 1e-6 * float(ref.v.abs().max())


def _sharded_worker(rank, world, port, q):
    import os

    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    d = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=d)
    try:
        from gaussian_splatting_b200.flat_adam import ShardedFlatAdam

        n_g = 20001
        ends = section_ends(n_g, 15)
        lrs = [0.002 * REFERENCE_LR_MULTIPLIERS[f] for f in FIELDS]
        total = ends[-1]
        p_sym = symm_mem.empty(total, dtype=torch.float32, device=d)
        g_sym = symm_mem.empty(total, dtype=torch.float32, device=d)
        gen = torch.Generator(device=d).manual_seed(5)       # same parameters on every rank
        p_sym.copy_(torch.randn(total, device=d, generator=gen))
        opt = ShardedFlatAdam(p_sym, g_sym, ends, lrs)
        ref_p = p_sym.clone()
        ref = FlatAdam(ref_p, ends, lrs)
        worst = 0.0
        for step in range(4):
            grank = torch.Generator(device=d).manual_seed(100 + 10 * step + rank)  # a different "view" per rank
            g_sym.copy_(torch.randn(total, device=d, generator=grank) * 1e-2)
            mean = g_sym.clone()
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)  # NCCL path: all-reduce + flat Adam
            mean /= world
            ref.step(mean)
            opt.step()
            torch.cuda.synchronize()
            worst = max(worst, float((p_sym - ref_p).abs().max()) / float(ref_p.abs().max()))
        gathered = [torch.empty_like(p_sym) for _ in range(world)]
        dist.all_gather(gathered, p_sym.clone())
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            q.put((worst, same))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NVLink peer memory)")
def test_sharded_adam_matches_allreduce_plus_flat():
    """ShardedFlatAdam over real symmetric memory on 2 GPUs: same parameters as NCCL all-reduce(avg) + FlatAdam
    (summation order differs from NCCL's by at most rounding), replicas bit-identical."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same, "parameter replicas differ"
    assert worst < 1e-6, worst
