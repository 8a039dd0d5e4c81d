"""View-parallel optimisation step on N GPUs of one node (second half):

    rank r:  render view (step*N + r)  ->  L1 loss against its target  ->  backward
    all:     ONE all-reduce (NCCL, average) of the flat gradient bucket the fused backward produced
    all:     Adam step on the replicated Gaussians (identical on every rank)

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port P tools/e2e/view_parallel_step.py --gaussians 3000000 --res 1080p --steps 10
(or plain `python tools/e2e/view_parallel_step.py` for N = 1: the collective is skipped).

Prints one JSON line from rank 0 with CUDA-event times (max over ranks) of the three phases and of the whole
step, the bytes reduced, whether the bucket was zero-copy, and a consistency check (parameters identical on
all ranks after the steps).  Targets are renders of a perturbed copy of the scene, so the loss decreases; the
LR convention is the reference's (1 view per optimizer step -> here the MEAN gradient of N views per step).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=3_000_000)
    ap.add_argument("--res", default="1080p")
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--lr", type=float, default=2e-3, help="base learning rate (reference: 0.002)")
    ap.add_argument("--stages", action="store_true", help="also report the rasterizer's per-stage times of the last step")
    ap.add_argument("--optimizer", choices=["torch", "flat", "sharded"], default="torch",
                    help="torch: NCCL all-reduce + torch.optim.Adam(fused); flat: NCCL all-reduce + gsr_adam_step on the "
                         "flat buffer; sharded: ONE kernel per rank doing reduce-scatter + Adam + all-gather over "
                         "NVLink peer memory (symmetric memory), optimizer state sharded")
    a = ap.parse_args()

    real_stdout = os.fdopen(os.dup(1), "w")  # NCCL prints its version line on stdout: keep ours clean
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    from gaussian_splatting_b200 import synth
    from gaussian_splatting_b200.rasterize import rasterize
    from gaussian_splatting_b200.view_parallel import PARAM_FIELDS, GradientBucket, broadcast_parameters, views_of_rank

    g = synth.make_gaussians(a.gaussians, a.res, sh_degree=a.sh_degree, seed=0, device=dev, requires_grad=True)
    broadcast_parameters(g)  # identical bits everywhere (they already are: same seed)
    cam = synth.make_camera(a.res, device=dev)
    bg = torch.full((3,), 0.5, device=dev)
    # learning-rate multipliers of the reference's optimizer (splat_py/config.py: xyz 0.1, quaternion 2, scale 5,
    # opacity 10, rgb 2, sh 0.1 times base_lr)
    from gaussian_splatting_b200.flat_adam import REFERENCE_LR_MULTIPLIERS as mult, FlatAdam, ShardedFlatAdam, flatten_gaussians

    grads_sym = None
    if a.optimizer == "torch":
        opt = torch.optim.Adam([dict(params=[getattr(g, f)], lr=a.lr * mult[f]) for f in PARAM_FIELDS
                                if getattr(g, f, None) is not None], fused=True)
    elif a.optimizer == "flat":
        flat, ends, names = flatten_gaussians(g)
        opt = FlatAdam(flat, ends, [a.lr * mult[f] for f in names])
    else:
        import torch.distributed._symmetric_memory as symm_mem

        assert world > 1, "--optimizer sharded needs torchrun with at least 2 ranks"
        n_rest = 0 if g.sh is None else g.sh.shape[2]
        from gaussian_splatting_b200.flat_adam import section_ends

        total = section_ends(a.gaussians, n_rest)[-1]
        params_sym = symm_mem.empty(total, dtype=torch.float32, device=dev)
        grads_sym = symm_mem.empty(total, dtype=torch.float32, device=dev)
        grads_sym.zero_()
        flat, ends, names = flatten_gaussians(g, flat=params_sym)
        opt = ShardedFlatAdam(params_sym, grads_sym, ends, [a.lr * mult[f] for f in names])
    params = [getattr(g, f) for f in PARAM_FIELDS if getattr(g, f, None) is not None]

    # targets: the same scene with brighter colours, rendered once per view
    poses = [synth.make_pose(v, a.views, device=dev) for v in range(a.views)]
    with torch.no_grad():
        g.rgb.mul_(1.15)
        targets = [rasterize(g, T, cam, 0.3, 500.0, 100, 3.0, True, bg)[0].clone() for T in poses]
        g.rgb.div_(1.15)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_render, t_reduce, t_adam, t_step, losses = [], [], [], [], []
    zero_copy = None
    for it in range(a.warmup + a.steps):
        v = views_of_rank(it, rank, world, a.views)
        for p in params:
            p.grad = None
        e = [ev() for _ in range(4)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e[0].record()
        prof = [] if (a.stages and it == a.warmup + a.steps - 1) else None
        image, _, _, state = rasterize(g, poses[v], cam, 0.3, 500.0, 100, 3.0, True, bg, return_state=True,
                                       grad_out=grads_sym, profile=prof)
        loss = (image - targets[v]).abs().mean()
        loss.backward()
        e[1].record()
        if a.optimizer == "sharded":
            e[2].record()
            opt.step()  # barrier, reduce-scatter + Adam + all-gather in one kernel, barrier
            bucket_bytes, zero_copy = grads_sym.numel() * 4, True
        else:
            bucket = GradientBucket.adopt(state.grad_flat, g)
            bucket.all_reduce(average=True)
            e[2].record()
            if a.optimizer == "flat":
                opt.step(bucket)  # the bucket carries its layout; FlatAdam rejects a non-native one
            else:
                opt.step()
            bucket_bytes, zero_copy = bucket.nbytes(), bucket.zero_copy
        e[3].record()
        torch.cuda.synchronize()
        if prof is not None:
            stage_ms = {name: round(e0.elapsed_time(e1), 3) for name, e0, e1 in prof}
        if it >= a.warmup:
            t_render.append(e[0].elapsed_time(e[1]))
            t_reduce.append(e[1].elapsed_time(e[2]))
            t_adam.append(e[2].elapsed_time(e[3]))
            t_step.append(e[0].elapsed_time(e[3]))
            losses.append(float(loss))

    def reduce_max(xs):
        t = torch.tensor([sum(xs) / len(xs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # replicas must still be identical: compare a checksum of the parameters across ranks
    chk = torch.stack([p.detach().double().sum() for p in params])
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).all())
    mean_loss = torch.tensor([losses[0], losses[-1]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mean_loss)
        mean_loss /= world
    line = {
        "what": "view-parallel step: render fwd+bwd of one view per GPU, one all-reduce of the gradient bucket, fused Adam",
        "n_gpus": world, "gaussians": a.gaussians, "res": a.res, "sh_degree": a.sh_degree, "steps": a.steps,
        "ms_render_fwd_bwd": round(reduce_max(t_render), 3), "ms_all_reduce": round(reduce_max(t_reduce), 3),
        "ms_optimizer": round(reduce_max(t_adam), 3), "ms_step": round(reduce_max(t_step), 3),
        "views_per_second": round(world * 1e3 / reduce_max(t_step), 1),
        "optimizer": a.optimizer, "bucket_bytes": bucket_bytes, "bucket_zero_copy": bool(zero_copy),
        "replicas_identical_after_steps": same,
        "loss_first_last": [round(float(mean_loss[0]), 6), round(float(mean_loss[1]), 6)],
        "device": torch.cuda.get_device_name(local),
    }
    if a.stages:
        line["stage_ms_rank0_last_step"] = stage_ms
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
