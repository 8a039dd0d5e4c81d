"""Adaptive density control on the flat buffers (gaussian_splatting_b200/densify.py): the PLAN is checked on the CPU
against a literal restatement of the reference's sequence (splat_py/trainer.py:114-295 + splat_py/
optimizer_manager.py:74-172: physically filter / concatenate every tensor after each of delete, clone, split), the
native APPLY pass is checked on the GPU (tests/test_densify_gpu.py)."""
import numpy as np
import pytest
import torch

from gaussian_splatting_b200.densify import DensificationStats, DensifyConfig, plan_adaptive_density_control
from gaussian_splatting_b200.structs import Gaussians
from gaussian_splatting_b200.utils import quaternion_to_rotation_torch

FIELDS = ("xyz", "quaternion", "scale", "opacity", "rgb", "sh")


def make_case(n, seed, device="cpu", n_rest=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    P = dict(xyz=r(n, 3), quaternion=r(n, 4), scale=r(n, 3) * 1.2 - 4.0, opacity=r(n, 1) * 2.0 - 1.0, rgb=r(n, 3),
             sh=r(n, 3, n_rest) * 0.1)
    M = {k: r(*v.shape) for k, v in P.items()}           # Adam exp_avg
    V = {k: r(*v.shape).abs() for k, v in P.items()}     # Adam exp_avg_sq
    uv_acc = r(n, 2).abs() * 1e-3
    xyz_acc = r(n, 3).abs() * 1e-3
    cnt = torch.randint(0, 5, (n,), generator=g, dtype=torch.int32)
    uv_acc[cnt == 0] = 0.0
    uv_acc[torch.rand(n, generator=g) < 0.03] = 0.0      # seen, but zero gradient
    to = lambda t: t.to(device)  # noqa: E731
    return ({k: to(v) for k, v in P.items()}, {k: to(v) for k, v in M.items()}, {k: to(v) for k, v in V.items()},
            to(uv_acc), to(xyz_acc), to(cnt))


def reference_sequence(P, M, V, uv_acc, xyz_acc, cnt, cfg, it):
    """The reference, literally: every tensor is filtered / concatenated after each step.  Returns new (P, M, V)."""
    from gaussian_splatting_b200.densify import _inverse_sigmoid

    P, M, V = dict(P), dict(M), dict(V)

    def filt(mask):                                     # trainer.py:114-121 + optimizer_manager.py:74-96
        nonlocal uv_acc, xyz_acc, cnt
        for k in FIELDS:
            P[k], M[k], V[k] = P[k][mask], M[k][mask], V[k][mask]
        uv_acc, xyz_acc, cnt = uv_acc[mask, :], xyz_acc[mask, :], cnt[mask]

    def append(new):                                    # structs.append + optimizer_manager.py:98-172
        for k in FIELDS:
            P[k] = torch.cat([P[k], new[k]], dim=0)
            M[k] = torch.cat([M[k], torch.zeros_like(new[k])], dim=0)
            V[k] = torch.cat([V[k], torch.zeros_like(new[k])], dim=0)

    keep_mask = (P["opacity"] > _inverse_sigmoid(cfg.delete_opacity_threshold)).squeeze(1)
    keep_mask &= ~(cnt == 0)
    keep_mask &= ~(torch.norm(uv_acc, dim=1) == 0.0)
    if int((~keep_mask).sum()) > 0 and cfg.use_delete:
        filt(keep_mask)
    if P["xyz"].shape[0] > cfg.max_gaussians:
        return P, M, V
    uv_grad_avg = uv_acc / cnt.unsqueeze(1).float()
    xyz_grad_avg = xyz_acc / cnt.unsqueeze(1).float()
    uv_norm = torch.norm(uv_grad_avg, dim=1)
    sf = (float(cfg.adaptive_control_end - it) / float(cfg.adaptive_control_end - cfg.adaptive_control_start) * 2.0
          if cfg.use_adaptive_fractional_densification else 1.0)
    if cfg.use_fractional_densification:
        uv_split_val = torch.quantile(uv_norm, 1.0 - (1.0 - cfg.uv_grad_percentile) * sf).item()
    else:
        uv_split_val = cfg.uv_grad_threshold
    densify_mask = uv_norm > uv_split_val
    scale_max = P["scale"].exp().max(dim=-1).values
    clone_mask = densify_mask & (scale_max <= cfg.clone_scale_threshold)
    if clone_mask.any() and cfg.use_clone:              # trainer.py:122-164
        new = {k: P[k][clone_mask].clone() for k in FIELDS}
        new["xyz"] -= xyz_grad_avg[clone_mask, :] * 0.01
        uv_acc = torch.cat([uv_acc, uv_acc[clone_mask, :]], dim=0)
        xyz_acc = torch.cat([xyz_acc, xyz_acc[clone_mask, :]], dim=0)
        cnt = torch.cat([cnt, cnt[clone_mask]], dim=0)
        append(new)
        densify_mask = torch.cat([densify_mask, densify_mask[clone_mask]], dim=0)
        scale_max = torch.cat([scale_max, scale_max[clone_mask]], dim=0)
    split_mask = densify_mask & (scale_max > cfg.clone_scale_threshold)
    scale_split = torch.quantile(scale_max, 1.0 - (1.0 - cfg.scale_norm_percentile) * sf).item()
    split_mask = split_mask | (scale_max > scale_split)
    if split_mask.any() and cfg.use_split:              # trainer.py:166-206
        s = cfg.num_split_samples
        new = {k: P[k][split_mask].clone().repeat(*([s] + [1] * (P[k].dim() - 1))) for k in FIELDS}
        rnd = torch.rand(int(split_mask.sum()) * s, 3, device=P["xyz"].device)
        rnd = rnd * torch.exp(new["scale"])
        new["quaternion"] = new["quaternion"] / torch.norm(new["quaternion"], dim=1, keepdim=True)
        rnd = torch.bmm(quaternion_to_rotation_torch(new["quaternion"]), rnd.unsqueeze(-1)).squeeze(-1)
        new["xyz"] += rnd
        new["scale"] = torch.log(torch.exp(new["scale"]) / cfg.split_scale_factor)
        filt(~split_mask)
        append(new)
    return P, M, V


def emulate_apply(plan, P, M, V):
    """What csrc/gsr_densify.cu does, with torch indexing (CPU stand-in for the native pass)."""
    if plan.is_identity():
        return dict(P), dict(M), dict(V)
    src = plan.src.long()
    fresh = torch.zeros(src.numel(), dtype=torch.bool, device=src.device)
    Po = {k: P[k][src].clone() for k in FIELDS}
    if plan.clone_row is not None:
        c = plan.clone_row.long()
        sel = c >= 0
        Po["xyz"][sel] = Po["xyz"][sel] - plan.xyz_sub[c[sel]]
        fresh |= sel
    if plan.split_row is not None:
        sr = plan.split_row.long()
        sel = sr >= 0
        Po["xyz"][sel] = Po["xyz"][sel] + plan.xyz_add[sr[sel]]
        Po["quaternion"][sel] = plan.q_set[sr[sel]]
        Po["scale"][sel] = plan.scale_set[sr[sel]]
        fresh |= sel
    Mo = {k: M[k][src].clone() for k in FIELDS}
    Vo = {k: V[k][src].clone() for k in FIELDS}
    for k in FIELDS:
        Mo[k][fresh] = 0.0
        Vo[k][fresh] = 0.0
    return Po, Mo, Vo


CASES = [
    dict(),                                                           # the reference's defaults (7k schedule)
    dict(use_adaptive_fractional_densification=False),
    dict(use_fractional_densification=False, use_adaptive_fractional_densification=False, uv_grad_threshold=4e-4),
    dict(use_delete=False),
    dict(use_clone=False),
    dict(use_split=False),
    dict(max_gaussians=100),                                          # early return after the delete
    dict(clone_scale_threshold=0.05, num_split_samples=3),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("it", [800, 3000])
def test_plan_reproduces_the_reference_sequence(case, it):
    cfg = DensifyConfig(**CASES[case])
    n = 4000
    P, M, V, uv_acc, xyz_acc, cnt = make_case(n, seed=17 + case)
    torch.manual_seed(5)
    Pr, Mr, Vr = reference_sequence(P, M, V, uv_acc, xyz_acc, cnt, cfg, it)
    g = Gaussians(P["xyz"], P["rgb"], P["opacity"], P["scale"], P["quaternion"], P["sh"])
    stats = DensificationStats.__new__(DensificationStats)
    stats.uv_grad_accum, stats.xyz_grad_accum, stats.grad_accum_count = uv_acc, xyz_acc, cnt
    torch.manual_seed(5)
    plan = plan_adaptive_density_control(g, stats, cfg, it)
    Po, Mo, Vo = emulate_apply(plan, P, M, V)
    assert plan.n_out == Pr["xyz"].shape[0], plan.info
    for k in FIELDS:
        assert torch.equal(Po[k], Pr[k]), (k, plan.info)
        assert torch.equal(Mo[k], Mr[k]) and torch.equal(Vo[k], Vr[k]), (k, plan.info)
    if case == 0:
        assert plan.info["deleted"] > 0 and plan.info["cloned"] > 0 and plan.info["split"] > 0, plan.info


def test_plan_identity_when_nothing_to_do():
    cfg = DensifyConfig(use_delete=False, use_clone=False, use_split=False)
    P, M, V, uv_acc, xyz_acc, cnt = make_case(100, seed=1)
    g = Gaussians(P["xyz"], P["rgb"], P["opacity"], P["scale"], P["quaternion"], P["sh"])
    stats = DensificationStats.__new__(DensificationStats)
    stats.uv_grad_accum, stats.xyz_grad_accum, stats.grad_accum_count = uv_acc, xyz_acc, cnt
    assert plan_adaptive_density_control(g, stats, cfg, 1000).is_identity()
