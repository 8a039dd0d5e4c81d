"""Deterministic synthetic scenes ("synth-v1") for parity tests and bench.py.

Everything is generated on the CPU from seeded torch generators and then moved to the target
device, so every box sees identical bits.  No dataset, no network.
"""
from __future__ import annotations

import math

import torch

from .structs import Camera, Gaussians

SH0 = 0.28209479177387814

RESOLUTIONS = {
    "4k": (2160, 3840, 2400.0),
    "1080p": (1080, 1920, 1200.0),
    "720p": (720, 1280, 800.0),
    "small": (192, 320, 200.0),
    "tiny": (64, 64, 60.0),
}


def make_camera(res="1080p", device="cpu", dtype=torch.float32):
    H, W, f = RESOLUTIONS[res]
    K = torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], dtype=dtype, device=device)
    return Camera(W, H, K)


def make_pose(view=0, n_views=1, device="cpu", dtype=torch.float32):
    """camera_T_world: yaw of (view-(V-1)/2)*3 degrees about the point (0,0,6); identity for V=1."""
    ang = math.radians((view - (n_views - 1) / 2.0) * 3.0)
    c, s = math.cos(ang), math.sin(ang)
    R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)
    pivot = torch.tensor([0.0, 0.0, 6.0], dtype=torch.float64)
    t = pivot - R @ pivot
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = R
    T[:3, 3] = t
    return T.to(dtype=dtype, device=device)


def make_gaussians(n, res="1080p", sh_degree=3, seed=0, device="cpu", requires_grad=False,
                   sigma_px=(1.2, 0.6, 0.3, 12.0), dtype=torch.float32):
    """N Gaussians filling (and overfilling by 25%) the view frustum of `make_camera(res)`.

    z ~ U[1,11]; centres uniform in 1.25x the image; per-axis pixel sigma LogNormal(ln mu, s)
    clipped to [lo, hi]; random (unnormalised) quaternions; opacity logit U[-2,4];
    DC colour U[0,1]/SH0; higher SH bands N(0, 0.05^2).
    """
    H, W, f = RESOLUTIONS[res]
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(n, generator=g, dtype=torch.float64) * 10.0 + 1.0
    tx = (torch.rand(n, generator=g, dtype=torch.float64) * 2.5 - 1.25) * (W / 2.0) / f
    ty = (torch.rand(n, generator=g, dtype=torch.float64) * 2.5 - 1.25) * (H / 2.0) / f
    xyz = torch.stack([z * tx, z * ty, z], dim=1)
    mu, s, lo, hi = sigma_px
    sig = torch.exp(torch.randn(n, 3, generator=g, dtype=torch.float64) * s + math.log(mu)).clamp(lo, hi)
    scale = torch.log(z[:, None] * sig / f)
    quaternion = torch.randn(n, 4, generator=g, dtype=torch.float64)
    opacity = torch.rand(n, 1, generator=g, dtype=torch.float64) * 6.0 - 2.0
    rgb = torch.rand(n, 3, generator=g, dtype=torch.float64) / SH0
    n_rest = (sh_degree + 1) ** 2 - 1
    sh = torch.randn(n, 3, n_rest, generator=g, dtype=torch.float64) * 0.05 if n_rest > 0 else None

    def prep(t):
        t = t.to(dtype=dtype, device=device).contiguous()
        return t.requires_grad_(True) if requires_grad else t

    return Gaussians(
        xyz=prep(xyz), rgb=prep(rgb), opacity=prep(opacity), scale=prep(scale), quaternion=prep(quaternion),
        sh=None if sh is None else prep(sh),
    )


def make_upstream_grad(res="1080p", seed=1, device="cpu", dtype=torch.float32):
    """G ~ N(0,1)/(3HW): the image gradient fed to backward()."""
    H, W, _ = RESOLUTIONS[res]
    g = torch.Generator().manual_seed(seed)
    G = torch.randn(H, W, 3, generator=g, dtype=torch.float64) / (3.0 * H * W)
    return G.to(dtype=dtype, device=device).contiguous()


DEFAULTS = dict(near_thresh=0.3, far_thresh=500.0, cull_mask_padding=100, mh_dist=3.0)
