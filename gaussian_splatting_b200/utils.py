"""Geometry helpers on the rasterization path (mirror of splat_py/utils.py:40-123)."""
from __future__ import annotations

import torch


def inverse_sigmoid_torch(x):
    """logit with the reference's clipping to [1e-4, 1-1e-4] (splat_py/utils.py:14-16)."""
    p = torch.clip(x, 1e-4, 1 - 1e-4)
    return torch.log(p / (1.0 - p))


def quaternion_to_rotation_torch(q):
    """[N,4] normalised (w,x,y,z) -> [N,3,3] (splat_py/utils.py:40-57)."""
    w, x, y, z = q.unbind(dim=1)
    rows = [
        1 - 2 * y**2 - 2 * z**2, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y,
        2 * x * y + 2 * w * z, 1 - 2 * x**2 - 2 * z**2, 2 * y * z - 2 * w * x,
        2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x**2 - 2 * y**2,
    ]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def transform_points_torch(pts, transform):
    """Apply a 4x4 rigid transform to [N,3] points (splat_py/utils.py:60-72).

    Same matmul formulation as the reference so the bits match its xyz_camera_frame; the
    NaN report of the reference (a host sync) is dropped — NaN points are culled downstream.
    """
    ones = torch.ones(pts.shape[0], 1, dtype=pts.dtype, device=pts.device)
    homog = torch.cat([pts, ones], dim=1)
    out = torch.matmul(transform, homog.unsqueeze(-1)).squeeze(-1)[:, :3]
    return out.contiguous()


def compute_rays(camera):
    """Unit ray per pixel in the camera frame, row-major [H*W,3] (splat_py/utils.py:75-109)."""
    K = camera.K
    u = torch.linspace(0, camera.width - 1, camera.width, dtype=K.dtype, device=K.device)
    v = torch.linspace(0, camera.height - 1, camera.height, dtype=K.dtype, device=K.device)
    v, u = torch.meshgrid(v, u, indexing="ij")
    u, v = u.flatten(), v.flatten()
    d = torch.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], torch.ones_like(u)], dim=-1)
    return d / torch.norm(d, dim=1, keepdim=True)


def compute_rays_in_world_frame(camera, camera_T_world):
    """Unit ray per pixel in the world frame, [H,W,3] contiguous (splat_py/utils.py:112-123)."""
    rays = compute_rays(camera)
    world_T_camera = torch.inverse(camera_T_world)
    rays = (world_T_camera[:3, :3] @ rays.T).T
    rays = rays / torch.norm(rays, dim=1, keepdim=True)
    return rays.reshape(camera.height, camera.width, 3).contiguous()
