"""render_depth(): mirror of splat_py/depth.py:17-88 (no gradients)."""
from __future__ import annotations

import torch

from . import native
from .cuda_autograd_functions import (
    CameraPointProjection,
    ComputeConic,
    ComputeProjectionJacobian,
    ComputeSigmaWorld,
)
from .structs import Tiles
from .tile_culling import get_splats
from .utils import transform_points_torch


def render_depth(gaussians, alpha_threshold, camera_T_world, camera, near_thresh, cull_mask_padding, mh_dist):
    """Range to the first surface where accumulated alpha exceeds `alpha_threshold`; -1 where none."""
    with torch.no_grad():
        xyz_cam = transform_points_torch(gaussians.xyz, camera_T_world)
        uv = CameraPointProjection.apply(xyz_cam, camera.K)
        keep = ~(
            (xyz_cam[:, 2] < near_thresh)
            | (uv[:, 0] < -1 * cull_mask_padding) | (uv[:, 0] > camera.width + cull_mask_padding)
            | (uv[:, 1] < -1 * cull_mask_padding) | (uv[:, 1] > camera.height + cull_mask_padding)
        )
        uv, xyz_cam = uv[keep, :].contiguous(), xyz_cam[keep, :].contiguous()
        opacity = torch.sigmoid(gaussians.opacity[keep]).contiguous()
        sigma_world = ComputeSigmaWorld.apply(gaussians.quaternion[keep, :].contiguous(),
                                              gaussians.scale[keep, :].contiguous())
        J = ComputeProjectionJacobian.apply(xyz_cam, camera.K)
        conic = ComputeConic.apply(sigma_world, J, camera_T_world)
        tiles = Tiles(camera.height, camera.width, uv.device)
        sorted_idx, tile_ranges = get_splats(uv, tiles, conic, xyz_cam, mh_dist)
        depth_image = torch.full((camera.height, camera.width, 1), -1.0, dtype=torch.float32, device=uv.device)
        native().render_depth_cuda(xyz_cam, uv, opacity, conic, tile_ranges, sorted_idx, alpha_threshold,
                                   depth_image)
        return depth_image
