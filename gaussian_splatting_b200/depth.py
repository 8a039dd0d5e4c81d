"""render_depth(): range image of the first opaque surface (mirror of splat_py/depth.py:17-88, no gradients)."""
from __future__ import annotations

import torch

from . import native
from .rasterize import project_and_bin


def render_depth(gaussians, alpha_threshold, camera_T_world, camera, near_thresh, cull_mask_padding, mh_dist):
    """[H,W,1] float32: distance from the camera to the gaussian at which the accumulated alpha first exceeds
    `alpha_threshold` (src/depth.cu:57-113); -1 where it never does.  The reference applies no far cull here."""
    with torch.no_grad():
        s = project_and_bin(gaussians, camera_T_world, camera, near_thresh, float("inf"), cull_mask_padding, mh_dist)
        depth_image = torch.full((camera.height, camera.width, 1), -1.0, dtype=torch.float32, device=s.uv.device)
        native().render_depth_cuda(s.xyz_cam, s.uv, s.opacity, s.conic, s.tile_ranges, s.sorted_idx,
                                   alpha_threshold, depth_image)
        return depth_image
