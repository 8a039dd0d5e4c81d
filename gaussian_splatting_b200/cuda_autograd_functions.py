"""The six autograd Functions of the reference's operator surface
(splat_py/cuda_autograd_functions.py:19-219), bound to the B200 library.

Forward argument order and backward arities are the reference's; outputs are allocated here and
filled in place by the native op (gradient buffers of RenderImage are zero-filled because the
native backward accumulates into them, as the reference's does).
"""
from __future__ import annotations

import torch

from . import native


def _like(t, *shape, zero=False):
    make = torch.zeros if zero else torch.empty
    return make(*shape, dtype=t.dtype, device=t.device)


class CameraPointProjection(torch.autograd.Function):
    """xyz_camera [N,3], K [3,3] -> uv [N,2]   (autograd:19-34)"""

    @staticmethod
    def forward(ctx, xyz_camera, K):
        uv = _like(xyz_camera, xyz_camera.shape[0], 2)
        native().camera_projection_cuda(xyz_camera, K, uv)
        ctx.save_for_backward(xyz_camera, K)
        return uv

    @staticmethod
    def backward(ctx, grad_uv):
        xyz_camera, K = ctx.saved_tensors
        grad_xyz = torch.zeros_like(xyz_camera)  # rows with z <= 0 are left at zero by the op
        native().camera_projection_backward_cuda(xyz_camera, K, grad_uv.contiguous(), grad_xyz)
        return grad_xyz, None


class ComputeSigmaWorld(torch.autograd.Function):
    """quaternion [N,4], scale [N,3] -> sigma_world [N,3,3]   (autograd:37-61)"""

    @staticmethod
    def forward(ctx, quaternion, scale):
        sigma = _like(quaternion, quaternion.shape[0], 3, 3)
        native().compute_sigma_world_cuda(quaternion, scale, sigma)
        ctx.save_for_backward(quaternion, scale)
        return sigma

    @staticmethod
    def backward(ctx, grad_sigma_world):
        quaternion, scale = ctx.saved_tensors
        gq, gs = torch.empty_like(quaternion), torch.empty_like(scale)
        native().compute_sigma_world_backward_cuda(quaternion, scale, grad_sigma_world.contiguous(), gq, gs)
        return gq, gs


class ComputeProjectionJacobian(torch.autograd.Function):
    """xyz_camera [N,3], K -> J [N,2,3]   (autograd:64-81)"""

    @staticmethod
    def forward(ctx, xyz_camera, K):
        J = _like(xyz_camera, xyz_camera.shape[0], 2, 3)
        native().compute_projection_jacobian_cuda(xyz_camera, K, J)
        ctx.save_for_backward(xyz_camera, K)
        return J

    @staticmethod
    def backward(ctx, grad_jacobian):
        xyz_camera, K = ctx.saved_tensors
        grad_xyz = torch.empty_like(xyz_camera)
        native().compute_projection_jacobian_backward_cuda(xyz_camera, K, grad_jacobian.contiguous(), grad_xyz)
        return grad_xyz, None


class ComputeConic(torch.autograd.Function):
    """sigma_world, J, camera_T_world -> conic [N,3] = [S00, S01+S10, S11]   (autograd:84-102)"""

    @staticmethod
    def forward(ctx, sigma_world, J, camera_T_world):
        conic = _like(sigma_world, J.shape[0], 3)
        native().compute_conic_cuda(sigma_world, J, camera_T_world, conic)
        ctx.save_for_backward(sigma_world, camera_T_world, J)
        return conic

    @staticmethod
    def backward(ctx, grad_conic):
        sigma_world, camera_T_world, J = ctx.saved_tensors
        g_sigma, g_J = torch.empty_like(sigma_world), torch.empty_like(J)
        native().compute_conic_backward_cuda(sigma_world, J, camera_T_world, grad_conic.contiguous(), g_sigma, g_J)
        return g_sigma, g_J, None


class PrecomputeRGBFromSH(torch.autograd.Function):
    """sh_coeffs [N,3(,K)], xyz [N,3], inverse(camera_T_world) -> rgb [N,3]   (autograd:105-127)

    No gradient flows to xyz (the reference drops the view-direction term too).
    """

    @staticmethod
    def forward(ctx, sh_coeffs, xyz, camera_T_world):
        rgb = _like(sh_coeffs, xyz.shape[0], 3)
        native().precompute_rgb_from_sh_cuda(xyz, sh_coeffs, camera_T_world, rgb)
        ctx.save_for_backward(xyz, camera_T_world)
        ctx.sh_shape = tuple(sh_coeffs.shape)  # kept on the host: no device round trip in backward
        return rgb

    @staticmethod
    def backward(ctx, grad_rgb):
        xyz, camera_T_world = ctx.saved_tensors
        grad_sh = _like(xyz, *ctx.sh_shape)
        native().precompute_rgb_from_sh_backward_cuda(xyz, camera_T_world, grad_rgb.contiguous(), grad_sh)
        return grad_sh, None, None


class RenderImage(torch.autograd.Function):
    """Tile compositing (autograd:130-219).  Returns image [H,W,3]; grads for rgb, opacity, uvs, conic."""

    @staticmethod
    def forward(ctx, rgb, opacity, uvs, conic, rays, splat_start_end_idx_by_tile_idx,
                sorted_gaussian_idx_by_splat_idx, image_size, background_rgb):
        h, w = int(image_size[0]), int(image_size[1])
        image = _like(rgb, h, w, 3, zero=True)
        num_splats_per_pixel = torch.zeros(h, w, dtype=torch.int32, device=rgb.device)
        final_weight_per_pixel = _like(rgb, h, w, zero=True)
        native().render_tiles_cuda(
            uvs, opacity, rgb, conic, rays, splat_start_end_idx_by_tile_idx,
            sorted_gaussian_idx_by_splat_idx, background_rgb, num_splats_per_pixel,
            final_weight_per_pixel, image,
        )
        ctx.save_for_backward(
            uvs, opacity, rgb, conic, rays, splat_start_end_idx_by_tile_idx,
            sorted_gaussian_idx_by_splat_idx, background_rgb, num_splats_per_pixel, final_weight_per_pixel,
        )
        return image

    @staticmethod
    def backward(ctx, grad_rendered_image):
        (uvs, opacity, rgb, conic, rays, ranges, sorted_idx, background_rgb, num_splats_per_pixel,
         final_weight_per_pixel) = ctx.saved_tensors
        grads = [torch.zeros_like(t) for t in (rgb, opacity, uvs, conic)]
        native().render_tiles_backward_cuda(
            uvs, opacity, rgb, conic, rays, ranges, sorted_idx, background_rgb, num_splats_per_pixel,
            final_weight_per_pixel, grad_rendered_image.contiguous(), *grads,
        )
        return (*grads, None, None, None, None, None)
