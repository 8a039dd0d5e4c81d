"""Operator surface of the rasterizer as torch.autograd.Functions, bound to the B200 library.

Public names, forward argument order and backward arity follow the reference module of the same name
(splat_py/cuda_autograd_functions.py:19-219) so that callers and tests written against it run unchanged:

    CameraPointProjection(xyz_camera [N,3], K [3,3])                 -> uv [N,2]
    ComputeSigmaWorld(quaternion [N,4], scale [N,3])                 -> sigma_world [N,3,3]
    ComputeProjectionJacobian(xyz_camera [N,3], K [3,3])             -> J [N,2,3]
    ComputeConic(sigma_world, J, camera_T_world [4,4])               -> conic [N,3] = [S00, S01+S10, S11]
    PrecomputeRGBFromSH(sh_coeffs [N,3(,K)], xyz [N,3], world_T_cam) -> rgb [N,3]   (no gradient to xyz)
    RenderImage(rgb, opacity, uvs, conic, rays, tile_ranges, sorted_idx, image_size, background) -> image [H,W,3]

The five per-Gaussian operators share one shape: allocate the output, call `<op>_cuda`, save the inputs;
backward allocates the input gradients and calls `<op>_backward_cuda`.  They are generated from the table
below.  Output / gradient buffers are filled in place by the native op; the render backward ACCUMULATES, so
its buffers are zero-filled (as in the reference, :195-198).
"""
from __future__ import annotations

import torch

from . import native


def _new(like, shape, zero=False):
    alloc = torch.zeros if zero else torch.empty
    return alloc(tuple(shape), dtype=like.dtype, device=like.device)


def _per_gaussian_function(name, doc, *, fwd, bwd, n_inputs, out_shape, grad_slots, fwd_order=None,
                           bwd_order=None, zero_grads=False):
    """Build an autograd.Function for a per-gaussian operator.

    fwd / bwd        names of the native entry points
    n_inputs         number of tensor inputs of forward()
    out_shape        f(inputs) -> shape of the output
    grad_slots       indices of the inputs that receive a gradient (the others get None)
    fwd_order        order in which the inputs are handed to the native forward (default: as given)
    bwd_order        same for the native backward (default: forward order)
    zero_grads       zero-fill gradient buffers (the native op leaves some rows untouched)
    """
    fwd_order = tuple(range(n_inputs)) if fwd_order is None else fwd_order
    bwd_order = fwd_order if bwd_order is None else bwd_order

    def forward(ctx, *inputs):
        out = _new(inputs[0], out_shape(inputs))
        getattr(native(), fwd)(*[inputs[i] for i in fwd_order], out)
        ctx.save_for_backward(*inputs)
        return out

    def backward(ctx, grad_out):
        inputs = ctx.saved_tensors
        grads = [_new(inputs[i], inputs[i].shape, zero=zero_grads) for i in grad_slots]
        getattr(native(), bwd)(*[inputs[i] for i in bwd_order], grad_out.contiguous(), *grads)
        full = [None] * n_inputs
        for slot, g in zip(grad_slots, grads):
            full[slot] = g
        return tuple(full)

    return type(name, (torch.autograd.Function,),
                {"forward": staticmethod(forward), "backward": staticmethod(backward), "__doc__": doc})


CameraPointProjection = _per_gaussian_function(
    "CameraPointProjection", "pinhole projection u = fx x/z + cx, v = fy y/z + cy (reference :19-34)",
    fwd="camera_projection_cuda", bwd="camera_projection_backward_cuda", n_inputs=2,
    out_shape=lambda a: (a[0].shape[0], 2), grad_slots=(0,),
    zero_grads=True,  # rows with z <= 0 are left untouched by the native backward
)

ComputeSigmaWorld = _per_gaussian_function(
    "ComputeSigmaWorld", "Sigma = R(q/|q|) diag(exp(2 s)) R^T (reference :37-61)",
    fwd="compute_sigma_world_cuda", bwd="compute_sigma_world_backward_cuda", n_inputs=2,
    out_shape=lambda a: (a[0].shape[0], 3, 3), grad_slots=(0, 1),
)

ComputeProjectionJacobian = _per_gaussian_function(
    "ComputeProjectionJacobian", "J = d(u,v)/d(x,y,z) of the pinhole projection (reference :64-81)",
    fwd="compute_projection_jacobian_cuda", bwd="compute_projection_jacobian_backward_cuda", n_inputs=2,
    out_shape=lambda a: (a[0].shape[0], 2, 3), grad_slots=(0,),
)

ComputeConic = _per_gaussian_function(
    "ComputeConic", "2-D covariance (J W) Sigma (J W)^T as [S00, S01+S10, S11] (reference :84-102)",
    fwd="compute_conic_cuda", bwd="compute_conic_backward_cuda", n_inputs=3,
    out_shape=lambda a: (a[1].shape[0], 3), grad_slots=(0, 1),
)


class PrecomputeRGBFromSH(torch.autograd.Function):
    """Per-gaussian SH -> RGB in the direction camera->gaussian, scaled by 1/SH_0 (reference :105-127)."""

    @staticmethod
    def forward(ctx, sh_coeffs, xyz, camera_T_world):
        rgb = _new(sh_coeffs, (xyz.shape[0], 3))
        native().precompute_rgb_from_sh_cuda(xyz, sh_coeffs, camera_T_world, rgb)
        ctx.save_for_backward(xyz, camera_T_world)
        ctx.coeff_shape = tuple(sh_coeffs.shape)  # host-side: the backward needs no device round trip
        return rgb

    @staticmethod
    def backward(ctx, grad_rgb):
        xyz, camera_T_world = ctx.saved_tensors
        grad_sh = _new(xyz, ctx.coeff_shape)
        native().precompute_rgb_from_sh_backward_cuda(xyz, camera_T_world, grad_rgb.contiguous(), grad_sh)
        return grad_sh, None, None


class RenderImage(torch.autograd.Function):
    """Depth-ordered alpha compositing of the tile lists (reference :130-219)."""

    _N_NONDIFF = 5  # rays, tile ranges, sorted idx, image size, background

    @staticmethod
    def forward(ctx, rgb, opacity, uvs, conic, rays, splat_start_end_idx_by_tile_idx,
                sorted_gaussian_idx_by_splat_idx, image_size, background_rgb):
        height, width = int(image_size[0]), int(image_size[1])
        image = _new(rgb, (height, width, 3), zero=True)
        n_walked = torch.zeros(height, width, dtype=torch.int32, device=rgb.device)
        last_weight = _new(rgb, (height, width), zero=True)
        tensors = (uvs, opacity, rgb, conic, rays, splat_start_end_idx_by_tile_idx,
                   sorted_gaussian_idx_by_splat_idx, background_rgb, n_walked, last_weight)
        native().render_tiles_cuda(*tensors, image)
        ctx.save_for_backward(*tensors)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        tensors = ctx.saved_tensors
        uvs, opacity, rgb, conic = tensors[:4]
        grads = [torch.zeros_like(t) for t in (rgb, opacity, uvs, conic)]
        native().render_tiles_backward_cuda(*tensors, grad_image.contiguous(), *grads)
        return (*grads, *([None] * RenderImage._N_NONDIFF))
