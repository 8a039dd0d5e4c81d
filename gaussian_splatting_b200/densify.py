"""Per-step densification statistics of the reference's trainer, on device in one kernel.

The reference accumulates, after every optimizer step (splat_py/trainer.py:376-385),
    uv_grad = uv.grad.detach(); uv_grad[:, 0] *= K[0, 0]; uv_grad[:, 1] *= K[1, 1]
    self.uv_grad_accum[~culling_mask] += torch.abs(uv_grad)
    self.xyz_grad_accum += torch.abs(self.gaussians.xyz.grad.detach())
    self.grad_accum_count += (~culling_mask).int()
(two boolean-mask scatters — each a nonzero() with a host sync — and four elementwise kernels over [N]-sized
tensors); `adaptive_density_control` (:208-295) later turns them into the clone / split / delete masks.  Here the
fused rasterizer already knows the indices of the visible Gaussians (`state.vis_idx`, ascending — the order of
the compact uv), so the whole update is `gsr_densify_accumulate`, with no sync.
"""
from __future__ import annotations

import torch

from . import native


class DensificationStats:
    """The three accumulators of splat_py/trainer.py:51-66 (`reset_grad_accum`), same names and shapes."""

    def __init__(self, n_gaussians: int, device, dtype=torch.float32):
        self.uv_grad_accum = torch.zeros(n_gaussians, 2, dtype=dtype, device=device)
        self.xyz_grad_accum = torch.zeros(n_gaussians, 3, dtype=dtype, device=device)
        self.grad_accum_count = torch.zeros(n_gaussians, dtype=torch.int32, device=device)

    def accumulate(self, state, uv: torch.Tensor, xyz: torch.Tensor, K: torch.Tensor) -> None:
        """state: the `_ViewState` rasterize(..., return_state=True) returned; uv: the compact uv it returned (with
        `.grad` retained); xyz: the position parameter (with `.grad`); K: the camera's 3x3 intrinsic matrix.
        Like the reference, scales `uv.grad` by the focal lengths in place."""
        assert uv.grad is not None, "call uv.retain_grad() before backward()"
        native().densify_accumulate(state.vis_idx32, uv.grad, xyz.grad.contiguous(), K.contiguous(),
                                    self.uv_grad_accum, self.xyz_grad_accum, self.grad_accum_count)
