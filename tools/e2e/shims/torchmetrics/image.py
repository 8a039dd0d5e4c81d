"""SSIM stand-in with the call signature the reference trainer uses (splat_py/trainer.py:24, :329, :371):

    ssim = StructuralSimilarityIndexMeasure(data_range=1.0).to(device)
    value = ssim(pred [B,C,H,W], target [B,C,H,W])          # scalar tensor, differentiable w.r.t. pred

Standard SSIM (Wang et al. 2004): 11x11 Gaussian window, sigma 1.5, K1 = 0.01, K2 = 0.03, per-channel
depthwise filtering, mean over the valid region.  The window is separable, so the five local moments are
filtered by ONE 11x1 and ONE 1x11 depthwise convolution over the stacked inputs (mathematically the 2-D window;
a 121-tap depthwise conv2d per moment costs ~10 ms per 1080p step and would dominate every arm).  All arms of
tools/e2e use this same shim, so the comparison between them is like for like.
"""
import torch
import torch.nn.functional as F


class StructuralSimilarityIndexMeasure(torch.nn.Module):
    def __init__(self, data_range=1.0, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
        super().__init__()
        self.c1 = (k1 * data_range) ** 2
        self.c2 = (k2 * data_range) ** 2
        x = torch.arange(kernel_size, dtype=torch.float32) - (kernel_size - 1) / 2
        g = torch.exp(-(x * x) / (2 * sigma * sigma))
        g = g / g.sum()
        self.register_buffer("window", g)

    def forward(self, pred, target):
        c = pred.shape[1]
        g = self.window.to(pred.dtype)
        stack = torch.cat([pred, target, pred * pred, target * target, pred * target], dim=1)  # [B, 5c, H, W]
        k = stack.shape[1]
        rows = F.conv2d(stack, g.view(1, 1, -1, 1).expand(k, 1, -1, 1), groups=k)
        blurred = F.conv2d(rows, g.view(1, 1, 1, -1).expand(k, 1, 1, -1), groups=k)
        mu_p, mu_t, pp, tt, pt = blurred.split(c, dim=1)
        var_p = pp - mu_p * mu_p
        var_t = tt - mu_t * mu_t
        cov = pt - mu_p * mu_t
        ssim = ((2 * mu_p * mu_t + self.c1) * (2 * cov + self.c2)) / (
            (mu_p * mu_p + mu_t * mu_t + self.c1) * (var_p + var_t + self.c2))
        return ssim.mean()
