"""SSIM stand-in with the call signature the reference trainer uses (splat_py/trainer.py:24, :329, :371):

    ssim = StructuralSimilarityIndexMeasure(data_range=1.0).to(device)
    value = ssim(pred [B,C,H,W], target [B,C,H,W])          # scalar tensor, differentiable w.r.t. pred

Standard SSIM (Wang et al. 2004): 11x11 Gaussian window, sigma 1.5, K1 = 0.01, K2 = 0.03, per-channel
depthwise filtering, mean over the valid region.  Both arms of tools/e2e/run_trainer.py use this same shim,
so the comparison between them is like for like.
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
        self.register_buffer("window", (g[:, None] * g[None, :])[None, None])

    def forward(self, pred, target):
        c = pred.shape[1]
        w = self.window.to(pred.dtype).expand(c, 1, -1, -1)
        blur = lambda t: F.conv2d(t, w, groups=c)  # noqa: E731
        mu_p, mu_t = blur(pred), blur(target)
        var_p = blur(pred * pred) - mu_p * mu_p
        var_t = blur(target * target) - mu_t * mu_t
        cov = blur(pred * target) - mu_p * mu_t
        ssim = ((2 * mu_p * mu_t + self.c1) * (2 * cov + self.c2)) / (
            (mu_p * mu_p + mu_t * mu_t + self.c1) * (var_p + var_t + self.c2))
        return ssim.mean()
