"""The optimizer-step oracle (oracle/adam_oracle.py) against torch.optim.Adam itself, on CPU."""
import numpy as np
import torch

from oracle import adam_oracle


def test_adam_oracle_matches_torch_adam():
    rng = np.random.default_rng(0)
    n = 4096
    p0 = rng.standard_normal(n).astype(np.float32)
    lr = 0.002 * 5.0
    pt = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([pt], lr=lr)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 8):
        g = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0, n)).astype(np.float32)
        g[rng.random(n) < 0.2] = 0.0  # culled gaussians have exactly zero gradients
        pt.grad = torch.tensor(g.copy())
        opt.step()
        p, m, v = adam_oracle.adam_step(p, g, m, v, lr, step)
        st = opt.state[pt]
        # torch's CPU kernels are free to fuse differently: a few ulp
        np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=3e-7, atol=1e-30)
        np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=3e-7, atol=1e-30)
        np.testing.assert_allclose(p, pt.detach().numpy(), rtol=1e-6, atol=1e-7)
