"""CPU restatement of the optimizer step on the flat parameter buffer.  TEST INFRASTRUCTURE ONLY — imported by
tests/ (and nothing in the product path).

Follows torch.optim.Adam as the reference configures it (splat_py/optimizer_manager.py:13-44; torch/optim/adam.py
`_single_tensor_adam` / `_multi_tensor_adam`: lerp_, mul_ + addcmul_, sqrt / bias_correction2_sqrt + eps, addcdiv_),
in float32 with the fused multiply-adds torch's CUDA kernels compile to (emulated through float64, exact up to
double rounding).  Pinned against torch.optim.Adam itself in tests/test_adam_oracle.py.
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """One Adam step on float32 arrays (in place on copies; returns p, m, v).  `lr` is a scalar or per-element."""
    p, g, m, v = (np.array(x, f32) for x in (p, g, m, v))
    w1, w2, b2 = f32(1.0 - beta1), f32(1.0 - beta2), f32(beta2)
    bias1, bias2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    neg_step = (-(np.asarray(lr, np.float64)) / bias1).astype(f32)
    inv_b2s = f32(1.0) / f32(np.sqrt(bias2))
    m = _fma(np.broadcast_to(w1, g.shape), (g - m).astype(f32), m)
    v = _fma((w2 * g).astype(f32), g, (v * b2).astype(f32))
    d = ((np.sqrt(v).astype(f32) * inv_b2s).astype(f32) + f32(eps)).astype(f32)
    p = _fma(np.broadcast_to(neg_step, g.shape).astype(f32), (m / d).astype(f32), p)
    return p, m, v
