"""ORACLE (test infrastructure, not product code): numpy restatement of the scheduler math.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows, in float32 with one rounding per operation (as eager PyTorch evaluates them):
  cosine tables   src/training/improved_diffusion_trainer.py:25-48
  q_sample        src/training/improved_diffusion_trainer.py:55-58 (+ clamp :363)
  ddpm step       src/training/improved_diffusion_trainer.py:543-567
  linear tables   src/training/final_trainer.py:22-40
  posterior step  src/training/final_trainer.py:52-71
  SmoothL1(0.1)   torch.nn.SmoothL1Loss, improved_diffusion_trainer.py:300
numpy's cos differs from torch's in the last ulp for a few entries, so the *tables* are pinned against the
reference through tests/golden/scheduler_tables.npz and the element-wise formulas are checked bit-exactly
given identical tables.  Pinned by oracle/make_golden.py.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def cosine_tables(num_timesteps=1000, beta_start=1e-4, beta_end=0.02, s=0.008):
    x = np.linspace(0, num_timesteps, num_timesteps + 1, dtype=f32)
    inner = ((x / f32(num_timesteps)) + f32(s)) / f32(1 + s) * f32(np.pi) * f32(0.5)
    abar = np.cos(inner.astype(f32)).astype(f32) ** 2
    abar = (abar / abar[0]).astype(f32)
    betas = (f32(1) - (abar[1:] / abar[:-1]).astype(f32)).astype(f32)
    betas = np.clip(betas, f32(beta_start), f32(beta_end)).astype(f32)
    alphas = (f32(1) - betas).astype(f32)
    alphas_cumprod = np.cumprod(alphas, dtype=f32)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": alphas_cumprod,
        "sqrt_alphas_cumprod": np.maximum(np.sqrt(alphas_cumprod), f32(1e-8)),
        "sqrt_one_minus_alphas_cumprod": np.maximum(np.sqrt(f32(1) - alphas_cumprod), f32(1e-8)),
    }


def q_sample(x0, noise, t, sqrt_ac, sqrt_1mac, clamp=None):
    x0 = x0.astype(f32)
    if clamp is not None:
        x0 = np.clip(x0, f32(-clamp), f32(clamp))
    shape = (-1,) + (1,) * (x0.ndim - 1)
    a = sqrt_ac[t].reshape(shape).astype(f32)
    b = sqrt_1mac[t].reshape(shape).astype(f32)
    out = ((a * x0).astype(f32) + (b * noise.astype(f32)).astype(f32)).astype(f32)
    if not np.isfinite(out).all():
        out = (x0 + (f32(0.1) * noise.astype(f32)).astype(f32)).astype(f32)
    return out


def ddpm_step(x, eps, z, t, betas, alphas, alphas_cumprod):
    c1 = (f32(1.0) / np.sqrt(alphas[t])).astype(f32)
    c2 = (betas[t] / np.sqrt(f32(1) - alphas_cumprod[t])).astype(f32)
    out = (c1 * (x - (c2 * eps).astype(f32)).astype(f32)).astype(f32)
    if z is not None:
        out = (out + (np.sqrt(betas[t]).astype(f32) * z).astype(f32)).astype(f32)
    return out


def posterior_step(x, eps, z, t, betas, sqrt_recip_alphas, sqrt_1mac, posterior_variance):
    mean = (sqrt_recip_alphas[t] * (x - ((betas[t] * eps).astype(f32) / sqrt_1mac[t]).astype(f32)).astype(f32)).astype(f32)
    if t > 0 and z is not None:
        return (mean + (np.sqrt(posterior_variance[t]).astype(f32) * z).astype(f32)).astype(f32)
    return mean


def smooth_l1(pred, target, beta=0.1):
    d = pred.astype(np.float64) - target.astype(np.float64)
    ad = np.abs(d)
    loss = np.where(ad < beta, 0.5 * d * d / beta, ad - 0.5 * beta)
    grad = np.where(ad < beta, d / beta, np.sign(d)) / d.size
    return float(loss.mean()), grad.astype(f32)
