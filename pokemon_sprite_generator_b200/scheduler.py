"""Drop-in noise schedulers backed by the fused CUDA kernels.

`NoiseScheduler` mirrors src/training/improved_diffusion_trainer.py:22-74 (clipped-cosine schedule, the one
stage-2 training uses): same constructor, same five tensor attributes, `add_noise(x_0, noise, timesteps)`,
`to(device)`.  `LinearNoiseScheduler` mirrors src/training/final_trainer.py:19-82 (linear betas + posterior
variance, used by FinalPokemonGenerator).  The tables are built with the same torch ops in the same order as
the reference so they are bit-identical to it in the same process (tests/test_scheduler.py).

The element-wise work (timestep gather, sqrt(abar)*x0 + sqrt(1-abar)*eps, reverse step) runs in
csrc/diffusion_ops.cu; the arithmetic there is un-contracted so results equal eager PyTorch bit for bit.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class NoiseScheduler:
    """Clipped-cosine DDPM schedule (reference: improved_diffusion_trainer.py:22-74)."""

    def __init__(self, num_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02):
        self.num_timesteps = num_timesteps
        self.betas = self._cosine_beta_schedule(num_timesteps, beta_start, beta_end).float()
        self.alphas = (1.0 - self.betas).float()
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0).float()
        self.sqrt_alphas_cumprod = torch.clamp(torch.sqrt(self.alphas_cumprod).float(), min=1e-8)
        self.sqrt_one_minus_alphas_cumprod = torch.clamp(torch.sqrt(1.0 - self.alphas_cumprod).float(), min=1e-8)
        self._build_step_tables()
        self._flag = None

    @staticmethod
    def _cosine_beta_schedule(timesteps, beta_start, beta_end, s=0.008):
        # reference :41-48 -- fp32 throughout, then clip to [beta_start, beta_end]
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float32)
        abar = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
        abar = abar / abar[0]
        betas = 1 - (abar[1:] / abar[:-1])
        return torch.clip(betas, beta_start, beta_end)

    def _build_step_tables(self):
        # Coefficients of ddpm_sample (reference :543-545,554-555,562), one fp32 entry per timestep, evaluated
        # exactly as the reference does at step time: 0-dim fp32 tensor ops, one timestep at a time.  (torch's
        # CPU sqrt is not correctly rounded, so a CUDA-resident reference may differ from these by 1 ulp; the
        # golden vectors come from the reference run on CPU.)
        c1, c2, sg = [], [], []
        for t in range(self.num_timesteps):
            alpha_t, abar_t, beta_t = self.alphas[t], self.alphas_cumprod[t], self.betas[t]
            c1.append(1.0 / torch.sqrt(alpha_t))
            c2.append(beta_t / torch.sqrt(1 - abar_t))
            sg.append(torch.sqrt(beta_t))
        self.step_coef1, self.step_coef2, self.step_sigma = torch.stack(c1), torch.stack(c2), torch.stack(sg)

    _TABLES = ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
               "step_coef1", "step_coef2", "step_sigma")

    def to(self, device):
        for name in self._TABLES:
            setattr(self, name, getattr(self, name).to(device, dtype=torch.float32).contiguous())
        return self

    # -- q_sample ---------------------------------------------------------------------------------
    def add_noise(self, x_0: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor, clamp: float | None = None,
                  check_finite: bool = True) -> torch.Tensor:
        """sqrt(abar_t) * x_0 + sqrt(1 - abar_t) * noise, with the reference's NaN/Inf fallback applied on device.

        `clamp=3.0` additionally fuses the trainer's `torch.clamp(latent, -3, 3)` (reference :363).
        """
        if not x_0.is_cuda:
            raise L.PsgError("NoiseScheduler.add_noise: CUDA tensors required (no CPU fallback)")
        self.to(x_0.device)
        x_0 = x_0.contiguous().float()
        noise = noise.contiguous().float()
        t = timesteps.to(device=x_0.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(x_0)
        b = x_0.shape[0]
        n_per = x_0.numel() // max(b, 1)
        flag = None
        if check_finite:
            if self._flag is None or self._flag.device != x_0.device:
                self._flag = torch.zeros(1, dtype=torch.int32, device=x_0.device)
            else:
                self._flag.zero_()
            flag = self._flag
        L.call("psg_q_sample", L.ptr(x_0), L.ptr(noise), L.ptr(t), L.ptr(self.sqrt_alphas_cumprod),
               L.ptr(self.sqrt_one_minus_alphas_cumprod), L.ptr(out), C.c_int(b), C.c_int(n_per),
               C.c_int(self.num_timesteps), C.c_int(1 if clamp is not None else 0),
               C.c_float(-(clamp or 0.0)), C.c_float(clamp or 0.0), L.ptr(flag), L.stream_ptr())
        return out

    # -- reverse step -----------------------------------------------------------------------------
    def ddpm_step(self, x_t: torch.Tensor, predicted_noise: torch.Tensor, t: int, noise: torch.Tensor | None) -> torch.Tensor:
        """One step of ddpm_sample (reference :543-567): (x - b/sqrt(1-abar) eps)/sqrt(a) [+ sqrt(b) z]."""
        self.to(x_t.device)
        x_t = x_t.contiguous()
        eps = predicted_noise.contiguous()
        out = torch.empty_like(x_t)
        L.call("psg_ddpm_step", L.ptr(x_t), L.ptr(eps), L.ptr(noise.contiguous() if noise is not None else None), L.ptr(out),
               C.c_longlong(x_t.numel()), C.c_int(0), L.ptr(self.step_coef1), L.ptr(self.step_coef2), L.ptr(self.step_sigma),
               L.ptr(None), C.c_int(int(t)), C.c_int(self.num_timesteps), L.stream_ptr())
        return out


class LinearNoiseScheduler:
    """Linear-beta schedule with posterior variance (reference: src/training/final_trainer.py:19-82)."""

    def __init__(self, num_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02):
        self.num_timesteps = num_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas = torch.sqrt(1.0 / self.alphas)
        posterior_variance = self.betas * (1.0 - torch.cat([torch.tensor([1.0]), self.alphas_cumprod[:-1]])) / (
            1.0 - self.alphas_cumprod)
        self.posterior_variance = torch.clamp(posterior_variance, min=1e-20)
        # sqrt(variance) evaluated per timestep as the reference does at step time (final_trainer.py:67-69)
        self.sqrt_posterior_variance = torch.stack([torch.sqrt(self.posterior_variance[t]) for t in range(num_timesteps)])

    _TABLES = ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
               "sqrt_recip_alphas", "posterior_variance", "sqrt_posterior_variance")

    def to(self, device):
        for name in self._TABLES:
            setattr(self, name, getattr(self, name).to(device).contiguous())
        return self

    def add_noise(self, x_0, noise, timesteps):
        self.to(x_0.device)
        x_0 = x_0.contiguous().float()
        out = torch.empty_like(x_0)
        b = x_0.shape[0]
        t = timesteps.to(device=x_0.device, dtype=torch.int64).contiguous()
        L.call("psg_q_sample", L.ptr(x_0), L.ptr(noise.contiguous().float()), L.ptr(t), L.ptr(self.sqrt_alphas_cumprod),
               L.ptr(self.sqrt_one_minus_alphas_cumprod), L.ptr(out), C.c_int(b), C.c_int(x_0.numel() // max(b, 1)),
               C.c_int(self.num_timesteps), C.c_int(0), C.c_float(0.0), C.c_float(0.0), L.ptr(None), L.stream_ptr())
        return out

    def sample_previous_timestep(self, x_t: torch.Tensor, predicted_noise: torch.Tensor, timestep: int,
                                 noise: torch.Tensor | None = None) -> torch.Tensor:
        """reference :52-71.  `noise` defaults to torch.randn_like(x_t) drawn here, as the reference does."""
        self.to(x_t.device)
        x_t = x_t.contiguous()
        if timestep > 0 and noise is None:
            noise = torch.randn_like(x_t)
        if timestep <= 0:
            noise = None
        out = torch.empty_like(x_t)
        L.call("psg_ddpm_step", L.ptr(x_t), L.ptr(predicted_noise.contiguous()), L.ptr(noise), L.ptr(out),
               C.c_longlong(x_t.numel()), C.c_int(1), L.ptr(self.sqrt_recip_alphas), L.ptr(self.betas),
               L.ptr(self.sqrt_one_minus_alphas_cumprod), L.ptr(self.sqrt_posterior_variance), C.c_int(int(timestep)),
               C.c_int(self.num_timesteps), L.stream_ptr())
        return out
