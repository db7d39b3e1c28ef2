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


def _reverse_step(x_t: torch.Tensor, eps: torch.Tensor, noise, mode: int, coef) -> torch.Tensor:
    x_t = x_t.float().contiguous()
    eps = eps.float().contiguous()
    if noise is not None:
        noise = noise.float().contiguous()
    out = torch.empty_like(x_t)
    arr = (C.c_float * 5)(*[float(c) for c in coef] + [0.0] * (5 - len(coef)))
    L.call("psg_reverse_step", L.ptr(x_t), L.ptr(eps), L.ptr(noise), L.ptr(out), C.c_longlong(x_t.numel()), C.c_int(mode), arr,
           L.stream_ptr())
    return out


class DiffusersNoiseScheduler:
    """The scheduler of the reference's diffusers-U-Net trainer (src/training/diffusers_trainer.py:27-100): the clipped-cosine
    schedule plus a posterior variance whose first entry is copied from the second, and the x0-prediction reverse step
    `sample_prev_timestep`.  Tables are built with the reference's torch ops in the reference's order (bit-identical on CPU);
    the step's scalars are evaluated as 0-dim fp32 tensor ops exactly as the reference evaluates them at step time."""

    def __init__(self, num_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02):
        self.num_timesteps = num_timesteps
        self.betas = NoiseScheduler._cosine_beta_schedule(num_timesteps, beta_start, beta_end).float()
        self.alphas = (1.0 - self.betas).float()
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0).float()
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod).float()
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod).float()
        self.posterior_variance = self.betas * (1.0 - torch.cat([torch.tensor([1.0]), self.alphas_cumprod[:-1]])) / (
            1.0 - self.alphas_cumprod)
        self.posterior_variance[0] = self.posterior_variance[1]
        self.sqrt_alphas_cumprod = torch.clamp(self.sqrt_alphas_cumprod, min=1e-8)
        self.sqrt_one_minus_alphas_cumprod = torch.clamp(self.sqrt_one_minus_alphas_cumprod, min=1e-8)
        self._coef = {}

    def to(self, device):
        return self          # the tables stay on the host: only five scalars per step reach the kernel

    def add_noise(self, x_0: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """reference :59-74 (the NaN/Inf clamp of the result is the caller's finite check here)."""
        return self._as_cosine().add_noise(x_0, noise, timesteps)      # same tables: the fused q_sample kernel

    def _as_cosine(self):
        if not hasattr(self, "_cos"):
            self._cos = NoiseScheduler(self.num_timesteps)
        return self._cos

    def step_coefficients(self, timestep: int):
        """(sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev), sqrt(posterior_variance_t)), reference :84-98."""
        c = self._coef.get(timestep)
        if c is None:
            abar_t = self.alphas_cumprod[timestep]
            abar_prev = self.alphas_cumprod[timestep - 1] if timestep > 0 else torch.tensor(1.0)
            c = (float(torch.sqrt(1 - abar_t)), float(torch.sqrt(abar_t)), float(torch.sqrt(abar_prev)),
                 float(torch.sqrt(1 - abar_prev)), float(torch.sqrt(self.posterior_variance[timestep])))
            self._coef[timestep] = c
        return c

    def sample_prev_timestep(self, x_t: torch.Tensor, noise_pred: torch.Tensor, timestep: int,
                             noise: torch.Tensor | None = None) -> torch.Tensor:
        """reference :76-100.  `noise` defaults to torch.randn_like(x_t) drawn here when timestep > 0, as the reference does."""
        timestep = int(timestep)
        if timestep > 0 and noise is None:
            noise = torch.randn_like(x_t)
        if timestep <= 0:
            noise = None
        return _reverse_step(x_t, noise_pred, noise, 2, self.step_coefficients(timestep))


class GradioLinearSchedule:
    """The linear schedule and the denoise / re-noise step of the reference's demo app (gradio_app.py:281-284, 342-359)."""

    def __init__(self, num_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02):
        self.num_timesteps = num_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)

    def step(self, latent: torch.Tensor, predicted_noise: torch.Tensor, t: int, next_t: int | None,
             noise: torch.Tensor | None = None) -> torch.Tensor:
        """One iteration of the loop: denoise at t; if there is a next step with next_t > 0, re-noise to it (drawing
        torch.randn_like(latent) unless `noise` is given).  next_t = None: the last step."""
        a_t = self.alphas[int(t)]
        c0 = float((1 - a_t) / torch.sqrt(1 - self.alphas_cumprod[int(t)]))
        c1 = float(torch.sqrt(a_t))
        if next_t is None or int(next_t) <= 0:
            return _reverse_step(latent, predicted_noise, None, 3, (c0, c1))
        a_n = self.alphas[int(next_t)]
        if noise is None:
            noise = torch.randn_like(latent)
        return _reverse_step(latent, predicted_noise, noise, 3, (c0, c1, float(torch.sqrt(a_n)), float(torch.sqrt(1 - a_n))))
