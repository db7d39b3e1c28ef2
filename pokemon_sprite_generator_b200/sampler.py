"""DDPM samplers over the CUDA U-Net: the reference's `ddpm_sample` (src/training/improved_diffusion_trainer.py:508-569)
and the 50-step posterior-variance loop of `FinalPokemonGenerator.forward` (src/training/final_trainer.py:183-204).

RNG contract: one `randn` for x_T, then one `randn` per step *after* the U-Net call, all on the
sampling device's default generator -- the same draw order as the reference.  `noise_fn(shape)` overrides the source
(tests inject pre-drawn CPU noise to compare bit-exactly with goldens made on CPU).  Prompt batches shard across
GPUs with no communication: each rank simply samples its own slice.

`use_cuda_graph=True` captures the U-Net forward (about 600 kernel launches) once and replays it per step; the
timestep and latent live in static device buffers.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .scheduler import LinearNoiseScheduler, NoiseScheduler


class _GraphedUNet:
    """Static-shape CUDA-graph replay of unet(x, t, text) in eval / no-grad mode."""

    def __init__(self, unet, x: torch.Tensor, t: torch.Tensor, text: torch.Tensor):
        self.x, self.t, self.text = x.clone(), t.clone(), text.clone()
        eng = unet.engine()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):  # warm-up: flattens/packs weights and sizes every workspace outside the capture
                unet(self.x, self.t, self.text)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = unet(self.x, self.t, self.text)
        self._version = eng.store.version()

    def __call__(self, x: torch.Tensor, t_value: int) -> torch.Tensor:
        self.x.copy_(x)
        self.t.fill_(t_value)
        self.graph.replay()
        return self.out


@torch.no_grad()
def ddpm_sample(unet, scheduler: NoiseScheduler, text_emb: torch.Tensor, num_samples: int, fast_sampling: bool = True,
                latent_dim: int = 8, noise_fn: Optional[Callable[[tuple], torch.Tensor]] = None,
                use_cuda_graph: bool = False) -> torch.Tensor:
    dev = text_emb.device
    shape = (num_samples, latent_dim, 27, 27)
    draw = noise_fn if noise_fn is not None else (lambda s: torch.randn(s, device=dev))
    was_training = unet.training
    unet.eval()
    x = draw(shape).to(dev)
    scheduler.to(dev)
    steps = list(range(0, scheduler.num_timesteps, 50)) if fast_sampling else list(range(scheduler.num_timesteps))
    graphed = None
    if use_cuda_graph:
        graphed = _GraphedUNet(unet, x, torch.zeros(num_samples, dtype=torch.long, device=dev), text_emb.float().contiguous())
    for t in reversed(steps):
        if graphed is not None:
            eps = graphed(x, t)
        else:
            eps = unet(x, torch.full((num_samples,), t, device=dev, dtype=torch.long), text_emb)
        z = draw(shape).to(dev) if t > 0 else None     # reference :560-567 (always true for t > 0 in both modes)
        x = scheduler.ddpm_step(x, eps, t, z)
    unet.train(was_training)
    return x


@torch.no_grad()
def posterior_sample(unet, scheduler: LinearNoiseScheduler, text_emb: torch.Tensor, num_inference_steps: int = 50,
                     latent_dim: int = 8, noise_fn: Optional[Callable[[tuple], torch.Tensor]] = None,
                     use_cuda_graph: bool = False) -> torch.Tensor:
    """Latents of FinalPokemonGenerator.forward(mode='generate') before the VAE decoder."""
    dev = text_emb.device
    B = text_emb.shape[0]
    shape = (B, latent_dim, 27, 27)
    draw = noise_fn if noise_fn is not None else (lambda s: torch.randn(s, device=dev))
    was_training = unet.training
    unet.eval()
    lat = draw(shape).to(dev)
    step = max(1, scheduler.num_timesteps // num_inference_steps)
    graphed = None
    if use_cuda_graph:
        graphed = _GraphedUNet(unet, lat, torch.zeros(B, dtype=torch.long, device=dev), text_emb.float().contiguous())
    for i in range(num_inference_steps):
        ts = max(0, scheduler.num_timesteps - 1 - i * step)
        eps = graphed(lat, ts) if graphed is not None else unet(lat, torch.full((B,), ts, device=dev, dtype=torch.long), text_emb)
        if ts > 0:
            lat = scheduler.sample_previous_timestep(lat, eps, ts, noise=draw(shape).to(dev))
        else:
            lat = lat - eps        # reference :202-204 (dead at 50 steps: t never reaches 0)
    unet.train(was_training)
    return lat
