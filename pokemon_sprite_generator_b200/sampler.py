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

from .scheduler import DiffusersNoiseScheduler, GradioLinearSchedule, LinearNoiseScheduler, NoiseScheduler


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


# One captured forward per (U-Net, batch, text length): capturing costs about a second (two warm-up forwards + the capture), which
# a 50-step text-to-sprite call would otherwise pay every time.  The graph reads the weights in place, so optimiser updates are
# seen by later replays; re-flattened / re-packed weights (another store version) or another shape re-capture.
_graph_cache = {}


def _graphed_for(unet, x: torch.Tensor, t: torch.Tensor, text: torch.Tensor) -> _GraphedUNet:
    eng = unet.engine()
    eng.prepare(x.device)
    key = (id(unet), tuple(x.shape), tuple(text.shape), str(x.device))
    g = _graph_cache.get(key)
    if g is None or g._version != eng.store.version():
        g = _GraphedUNet(unet, x, t, text)
        _graph_cache.clear()                      # one resident graph: its private memory pool is a full forward's activations
        _graph_cache[key] = g
    else:
        g.text.copy_(text)
    return g


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
        graphed = _graphed_for(unet, x, torch.zeros(num_samples, dtype=torch.long, device=dev), text_emb.float().contiguous())
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
        graphed = _graphed_for(unet, lat, torch.zeros(B, dtype=torch.long, device=dev), text_emb.float().contiguous())
    for i in range(num_inference_steps):
        ts = max(0, scheduler.num_timesteps - 1 - i * step)
        eps = graphed(lat, ts) if graphed is not None else unet(lat, torch.full((B,), ts, device=dev, dtype=torch.long), text_emb)
        if ts > 0:
            lat = scheduler.sample_previous_timestep(lat, eps, ts, noise=draw(shape).to(dev))
        else:
            lat = lat - eps        # reference :202-204 (dead at 50 steps: t never reaches 0)
    unet.train(was_training)
    return lat


@torch.no_grad()
def gradio_sample(unet, schedule: GradioLinearSchedule, text_emb: torch.Tensor, num_inference_steps: int = 50, latent_dim: int = 8,
                  initial_latent: Optional[torch.Tensor] = None, noise_fn: Optional[Callable[[tuple], torch.Tensor]] = None) -> torch.Tensor:
    """The demo app's sampling loop (gradio_app.py:297-361): timesteps linspace(T-1, 0, steps) truncated to integers; per step
    denoise with (1-alpha_t)/sqrt(1-abar_t), then re-noise to the next timestep while it is > 0."""
    dev = text_emb.device
    B = text_emb.shape[0]
    shape = (B, latent_dim, 27, 27)
    draw = noise_fn if noise_fn is not None else (lambda s: torch.randn(s, device=dev))
    was_training = getattr(unet, "training", False)
    if hasattr(unet, "eval"):
        unet.eval()
    lat = draw(shape).to(dev) if initial_latent is None else initial_latent.clone()
    ts = torch.linspace(schedule.num_timesteps - 1, 0, num_inference_steps, dtype=torch.long).tolist()
    for i, t in enumerate(ts):
        eps = unet(lat, torch.full((B,), t, device=dev, dtype=torch.long), text_emb)
        if i < len(ts) - 1:
            nxt = ts[i + 1]
            lat = schedule.step(lat, eps, t, nxt, noise=draw(shape).to(dev) if nxt > 0 else None)
        else:
            lat = schedule.step(lat, eps, t, None)
    if hasattr(unet, "train"):
        unet.train(was_training)
    return lat


@torch.no_grad()
def text_to_sprite(unet, vae_decoder, text_emb: torch.Tensor, scheduler: Optional[LinearNoiseScheduler] = None,
                   num_inference_steps: int = 50, use_cuda_graph: bool = True) -> torch.Tensor:
    """BASELINE config 5 downstream of the text encoder: 50 posterior DDPM steps of the U-Net (FinalPokemonGenerator.forward,
    src/training/final_trainer.py:183-204), then the VAE decoder to [B, 3, 215, 215], mapped to [0, 1] as `generate_samples` does
    (src/training/improved_diffusion_trainer.py:598)."""
    scheduler = scheduler or LinearNoiseScheduler()
    lat = posterior_sample(unet, scheduler, text_emb, num_inference_steps, use_cuda_graph=use_cuda_graph)
    img = vae_decoder(lat, text_emb)
    return torch.clamp((img + 1.0) / 2.0, 0, 1)
