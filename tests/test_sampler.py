"""The sampler FUNCTIONS themselves (sampler.ddpm_sample / sampler.posterior_sample / the CUDA-graph replay the `denoise`
bench number is timed on) against the goldens produced by executing the reference's loops (oracle/make_golden.py):
`ImprovedDiffusionTrainer.ddpm_sample` (src/training/improved_diffusion_trainer.py:508-569, driven unbound with a stub
U-Net) and the 50-step loop of `FinalPokemonGenerator.forward` (src/training/final_trainer.py:183-204).

The stub U-Net (x*0.5 - 0.125: power-of-two scale, exactly representable offset) is bit-exact on CPU and CUDA, so draw
order, step lists, the `t > 0` rule and every coefficient are pinned with torch.equal.
"""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "scheduler_tables.npz"


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


class StubUNet(nn.Module):
    """Same arithmetic as oracle/make_golden.py:_stub_unet; records the timesteps it was called with."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, x, t, text):
        assert t.dtype == torch.long and t.shape == (x.shape[0],) and bool((t == t[0]).all())
        self.calls.append(int(t[0]))
        return x * 0.5 - 0.125


def _cpu_noise(seed):
    """noise_fn drawing from the CPU default generator in the reference's order (goldens were made on CPU)."""
    torch.manual_seed(seed)
    return lambda shape: torch.randn(shape)


@pytest.mark.parametrize("fast", [True, False])
def test_ddpm_sample_function_bit_exact(cuda_device, gold, fast):
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    stub = StubUNet().to(cuda_device)
    stub.train()
    text = torch.zeros(2, 4, 256, device=cuda_device)
    x = sampler.ddpm_sample(stub, NoiseScheduler(), text, 2, fast_sampling=fast, noise_fn=_cpu_noise(99))
    assert torch.equal(x.cpu(), torch.from_numpy(gold["ddpm_fast" if fast else "ddpm_full"]))
    want = list(reversed(range(0, 1000, 50))) if fast else list(reversed(range(1000)))
    assert stub.calls == want
    assert stub.training, "ddpm_sample must restore train mode"


def test_posterior_sample_function_bit_exact(cuda_device, gold):
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler
    stub = StubUNet().to(cuda_device).eval()
    text = torch.zeros(2, 4, 256, device=cuda_device)
    lat = sampler.posterior_sample(stub, LinearNoiseScheduler(), text, 50, noise_fn=_cpu_noise(123))
    assert torch.equal(lat.cpu(), torch.from_numpy(gold["posterior_50"]))
    assert stub.calls == [999 - 20 * i for i in range(50)]
    # more steps than timesteps/1 reaches t == 0: x - eps branch (final_trainer.py:202-204)
    stub.calls.clear()
    lat2 = sampler.posterior_sample(stub, LinearNoiseScheduler(num_timesteps=4), text, 6, noise_fn=_cpu_noise(1))
    assert stub.calls == [3, 2, 1, 0, 0, 0] and bool(torch.isfinite(lat2).all())


@pytest.fixture(scope="module")
def real_unet(cuda_device):
    from pokemon_sprite_generator_b200.unet import UNet
    torch.manual_seed(0)
    return UNet(num_heads=4).to(cuda_device).eval()


def test_graphed_unet_replay_equals_eager(cuda_device, real_unet):
    """_GraphedUNet (the path bench.py times for `denoise`): replay == eager forward, bit for bit, at several timesteps and
    for changing latents, on the real U-Net (bf16 tensor-core path)."""
    from pokemon_sprite_generator_b200.sampler import _GraphedUNet
    g = torch.Generator(device="cuda").manual_seed(11)
    B = 3
    text = torch.randn(B, 32, 256, device=cuda_device, generator=g)
    x0 = torch.randn(B, 8, 27, 27, device=cuda_device, generator=g)
    graphed = _GraphedUNet(real_unet, x0, torch.zeros(B, dtype=torch.long, device=cuda_device), text)
    for t in (999, 500, 37, 0):
        x = torch.randn(B, 8, 27, 27, device=cuda_device, generator=g)
        with torch.no_grad():
            eager = real_unet(x, torch.full((B,), t, device=cuda_device, dtype=torch.long), text)
        rep = graphed(x, t).clone()
        assert torch.equal(rep, eager), f"t={t}: max diff {(rep - eager).abs().max().item():.3e}"


def test_ddpm_sample_graph_path_equals_eager_path(cuda_device, real_unet):
    """20-step fast sampling on the real U-Net: use_cuda_graph=True produces exactly the eager path's latents."""
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler, NoiseScheduler
    text = torch.randn(2, 32, 256, device=cuda_device, generator=torch.Generator(device="cuda").manual_seed(5))
    a = sampler.ddpm_sample(real_unet, NoiseScheduler(), text, 2, fast_sampling=True, noise_fn=_cpu_noise(7))
    b = sampler.ddpm_sample(real_unet, NoiseScheduler(), text, 2, fast_sampling=True, noise_fn=_cpu_noise(7), use_cuda_graph=True)
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    c = sampler.posterior_sample(real_unet, LinearNoiseScheduler(), text, 10, noise_fn=_cpu_noise(8))
    d = sampler.posterior_sample(real_unet, LinearNoiseScheduler(), text, 10, noise_fn=_cpu_noise(8), use_cuda_graph=True)
    assert torch.equal(c, d) and bool(torch.isfinite(c).all())


def test_ddpm_sample_default_rng_order(cuda_device):
    """Without noise_fn the draws come from the device generator: x_T first, then one draw per step after the U-Net call
     -- re-seeding reproduces the run, and the manual loop with the same draws gives the same latents."""
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    stub = StubUNet().to(cuda_device).eval()
    text = torch.zeros(2, 4, 256, device=cuda_device)
    ns = NoiseScheduler()
    torch.manual_seed(42)
    a = sampler.ddpm_sample(stub, ns, text, 2, fast_sampling=True)
    torch.manual_seed(42)
    x = torch.randn(2, 8, 27, 27, device=cuda_device)
    for t in reversed(range(0, 1000, 50)):
        eps = x * 0.5 - 0.125
        z = torch.randn(2, 8, 27, 27, device=cuda_device) if t > 0 else None
        x = ns.ddpm_step(x, eps, t, z)
    assert torch.equal(a, x)
