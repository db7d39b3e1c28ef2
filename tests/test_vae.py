"""Rows n1 / n4 of SURVEY.md §8f against goldens produced by EXECUTING the reference (oracle/make_golden_vae.py):
  * VAEDecoder (src/models/vae_decoder.py:128-222): module tree / seed-0 initialisation identical to the reference (CPU), and
    its output on CUDA in fp32 parity mode and in bf16 mode, for the default initialisation and an O(1)-gain state;
  * the two remaining reverse-step variants, bit-exact: `sample_prev_timestep` (src/training/diffusers_trainer.py:76-100) and
    the demo app's sampling loop (gradio_app.py:297-361), driven by the same bit-exact stub U-Net as tests/test_sampler.py.
"""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import record_metric

GOLD_VAE = Path(__file__).parent / "golden" / "vae_decoder.pt"
GOLD_REV = Path(__file__).parent / "golden" / "reverse_steps.npz"


@pytest.fixture(scope="module")
def gold_vae():
    return torch.load(GOLD_VAE, weights_only=False)


@pytest.fixture(scope="module")
def gold_rev():
    return dict(np.load(GOLD_REV))


def _vae_inputs(batch, text_len, seed):          # same draws as oracle/make_golden_vae.py:vae_inputs
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 8, 27, 27, generator=g), torch.randn(batch, text_len, 256, generator=g)


def _amplified(sd, seed):                        # same draws as oracle/make_golden_vae.py:amplified_vae_state
    g = torch.Generator().manual_seed(seed)
    amp = {}
    for k, v in sd.items():
        if v.dim() >= 2:
            amp[k] = torch.randn(v.shape, generator=g) * (1.6 / v[0].numel() ** 0.5)
        elif k.endswith("weight"):
            amp[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        else:
            amp[k] = 0.1 * torch.randn(v.shape, generator=g)
    return amp


# ---- CPU: module contract and schedule tables ---------------------------------------------------------------------------
def test_vae_decoder_init_matches_reference(gold_vae):
    from pokemon_sprite_generator_b200.vae import VAEDecoder
    torch.manual_seed(0)
    dec = VAEDecoder(latent_dim=8, text_dim=256, output_channels=3)
    sd = dec.state_dict()
    assert list(sd.keys()) == list(gold_vae["shapes"].keys())
    assert sum(p.numel() for p in dec.parameters()) == gold_vae["num_params"]
    for k, v in sd.items():
        assert tuple(v.shape) == gold_vae["shapes"][k], k
        s, a = gold_vae["checksums"][k]
        assert float(v.double().sum()) == s and float(v.double().abs().sum()) == a, k


def test_vae_decoder_refuses_cpu():
    from pokemon_sprite_generator_b200._lib import PsgError
    from pokemon_sprite_generator_b200.vae import VAEDecoder, ResNetBlock
    with pytest.raises(PsgError):
        VAEDecoder()(torch.zeros(1, 8, 27, 27), torch.zeros(1, 4, 256))
    with pytest.raises(PsgError):
        ResNetBlock(64, 64)(torch.zeros(1))


def test_alternate_schedule_tables_bit_identical(gold_rev):
    from pokemon_sprite_generator_b200.scheduler import DiffusersNoiseScheduler, GradioLinearSchedule
    ds = DiffusersNoiseScheduler()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
        assert np.array_equal(getattr(ds, k).numpy(), gold_rev["dt_" + k]), k
    gs = GradioLinearSchedule()
    for k in ("betas", "alphas", "alphas_cumprod"):
        assert np.array_equal(getattr(gs, k).numpy(), gold_rev["gr_" + k]), k


# ---- GPU: decoder parity --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("state", ["init", "amp"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_vae_decoder_matches_reference(cuda_device, gold_vae, state, mode):
    """fp32 mode: <= 2e-4 max-abs on the [-1, 1] image.  bf16 mode: the image passes through ~30 bf16 layers whose residual stream
    is re-rounded to 8 mantissa bits 15 times; measured on B200 (profiles/r02_parity_metrics.jsonl) mean-abs 4.5e-3 (0.6 LSB of
    an 8-bit image), max-abs 4.9e-2: the bounds are 2x those figures."""
    from pokemon_sprite_generator_b200.vae import VAEDecoder
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    torch.manual_seed(0)
    dec = VAEDecoder(compute_dtype=dt)
    if state == "amp":
        dec.load_state_dict(_amplified(dec.state_dict(), gold_vae["amp_seed"]))
    dec = dec.to(cuda_device).eval()
    tol_max, tol_mean = (2e-4, 2e-5) if mode == "fp32" else (1e-1, 9e-3)
    c1 = gold_vae["cases"][f"{state}_b1_l32"]
    lat, txt = _vae_inputs(1, 32, c1["seed"])
    y = dec(lat.to(cuda_device), txt.to(cuda_device)).cpu()
    assert y.shape == (1, 3, 215, 215) and torch.isfinite(y).all()
    d = (y - c1["output"]).abs()
    c2 = gold_vae["cases"][f"{state}_b2_l7"]
    lat, txt = _vae_inputs(2, 7, c2["seed"])
    y2 = dec(lat.to(cuda_device), txt.to(cuda_device)).cpu()
    d2 = (y2.flatten()[::c2["stride"]] - c2["output_strided"]).abs()
    print(f"[vae {state} {mode}] b1: max {d.max():.3e} mean {d.mean():.3e}; b2 strided: max {d2.max():.3e} mean {d2.mean():.3e}")
    record_metric(f"vae_decoder_{state}_{mode}", b1_max=float(d.max()), b1_mean=float(d.mean()), b2_max=float(d2.max()),
                  b2_mean=float(d2.mean()))
    assert d.max() <= tol_max and d.mean() <= tol_mean
    assert d2.max() <= tol_max and d2.mean() <= tol_mean
    assert abs(float(y2.std()) - c2["std"]) <= (1e-4 if mode == "fp32" else 1e-2)


# ---- GPU: reverse-step variants, bit-exact ----------------------------------------------------------------------------------
class _Stub(nn.Module):
    def forward(self, x, t, text):
        return x * 0.5 - 0.125


@pytest.mark.gpu
def test_diffusers_trainer_reverse_step_bit_exact(cuda_device, gold_rev):
    from pokemon_sprite_generator_b200.scheduler import DiffusersNoiseScheduler
    ds = DiffusersNoiseScheduler()
    x = torch.from_numpy(gold_rev["x"]).to(cuda_device)
    e = torch.from_numpy(gold_rev["eps"]).to(cuda_device)
    for t in (0, 1, 500, 999):
        torch.manual_seed(100 + t)
        z = torch.randn(2, 8, 27, 27).to(cuda_device) if t > 0 else None      # the reference draws randn_like(x_t) only for t > 0
        out = ds.sample_prev_timestep(x, e, t, noise=z)
        assert torch.equal(out.cpu(), torch.from_numpy(gold_rev[f"dt_step_t{t}"])), t
    torch.manual_seed(31)
    lat = torch.randn(2, 8, 27, 27).to(cuda_device)
    stub = _Stub()
    for t in list(range(999, -1, -34)) + [0]:
        z = torch.randn(2, 8, 27, 27).to(cuda_device) if t > 0 else None
        lat = ds.sample_prev_timestep(lat, stub(lat, None, None), t, noise=z)
    assert torch.equal(lat.cpu(), torch.from_numpy(gold_rev["dt_loop"]))


@pytest.mark.gpu
@pytest.mark.parametrize("steps", [20, 50])
def test_gradio_sampling_loop_bit_exact(cuda_device, gold_rev, steps):
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.scheduler import GradioLinearSchedule
    torch.manual_seed(41 + steps)
    out = sampler.gradio_sample(_Stub(), GradioLinearSchedule(), torch.zeros(2, 4, 256, device=cuda_device), num_inference_steps=steps,
                                noise_fn=lambda shape: torch.randn(shape))
    assert torch.equal(out.cpu(), torch.from_numpy(gold_rev[f"gr_loop_{steps}"]))


@pytest.mark.gpu
def test_text_to_sprite_runs_end_to_end(cuda_device):
    """Config 5 downstream of the text encoder: 50 posterior steps (CUDA-graph U-Net) + VAE decoder; shape / range / determinism."""
    from pokemon_sprite_generator_b200 import sampler
    from pokemon_sprite_generator_b200.unet import UNet
    from pokemon_sprite_generator_b200.vae import VAEDecoder
    torch.manual_seed(0)
    unet = UNet(num_heads=4).to(cuda_device).eval()
    dec = VAEDecoder().to(cuda_device).eval()
    text = torch.randn(2, 32, 256, device=cuda_device)
    torch.manual_seed(5)
    a = sampler.text_to_sprite(unet, dec, text, num_inference_steps=4, use_cuda_graph=True)
    torch.manual_seed(5)
    b = sampler.text_to_sprite(unet, dec, text, num_inference_steps=4, use_cuda_graph=False)
    assert a.shape == (2, 3, 215, 215) and torch.isfinite(a).all() and a.min() >= 0 and a.max() <= 1
    assert torch.equal(a, b)
