"""ORACLE / test infrastructure: golden vectors for the rows SURVEY.md §8 marks "next" -- the VAE decoder (n1) and the two
remaining reverse-step variants (n4) -- produced by EXECUTING the unmodified reference from /root/reference (build container only).

    python -m oracle.make_golden_vae           # ~1 min on 8 CPU threads

Outputs:
  tests/golden/vae_decoder.pt     seed-0 initialisation checksums of the reference VAEDecoder, and its output for seeded latents /
                                  text embeddings: the full image for one sample + strided samples and statistics for a batch of 2
                                  at another text length (src/models/vae_decoder.py:128-222)
  tests/golden/vae_encoder.pt     the same for the reference VAEEncoder (src/models/vae_decoder.py:68-125): (latent, mu, logvar) with the
                                  reparameterisation noise pinned by a seed
  tests/golden/reverse_steps.npz  src/training/diffusers_trainer.py:76-100 (`sample_prev_timestep`) single steps and a 30-step loop,
                                  gradio_app.py:297-361 (`ddpm_sample`) 20- and 50-step loops with a stub U-Net, and the schedule
                                  tables both use
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

from . import ref_loader

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"


def _stub_unet(x, t, text):
    return x * 0.5 - 0.125      # bit-exact on CPU and CUDA: power-of-two scale, exactly representable offset


def _load_extra():
    ref_loader.install_stubs()
    root = ref_loader.REFERENCE_ROOT
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    if "gradio" not in sys.modules:
        m = types.ModuleType("gradio")
        m.__dict__.update(Blocks=_Any(), themes=_Any())
        m.__path__ = []
        sys.modules["gradio"] = m
    from src.models.vae_decoder import VAEDecoder, VAEEncoder  # type: ignore
    from src.training import diffusers_trainer  # type: ignore
    import gradio_app  # type: ignore
    return VAEDecoder, diffusers_trainer.NoiseScheduler, gradio_app.PokemonGradioGenerator, VAEEncoder


def vae_inputs(batch: int, text_len: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    latent = torch.randn(batch, 8, 27, 27, generator=g)
    text = torch.randn(batch, text_len, 256, generator=g)
    return latent, text


def vae_golden(VAEDecoder):
    torch.manual_seed(0)
    dec = VAEDecoder(latent_dim=8, text_dim=256, output_channels=3).eval()
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    golden = {"checksums": {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()},
              "shapes": {k: tuple(v.shape) for k, v in sd.items()}, "num_params": sum(p.numel() for p in dec.parameters()), "cases": {}}
    # the default initialisation gives tiny activations; a second state with O(1) gains exercises every layer's numerics
    amp = amplified_vae_state(sd, 11)
    golden["amp_seed"] = 11
    for name, state in (("init", sd), ("amp", amp)):
        dec.load_state_dict(state)
        with torch.no_grad():
            lat, txt = vae_inputs(1, 32, 777)
            y = dec(lat, txt)
            golden["cases"][f"{name}_b1_l32"] = {"batch": 1, "text_len": 32, "seed": 777, "output": y.clone()}
            lat, txt = vae_inputs(2, 7, 778)
            y2 = dec(lat, txt)
            golden["cases"][f"{name}_b2_l7"] = {"batch": 2, "text_len": 7, "seed": 778, "stride": 13,
                                                "output_strided": y2.flatten()[::13].clone(), "mean": float(y2.mean()), "std": float(y2.std()),
                                                "absmax": float(y2.abs().max())}
        print(f"vae {name}: b1 |y|max={y.abs().max():.4f} std={y.std():.4f}; b2 std={y2.std():.4f}", flush=True)
    torch.save(golden, OUT / "vae_decoder.pt")
    print("wrote vae_decoder.pt")


def encoder_inputs(batch: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, 215, 215, generator=g) * 0.5


def encoder_golden(VAEEncoder):
    """src/models/vae_decoder.py:68-125: (latent, mu, logvar) of the seed-0 reference encoder and of an O(1)-gain state; the
    reparameterisation noise is the first draw after torch.manual_seed(55) (the encoder draws nothing else in eval mode)."""
    torch.manual_seed(0)
    enc = VAEEncoder(input_channels=3, latent_dim=8).eval()
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    golden = {"checksums": {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()},
              "shapes": {k: tuple(v.shape) for k, v in sd.items()}, "num_params": sum(p.numel() for p in enc.parameters()),
              "amp_seed": 12, "noise_seed": 55, "cases": {}}
    for name, state in (("init", sd), ("amp", amplified_vae_state(sd, 12))):
        enc.load_state_dict(state)
        img = encoder_inputs(2, 888)
        torch.manual_seed(55)
        with torch.no_grad():
            lat, mu, lv = enc(img)
        golden["cases"][name] = {"batch": 2, "seed": 888, "latent": lat.clone(), "mu": mu.clone(), "logvar": lv.clone()}
        print(f"encoder {name}: mu std {mu.std():.4f} logvar std {lv.std():.4f} latent std {lat.std():.4f}", flush=True)
    torch.save(golden, OUT / "vae_encoder.pt")
    print("wrote vae_encoder.pt")


def amplified_vae_state(sd, seed: int = 11):
    """The O(1)-gain state of vae_golden, regenerated from the seed (tests rebuild it instead of shipping 40 MB of weights)."""
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


def reverse_golden(DiffSched, Gradio):
    out = {}
    ds = DiffSched()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
        out["dt_" + k] = getattr(ds, k).numpy()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 8, 27, 27, generator=g)
    e = torch.randn(2, 8, 27, 27, generator=g)
    out["x"], out["eps"] = x.numpy(), e.numpy()
    for t in (0, 1, 500, 999):
        torch.manual_seed(100 + t)
        out[f"dt_step_t{t}"] = ds.sample_prev_timestep(x, e, t).numpy()
    torch.manual_seed(31)
    lat = torch.randn(2, 8, 27, 27)
    for t in range(999, -1, -34):       # 30 steps, ends at t = 13; one more at t = 0
        lat = ds.sample_prev_timestep(lat, _stub_unet(lat, None, None), t)
    lat = ds.sample_prev_timestep(lat, _stub_unet(lat, None, None), 0)
    out["dt_loop"] = lat.numpy()
    # gradio loop: the app's linear schedule (gradio_app.py:281-284) and its ddpm_sample (:297-361), driven unbound
    for steps in (20, 50):
        fake = types.SimpleNamespace(config={"model": {"latent_dim": 8}}, device=torch.device("cpu"), num_timesteps=1000,
                                     use_diffusers=False, unet=_stub_unet)
        fake.betas = torch.linspace(0.0001, 0.02, 1000)
        fake.alphas = 1.0 - fake.betas
        fake.alphas_cumprod = torch.cumprod(fake.alphas, dim=0)
        torch.manual_seed(41 + steps)
        out[f"gr_loop_{steps}"] = Gradio.ddpm_sample(fake, torch.zeros(2, 4, 256), num_inference_steps=steps).numpy()
    out["gr_betas"], out["gr_alphas"], out["gr_alphas_cumprod"] = fake.betas.numpy(), fake.alphas.numpy(), fake.alphas_cumprod.numpy()
    np.savez_compressed(OUT / "reverse_steps.npz", **out)
    print("wrote reverse_steps.npz", {k: v.shape for k, v in out.items() if "loop" in k or "step" in k})


def main():
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    VAEDecoder, DiffSched, Gradio, VAEEncoder = _load_extra()
    if "--encoder-only" not in sys.argv:
        reverse_golden(DiffSched, Gradio)
        vae_golden(VAEDecoder)
    encoder_golden(VAEEncoder)


if __name__ == "__main__":
    main()
