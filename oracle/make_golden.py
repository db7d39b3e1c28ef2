"""Generate tests/golden/* by EXECUTING the unmodified reference from /root/reference (build container only).

    python -m oracle.make_golden            # ~2-3 min on 8 CPU threads

Outputs (all small):
  tests/golden/scheduler_tables.npz   cosine + linear schedule tables, q_sample / reverse-step known answers
  tests/golden/unet_cases.pt          U-Net outputs, losses, per-parameter grad norms for seeded inputs and the
                                      seed-0 reference initialisation (+ its per-key checksums)
"""
from __future__ import annotations

import types
from pathlib import Path

import numpy as np
import torch

from . import inputs, ref_loader

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"


def _stub_unet(x, t, text):
    # bit-exact on CPU and CUDA: power-of-two scale, exactly representable offset
    return x * 0.5 - 0.125


def scheduler_golden(ref):
    out = {}
    ns = ref.NoiseScheduler()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        out["cos_" + k] = getattr(ns, k).numpy()
    ls = ref.FinalNoiseScheduler()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas",
              "posterior_variance"):
        out["lin_" + k] = getattr(ls, k).numpy()
    # q_sample known answers
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(6, 8, 27, 27, generator=g) * 2.0
    eps = torch.randn(6, 8, 27, 27, generator=g)
    t = torch.tensor([0, 12, 500, 902, 999, 1])
    out["qs_x0"], out["qs_eps"], out["qs_t"] = x0.numpy(), eps.numpy(), t.numpy()
    out["qs_cos"] = ns.add_noise(x0, eps, t).numpy()
    out["qs_cos_clamped"] = ns.add_noise(torch.clamp(x0, -3.0, 3.0), eps, t).numpy()
    out["qs_lin"] = ls.add_noise(x0, eps, t).numpy()
    # ddpm_sample driven unbound with a stub U-Net
    for fast in (True, False):
        torch.manual_seed(99)
        fake = types.SimpleNamespace(config={"model": {"latent_dim": 8}}, device=torch.device("cpu"), unet=_stub_unet,
                                     noise_scheduler=ref.NoiseScheduler())
        text = torch.zeros(2, 4, 256)
        out["ddpm_fast" if fast else "ddpm_full"] = ref.ImprovedDiffusionTrainer.ddpm_sample(fake, text, 2, fast_sampling=fast).numpy()
    # FinalPokemonGenerator's 50-step loop (final_trainer.py:186-204) around sample_previous_timestep
    torch.manual_seed(123)
    lat = torch.randn(2, 8, 27, 27)
    step = max(1, ls.num_timesteps // 50)
    for i in range(50):
        ts = max(0, ls.num_timesteps - 1 - i * step)
        pred = _stub_unet(lat, None, None)
        lat = ls.sample_previous_timestep(lat, pred, ts) if ts > 0 else lat - pred
    out["posterior_50"] = lat.numpy()
    # single posterior steps at t=0 and t=999 with explicit noise
    torch.manual_seed(5)
    x = torch.randn(2, 8, 27, 27)
    e = torch.randn(2, 8, 27, 27)
    out["post_x"], out["post_eps"] = x.numpy(), e.numpy()
    out["post_t0"] = ls.sample_previous_timestep(x, e, 0).numpy()
    torch.manual_seed(6)
    out["post_t999"] = ls.sample_previous_timestep(x, e, 999).numpy()
    np.savez_compressed(OUT / "scheduler_tables.npz", **out)
    print("wrote scheduler_tables.npz", {k: v.shape for k, v in out.items() if k.startswith(("ddpm", "post"))})


def _run_case(ref, unet, heads, batch, text_len, with_grad, seed=1234, fp64=False):
    latent, text, t, noise = inputs.make_inputs(batch, text_len, seed)
    ns = ref.NoiseScheduler()
    noisy = ns.add_noise(latent, noise, t)
    for blk in unet.modules():
        if hasattr(blk, "num_heads") and hasattr(blk, "in_proj_weight"):   # nn.MultiheadAttention
            blk.num_heads = heads
            blk.head_dim = blk.embed_dim // heads
    unet.zero_grad(set_to_none=True)
    case = {"heads": heads, "batch": batch, "text_len": text_len, "seed": seed}
    if with_grad:
        pred = unet(noisy, t, text)
        loss = torch.nn.SmoothL1Loss(beta=0.1)(pred, noise)
        loss.backward()
        grads = {k: p.grad for k, p in unet.named_parameters()}
        case["loss"] = loss.item()
        case["grad_norms"] = {k: grads[k].norm().item() for k in grads}
        case["grad_total_norm"] = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())))
        case["grad_samples"] = {k: grads[k].flatten()[:: max(1, grads[k].numel() // 64)][:64].clone() for k in inputs.GRAD_KEYS}
    else:
        with torch.no_grad():
            pred = unet(noisy, t, text)
    case["output"] = pred.detach().clone()
    if with_grad and fp64:
        # The reference's fp32 CPU conv weight-gradients lose precision on the deep levels (up to 3e-3 on a
        # parameter's grad norm vs exact); an fp64 run of the SAME reference module pins the exact values.
        unet.double()
        unet.zero_grad(set_to_none=True)
        pred64 = unet(noisy.double(), t, text.double())
        loss64 = torch.nn.SmoothL1Loss(beta=0.1)(pred64, noise.double())
        loss64.backward()
        g64 = {k: p.grad for k, p in unet.named_parameters()}
        case["loss_fp64"] = loss64.item()
        case["output_fp64"] = pred64.detach().float().clone()
        case["grad_norms_fp64"] = {k: g64[k].norm().item() for k in g64}
        case["grad_total_norm_fp64"] = float(torch.sqrt(sum(g.pow(2).sum() for g in g64.values())))
        case["grad_samples_fp64"] = {k: g64[k].flatten()[:: max(1, g64[k].numel() // 64)][:64].float().clone() for k in inputs.GRAD_KEYS}
        unet.float()
        unet.zero_grad(set_to_none=True)
    print(f"case heads={heads} B={batch} L={text_len} grad={with_grad}: |y|max={pred.abs().max():.4f} std={pred.std():.4f}"
          + (f" loss={case['loss']:.6f} gnorm={case['grad_total_norm']:.4f}" if with_grad else ""))
    return case


def unet_golden(ref):
    torch.manual_seed(0)
    unet = ref.UNet(latent_dim=8, text_dim=256, time_emb_dim=128, num_heads=8).eval()   # eval: dropout off
    sd = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    golden = {
        "checksums": {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()},
        "shapes": {k: tuple(v.shape) for k, v in sd.items()},
        "num_params": sum(p.numel() for p in unet.parameters()),
        "cases": {},
    }
    golden["cases"]["init_h8_b2_l32"] = _run_case(ref, unet, 8, 2, 32, True, fp64=True)
    golden["cases"]["init_h4_b2_l32"] = _run_case(ref, unet, 4, 2, 32, False)
    golden["cases"]["init_h8_b1_l7"] = _run_case(ref, unet, 8, 1, 7, False)
    golden["cases"]["init_h8_b3_l77"] = _run_case(ref, unet, 8, 3, 77, False)
    unet.load_state_dict(inputs.amplify_state_dict(sd))
    golden["cases"]["amp_h8_b2_l32"] = _run_case(ref, unet, 8, 2, 32, True, fp64=True)
    golden["cases"]["amp_h4_b2_l32"] = _run_case(ref, unet, 4, 2, 32, True)
    torch.save(golden, OUT / "unet_cases.pt")
    print("wrote unet_cases.pt")


def main():
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    ref = ref_loader.load()
    scheduler_golden(ref)
    unet_golden(ref)


if __name__ == "__main__" and "--b256" not in __import__("sys").argv:
    main()


def unet_b256_golden(ref, batch: int = 256, seed: int = 4321, stride: int = 23):
    """One reference golden AT THE BENCHMARK SHAPE (BASELINE config 2: batch 256, heads 4, 32 text tokens), seed-0 init and the
    O(1)-gain re-init: loss, per-parameter gradient norms, gradient samples and a strided sample of the noise prediction
    (every `stride`-th element of the flattened output + samples 0 and batch-1 in full).  ~10 CPU-minutes and ~40 GB per case
    on 8 threads; written to tests/golden/unet_b256.pt (small).        python -m oracle.make_golden --b256"""
    torch.manual_seed(0)
    unet = ref.UNet(latent_dim=8, text_dim=256, time_emb_dim=128, num_heads=4).eval()
    sd = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    golden = {"batch": batch, "seed": seed, "stride": stride, "heads": 4, "text_len": 32, "cases": {}}
    for name, state in (("init", sd), ("amp", inputs.amplify_state_dict(sd))):
        unet.load_state_dict(state)
        case = _run_case(ref, unet, 4, batch, 32, True, seed=seed)
        out = case.pop("output")
        case["output_strided"] = out.flatten()[::stride].clone()
        case["output_first"], case["output_last"] = out[0].clone(), out[-1].clone()
        case["output_absmax"], case["output_std"] = float(out.abs().max()), float(out.std())
        golden["cases"][name] = case
        torch.save(golden, OUT / "unet_b256.pt")
        print(f"wrote unet_b256.pt ({name})", flush=True)


if __name__ == "__main__" and "--b256" in __import__("sys").argv:
    torch.set_num_threads(8)
    unet_b256_golden(ref_loader.load())
