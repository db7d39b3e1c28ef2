"""Trainer-level parity: the fused train step against the reference algorithm's step
(oracle forward + autograd + clip_grad_norm_(0.7) + torch.optim.AdamW(eps=1e-6), improved_diffusion_trainer.py:256-322,
363-413), and the DiffusionTrainer class contract (train / validate / ddpm_sample / checkpoints, :82-126,335-655)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_steps(sd0, emb_coeff, latent, text, ts, noises, lrs, heads):
    """The reference step on CPU in fp32 from the same initial weights; returns losses, grad norms and final parameters."""
    from oracle import unet_oracle
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd0.items()}
    sd = dict(params)
    sd["time_embed.emb_coeff"] = emb_coeff
    opt = torch.optim.AdamW(list(params.values()), lr=lrs[0], betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-6)
    crit = torch.nn.SmoothL1Loss(beta=0.1)
    ns = NoiseScheduler()
    losses, norms = [], []
    for t, noise, lr in zip(ts, noises, lrs):
        for g in opt.param_groups:
            g["lr"] = lr
        lat = torch.clamp(latent, -3.0, 3.0)                                                    # :363
        noisy = ns.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1) * lat + ns.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1) * noise
        opt.zero_grad()
        loss = crit(unet_oracle.unet_forward(sd, noisy, t, text, num_heads=heads), noise)
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(list(params.values()), max_norm=0.7)))    # :410
        opt.step()
        losses.append(loss.item())
    return losses, norms, {k: v.detach() for k, v in params.items()}


def test_train_step_matches_reference_algorithm(cuda_device):
    """Three optimisation steps in the fp32 parity mode: loss, global gradient norm and the parameter update agree with
    the reference algorithm run on the CPU from the same weights, timesteps and noise."""
    from oracle import inputs
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
    from pokemon_sprite_generator_b200.unet import UNet
    dev = torch.device("cuda:0")
    heads, steps = 4, 3
    torch.manual_seed(0)
    unet = UNet(num_heads=heads, compute_dtype=torch.float32)
    with torch.no_grad():       # O(1)-gain re-init of the 0.02-gain layers: gradients well above Adam's eps everywhere
        for name, p in unet.named_parameters():
            if p.dim() >= 2 and float(p.std()) < 5e-3:
                p.mul_(20.0)
    sd0 = {k: v.detach().clone() for k, v in unet.named_parameters()}
    emb_coeff = unet.time_embed.emb_coeff.detach().clone()
    unet = unet.to(dev).eval()          # eval: dropout off (parity is defined without dropout)
    latent, text, _, _ = inputs.make_inputs(2, 32, 1234)
    g = torch.Generator().manual_seed(99)
    ts = [torch.randint(0, 1000, (2,), generator=g) for _ in range(steps)]
    noises = [torch.randn(2, 8, 27, 27, generator=g) for _ in range(steps)]
    opt = FusedAdamW(unet, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-4, max_grad_norm=0.7)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=40, pct_start=0.1, anneal_strategy="cos", cycle_momentum=False)
    step = TrainStep(unet, NoiseScheduler().to(dev), opt, sched)
    losses, norms, lrs = [], [], []
    for t, noise in zip(ts, noises):
        lrs.append(opt.param_groups[0]["lr"])
        losses.append(step(latent.to(dev), text.to(dev), timesteps=t.to(dev), noise=noise.to(dev)).item())
        norms.append(opt.clip_state[0].item())
        assert opt.clip_state[2].item() == 1.0          # finite-gradient flag: the update was applied
    ref_losses, ref_norms, ref_params = _reference_steps(sd0, emb_coeff, latent, text, ts, noises, lrs, heads)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, ref_losses)
    for a, b in zip(norms, ref_norms):
        assert abs(a - b) <= 5e-3 * abs(b), (norms, ref_norms)
    # parameter updates: compared as one vector over the whole model (Adam's g / (sqrt(v) + eps) amplifies fp32 noise on the
    # few entries whose gradient is near eps, so the bound is on the update's direction and size, not element by element)
    num = den = dot = n_ours = 0.0
    for k, p in unet.named_parameters():
        d_ours = (p.detach().cpu().double() - sd0[k].double()).flatten()
        d_ref = (ref_params[k].double() - sd0[k].double()).flatten()
        num += float((d_ours - d_ref).pow(2).sum())
        den += float(d_ref.pow(2).sum())
        dot += float((d_ours * d_ref).sum())
        n_ours += float(d_ours.pow(2).sum())
    assert den > 0 and math.sqrt(num / den) <= 0.05, f"relative update error {math.sqrt(num / den):.3e}"
    assert dot / math.sqrt(den * n_ours) >= 0.998


def _loaders(n_batches, batch, seed):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: [{"latent": torch.randn(batch, 8, 27, 27, generator=g), "text_emb": torch.randn(batch, 32, 256, generator=g)}   # noqa: E731
                  for _ in range(n_batches)]
    return {"train": mk(), "val": mk()[:1], "test": []}


def test_diffusion_trainer_contract(cuda_device, tmp_path):
    """DiffusionTrainer(config, vae_checkpoint_path, experiment_name): attributes, train(), validate_epoch, ddpm_sample and a
    checkpoint with the reference's keys that restores weights and optimiser state exactly."""
    from pokemon_sprite_generator_b200.trainer import DiffusionTrainer
    cfg = {"experiment_dir": str(tmp_path), "model": {"latent_dim": 8, "text_embedding_dim": 256, "num_heads": 4},
           "training": {"diffusion_epochs": 1, "save_every": 1, "sample_every": 10, "log_every": 1},
           "unet_optimization": {"learning_rate": 1e-4, "weight_decay": 1e-4, "max_grad_norm": 0.7, "scheduler": "cosine"}}
    torch.manual_seed(0)
    tr = DiffusionTrainer(cfg, None, "t0", components={"data_loaders": _loaders(12, 2, 5)})
    for attr in ("unet", "noise_scheduler", "optimizer", "scheduler", "criterion", "data_loaders", "device", "global_step",
                 "current_epoch", "best_val_loss", "writer", "logger"):
        assert hasattr(tr, attr), attr
    tr.train()
    assert tr.global_step == 12 and math.isfinite(tr.best_val_loss)
    path = tr.checkpoint_dir / "diffusion_best_model.pth"
    assert path.exists()
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ckpt) == {"epoch", "global_step", "unet_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_val_loss", "config"}
    assert len(ckpt["unet_state_dict"]) == 479
    m = tr.validate_epoch(0)
    assert math.isfinite(m["val_loss"]) and 0.0 < m["val_loss"] < 2.0
    # a fresh trainer resumes from the checkpoint: same weights, same optimiser moments, and the same next step
    torch.manual_seed(1)
    tr2 = DiffusionTrainer(cfg, None, "t1", components={"data_loaders": _loaders(12, 2, 5)})
    tr2.load_checkpoint(str(path))
    assert tr2.global_step == 12 and tr2.current_epoch == 0
    for (k, a), (_, b) in zip(tr.unet.state_dict().items(), tr2.unet.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = tr.optimizer.state_dict()["state"], tr2.optimizer.state_dict()["state"]
    assert all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) and torch.equal(s1[i]["exp_avg_sq"], s2[i]["exp_avg_sq"]) for i in s1)
    tr.unet.eval(); tr2.unet.eval()
    tr._step.lr_sched = tr2._step.lr_sched = None      # (the 12-step OneCycleLR schedule is exhausted)
    batch = tr.data_loaders["train"][0]
    g = torch.Generator(device="cuda").manual_seed(3)
    t = torch.randint(0, 1000, (2,), device="cuda", generator=g)
    noise = torch.randn(2, 8, 27, 27, device="cuda", generator=g)
    l1 = tr._step(batch["latent"].cuda(), batch["text_emb"].cuda(), timesteps=t, noise=noise)
    l2 = tr2._step(batch["latent"].cuda(), batch["text_emb"].cuda(), timesteps=t, noise=noise)
    assert l1.item() == l2.item()
    for (k, a), (_, b) in zip(tr.unet.state_dict().items(), tr2.unet.state_dict().items()):
        assert torch.equal(a, b), k
    # sampling through the trainer API (fast schedule: 20 reverse steps, reference :508-569)
    x = tr.ddpm_sample(batch["text_emb"][:2], 2, fast_sampling=True)
    assert x.shape == (2, 8, 27, 27) and bool(torch.isfinite(x).all())
