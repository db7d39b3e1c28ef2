"""Trainer-level parity: the fused train step against the reference algorithm's step
(oracle forward + autograd + clip_grad_norm_(0.7) + torch.optim.AdamW(eps=1e-6), improved_diffusion_trainer.py:256-322,
363-413), and the DiffusionTrainer class contract (train / validate / ddpm_sample / checkpoints, :82-126,335-655)."""
import math

import pytest
import torch

from conftest import record_metric

pytestmark = pytest.mark.gpu
UPDATE_REL_BOUND = 3e-4     # measured 1.1e-4 / 1.2e-4 (profiles/r02_parity_metrics.jsonl); round 1 allowed 5e-2


def _reference_steps(sd0, emb_coeff, latent, text, ts, noises, lrs, heads, coupled=False):
    """The reference step on CPU in fp32 from the same initial weights; returns losses, grad norms and final parameters.
    Its own OneCycleLR (torch defaults: cycle_momentum=True, i.e. beta1 cycles 0.95 -> 0.85 -> 0.95) is stepped per batch
    exactly as improved_diffusion_trainer.py:313-320,413 does; `lrs` only cross-checks the two schedules."""
    from oracle import unet_oracle
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd0.items()}
    sd = dict(params)
    sd["time_embed.emb_coeff"] = emb_coeff
    cls = torch.optim.Adam if coupled else torch.optim.AdamW
    opt = cls(list(params.values()), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-6)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=40, pct_start=0.1, anneal_strategy="cos")
    crit = torch.nn.SmoothL1Loss(beta=0.1)
    ns = NoiseScheduler()
    losses, norms = [], []
    for t, noise, lr in zip(ts, noises, lrs):
        assert abs(opt.param_groups[0]["lr"] - lr[0]) <= 1e-12 and abs(opt.param_groups[0]["betas"][0] - lr[1]) <= 1e-12
        lat = torch.clamp(latent, -3.0, 3.0)                                                    # :363
        noisy = ns.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1) * lat + ns.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1) * noise
        opt.zero_grad()
        loss = crit(unet_oracle.unet_forward(sd, noisy, t, text, num_heads=heads), noise)
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(list(params.values()), max_norm=0.7)))    # :410
        opt.step()
        sched.step()
        losses.append(loss.item())
    return losses, norms, {k: v.detach() for k, v in params.items()}


@pytest.mark.parametrize("coupled", [False, True], ids=["adamw", "adam_l2"])
def test_train_step_matches_reference_algorithm(cuda_device, coupled):
    """Three optimisation steps in the fp32 parity mode: loss, global gradient norm and the parameter update agree with
    the reference algorithm run on the CPU from the same weights, timesteps and noise -- OneCycleLR with torch's default
    cycle_momentum=True on both sides (beta1 cycling reaches the fused kernel), AdamW and the reference's coupled-decay
    `Adam` branch (improved_diffusion_trainer.py:276-292)."""
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
    opt = FusedAdamW(unet, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-4, max_grad_norm=0.7, adamw=not coupled)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=40, pct_start=0.1, anneal_strategy="cos")
    step = TrainStep(unet, NoiseScheduler().to(dev), opt, sched)
    losses, norms, lrs = [], [], []
    for t, noise in zip(ts, noises):
        lrs.append((opt.param_groups[0]["lr"], opt.param_groups[0]["betas"][0]))
        losses.append(step(latent.to(dev), text.to(dev), timesteps=t.to(dev), noise=noise.to(dev)).item())
        norms.append(opt.clip_state[0].item())
        assert opt.clip_state[2].item() == 1.0          # finite-gradient flag: the update was applied
    assert lrs[0][1] == pytest.approx(0.95) and lrs[1][1] < lrs[0][1], "OneCycleLR must be cycling beta1"
    assert opt.applied_steps() == steps
    ref_losses, ref_norms, ref_params = _reference_steps(sd0, emb_coeff, latent, text, ts, noises, lrs, heads, coupled)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-6 * abs(b), (losses, ref_losses)    # measured 5e-7
    for a, b in zip(norms, ref_norms):
        assert abs(a - b) <= 1e-3 * abs(b), (norms, ref_norms)      # measured 3.8e-4 (the reference's own fp32 CPU wgrad accuracy)
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
    rel, cos = math.sqrt(num / den), dot / math.sqrt(den * n_ours)
    print(f"[trainer {'adam_l2' if coupled else 'adamw'}] losses {losses} vs {ref_losses}; norms {norms} vs {ref_norms}; "
          f"relative update error {rel:.3e}, cosine {cos:.6f}")
    record_metric(f"trainer_{'adam_l2' if coupled else 'adamw'}", update_rel=rel, update_cos=cos,
                  loss_rel=max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)),
                  norm_rel=max(abs(a - b) / abs(b) for a, b in zip(norms, ref_norms)))
    assert den > 0 and rel <= UPDATE_REL_BOUND, f"relative update error {rel:.3e}"
    assert cos >= 0.998


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


def test_reference_adamw_state_dict_loads_element_for_element(cuda_device):
    """A `torch.optim.AdamW` state_dict (the reference's `optimizer_state_dict`) loads into FusedAdamW in the parameters'
    logical layout -- including the 3x3 conv weights that live tap-major in the flat buffers -- and one further step from
    it equals torch.optim.AdamW's step; the exported state round-trips back into torch.optim.AdamW unchanged."""
    from pokemon_sprite_generator_b200.trainer import FusedAdamW
    from pokemon_sprite_generator_b200.unet import UNet
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    unet = UNet(num_heads=4).to(dev)
    names = [n for n, _ in unet.named_parameters()]
    # the reference side: plain tensors with the same shapes, a torch AdamW that has taken two steps on random gradients
    g = torch.Generator(device="cuda").manual_seed(5)
    ref_params = [p.detach().clone().requires_grad_(True) for p in unet.parameters()]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-6)
    for _ in range(2):
        for p in ref_params:
            p.grad = torch.randn(p.shape, device=dev, generator=g) * 0.01
        ref_opt.step()
    with torch.no_grad():
        for p, r in zip(unet.parameters(), ref_params):
            p.copy_(r)
    opt = FusedAdamW(unet, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-2, max_grad_norm=0.0)
    opt.load_state_dict(ref_opt.state_dict())
    assert opt.applied_steps() == 2
    for i, (p, r) in enumerate(zip(unet.parameters(), ref_params)):
        assert torch.equal(opt.state[p]["exp_avg"], ref_opt.state[r]["exp_avg"]), names[i]
        assert torch.equal(opt.state[p]["exp_avg_sq"], ref_opt.state[r]["exp_avg_sq"]), names[i]
    # one more step on both sides with the same gradients, written into the engine's flat gradient buffer
    store = unet.engine().store
    grads = [torch.randn(p.shape, device=dev, generator=g) * 0.01 for p in ref_params]
    for (name, p), gr in zip(unet.named_parameters(), grads):
        store.grad_of(p).copy_(gr)
    for r, gr in zip(ref_params, grads):
        r.grad = gr
    ref_opt.step()
    opt.step()
    worst = 0.0
    for i, (p, r) in enumerate(zip(unet.parameters(), ref_params)):
        d = (p.detach() - r.detach()).abs().max().item()
        worst = max(worst, d)
        assert d <= 2e-6, (names[i], d)      # lr 1e-3 * O(1) update, fp32 rounding of the fused vs foreach formulation
        assert torch.allclose(opt.state[p]["exp_avg"], ref_opt.state[r]["exp_avg"], rtol=1e-5, atol=1e-8), names[i]
        assert torch.allclose(opt.state[p]["exp_avg_sq"], ref_opt.state[r]["exp_avg_sq"], rtol=1e-5, atol=1e-12), names[i]
    print(f"[adamw state] worst parameter difference after the step: {worst:.3e}")
    # export -> a fresh torch AdamW accepts it and carries the same moments
    back = torch.optim.AdamW([p.detach().clone().requires_grad_(True) for p in unet.parameters()], lr=1e-3)
    back.load_state_dict(opt.state_dict())
    for (p, st_b) in zip(unet.parameters(), back.state.values()):
        assert torch.equal(st_b["exp_avg_sq"], opt.state[p]["exp_avg_sq"]) and float(st_b["step"]) == 3.0


def test_skipped_step_does_not_advance_optimizer_or_schedule(cuda_device):
    """Non-finite gradients: the device skips the update AND the applied-step counter (Adam's bias correction); the host LR
    schedule is held back one tick as soon as the flag has landed (reference: `continue` before optimizer.step() /
    scheduler.step(), improved_diffusion_trainer.py:395-397,413)."""
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
    from pokemon_sprite_generator_b200.unet import UNet
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    unet = UNet(num_heads=4).to(dev).eval()
    opt = FusedAdamW(unet, lr=1e-4, max_grad_norm=0.7)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=40, pct_start=0.1, anneal_strategy="cos")
    step = TrainStep(unet, NoiseScheduler().to(dev), opt, sched)
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = torch.randn(2, 8, 27, 27, device=dev, generator=g)
    text = torch.randn(2, 32, 256, device=dev, generator=g)
    step(lat, text)
    before = {k: v.detach().clone() for k, v in unet.named_parameters()}
    bad_text = text.clone()
    bad_text[0, 0, 0] = float("inf")
    step(lat, bad_text)                       # non-finite activations -> non-finite gradients -> skipped on the device
    assert opt.clip_state[2].item() == 0.0 and opt.applied_steps() == 1
    for k, v in unet.named_parameters():
        assert torch.equal(v.detach(), before[k]), k
    step(lat, text)                           # the host now knows: this step's scheduler tick is withheld
    assert opt.applied_steps() == 2 and step.skipped_steps == 1
    assert sched.last_epoch == 2, sched.last_epoch     # 3 calls, 2 applied steps


def test_validate_epoch_matches_oracle(cuda_device, tmp_path):
    """validate_epoch (reference :447-506) with injected timestep / noise draws equals the oracle's loss on the same tensors."""
    from oracle import unet_oracle
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.trainer import DiffusionTrainer
    cfg = {"experiment_dir": str(tmp_path), "model": {"latent_dim": 8, "text_embedding_dim": 256, "num_heads": 4},
           "training": {"diffusion_epochs": 1}, "unet_optimization": {"scheduler": "constant"}}
    loaders = _loaders(2, 2, 9)
    loaders["val"] = loaders["train"]
    torch.manual_seed(0)
    tr = DiffusionTrainer(cfg, None, "v0", components={"data_loaders": loaders}, compute_dtype=torch.float32)
    g = torch.Generator().manual_seed(17)
    drawn = []

    def draws(latent):
        t = torch.randint(0, 1000, (latent.shape[0],), generator=g)
        n = torch.randn(latent.shape, generator=g)
        drawn.append((t, n))
        return t, n

    got = tr.validate_epoch(0, draws=draws)["val_loss"]
    sd = {k: v.detach().cpu() for k, v in tr.unet.state_dict().items()}
    ns = NoiseScheduler()
    want = []
    for batch, (t, n) in zip(loaders["val"], drawn):
        lat = torch.clamp(batch["latent"], -3.0, 3.0)
        noisy = ns.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1) * lat + ns.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1) * n
        with torch.no_grad():
            pred = unet_oracle.unet_forward(sd, noisy, t, batch["text_emb"], num_heads=4)
        want.append(torch.nn.functional.smooth_l1_loss(pred, n, beta=0.1).item())
    want = sum(want) / len(want)
    print(f"[validate_epoch] {got:.7f} vs oracle {want:.7f}")
    assert abs(got - want) <= 1e-5 * want
