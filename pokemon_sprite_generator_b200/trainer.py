"""Drop-in for the reference's stage-2 diffusion trainer (`src.training.DiffusionTrainer`, i.e.
`ImprovedDiffusionTrainer`, src/training/improved_diffusion_trainer.py:77-692).

Same constructor `(config, vae_checkpoint_path, experiment_name)`, same public methods (`train`, `train_epoch`,
`validate_epoch`, `ddpm_sample`, `generate_samples`, `save_checkpoint`, `load_checkpoint`) and attributes (`unet`,
`noise_scheduler`, `optimizer`, `scheduler`, `criterion`, `data_loaders`, `device`, `global_step`, ...).

What changes underneath:
  * the step is `TrainStep`: q_sample -> U-Net fwd -> SmoothL1 fwd+bwd -> U-Net bwd -> (NCCL all-reduce) -> global-norm
    clip -> AdamW, all hand-written kernels on flat fp32 buffers, with no host synchronisation inside the step: the
    reference's five NaN checks and 478 `.item()` norm reads (:328-333,353-404) become one device-side finite flag
    folded into the clip coefficient (non-finite gradients skip the update, as the reference skips the batch);
  * data parallelism: one process per GPU, gradients averaged with one all-reduce over the flat buffer;
  * the frozen VAE encoder / text encoder / dataset are out of scope: they are imported from the
    reference package when it is importable, or injected (`components=`) -- e.g. synthetic latents for benchmarks;
    the VAE decoder that `generate_samples` needs and the VAE encoder are this package's own drop-ins (`vae.VAEDecoder`,
    `vae.VAEEncoder`); `config['data']['latent_cache'] = True` encodes the training set once
    (`vae.LatentCache`) instead of re-encoding every image every epoch.
"""
from __future__ import annotations

import logging
import math
import os
from pathlib import Path
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops as K
from .losses import SmoothL1Loss, smooth_l1_fwd_bwd
from .parallel import GradSync, allreduce_mean_
from .scheduler import NoiseScheduler
from .unet import UNet


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the sqrt) -- or, with
    `adamw=False`, torch.optim.Adam's coupled L2 decay (the reference's `else` branch, improved_diffusion_trainer.py:285-292)
    -- as ONE kernel over the U-Net's flat parameter / gradient buffers, fused with the global-norm clip coefficient and the
    skip-on-non-finite predicate.  Exposes param_groups / state_dict like a torch optimizer so LR schedulers (OneCycleLR,
    which also cycles beta1) and checkpoints keep working: `state[p]['exp_avg' / 'exp_avg_sq']` are views of the flat
    moment buffers with the parameter's LOGICAL layout (tap-major conv weights through the same permute as the
    parameter itself), so a reference `torch.optim.AdamW` state_dict loads and exports element for element.

    Step counting: the device keeps the number of APPLIED steps (`clip_state[3]`, incremented only for finite gradients)
    and the kernel derives Adam's bias corrections from it, so a skipped batch does not advance them -- the reference
    `continue`s before `optimizer.step()` (:395-397).  `applied_steps()` reads it back (one sync; used by state_dict)."""

    def __init__(self, unet: UNet, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-4, max_grad_norm: float = 0.0,
                 adamw: bool = True):
        params = list(unet.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.unet = unet
        self.max_grad_norm = max_grad_norm
        self.decoupled = adamw
        self._m = self._v = None
        self._sumsq = None
        self.clip_state = None      # device [4]: total_norm, clip coefficient, finite flag, applied steps
        self.step_count = 0         # step() calls issued (host side); see applied_steps() for the device-side truth
        self._applied0 = 0          # applied steps to seed the device counter with when the state is (re)built

    def _ensure_state(self):
        eng = self.unet.engine()
        store = eng.store
        if store.flat is None:      # no forward yet (e.g. load_state_dict on a fresh trainer): build the flat buffers now
            dev = next(self.unet.parameters()).device
            if dev.type != "cuda":
                raise L.PsgError("FusedAdamW needs the U-Net on a CUDA device (there is no CPU path)")
            eng.prepare(dev)
        if self._m is None or self._m.device != store.flat.device or self._m.numel() != store.total:
            dev = store.flat.device
            self._m = torch.zeros(store.total, dtype=torch.float32, device=dev)
            self._v = torch.zeros(store.total, dtype=torch.float32, device=dev)
            self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
            self.clip_state = torch.zeros(4, dtype=torch.float32, device=dev)
            self.clip_state[3] = float(self._applied0)
            for name, p in store.named:
                off = store.offsets[name]
                self.state[p] = {"step": torch.tensor(float(self._applied0)), "exp_avg": store._view(self._m, p, off),
                                 "exp_avg_sq": store._view(self._v, p, off)}
        return store

    def applied_steps(self) -> int:
        """Number of optimiser steps actually applied (non-finite batches are skipped on the device).  Synchronises."""
        if self.clip_state is None:
            return self._applied0
        return int(round(float(self.clip_state[3].item())))

    @torch.no_grad()
    def step(self, closure=None):
        eng = self.unet.engine()
        store = self._ensure_state()
        g = self.param_groups[0]
        self.step_count += 1
        K.sumsq(store.grads, self._sumsq)
        K.clip_coef(self._sumsq, float(self.max_grad_norm or 0.0), self.clip_state, count_steps=True)
        K.adamw_step(store.flat, store.grads, self._m, self._v, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                     0, self.clip_state, shadow=store.shadow, coupled_l2=not self.decoupled)
        if store.shadow is not None:
            eng.mark_shadow_fresh()     # the kernel rewrote the bf16 shadow the GEMMs read
        else:
            eng.mark_params_dirty()
        return None

    def zero_grad(self, set_to_none: bool = True):
        for _, p in self.unet.engine().store.named:
            p.grad = None

    def state_dict(self):
        self._ensure_state()
        n = float(self.applied_steps())
        for st in self.state.values():
            st["step"] = torch.tensor(n)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch replaced the per-parameter state tensors: fold them back into the flat buffers (element for element in the
        # parameter's logical layout -- the views permute tap-major conv weights exactly like the parameter view does)
        old = {p: dict(st) for p, st in self.state.items()}
        steps = [int(float(st["step"])) for st in old.values() if "step" in st]
        self._applied0 = max(steps) if steps else 0
        self.step_count = self._applied0
        self._m = None
        self._ensure_state()
        for p, st in old.items():
            if "exp_avg" in st:
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])


class TrainStep:
    """One data-parallel optimisation step on device-resident tensors (the unit bench.py times).

    Data parallel (world > 1): rank 0's parameters and optimiser moments are broadcast once at construction, so replicas
    start identical whatever each rank's RNG state was at `UNet()` time; from then on identical averaged gradients keep
    them identical.  Skipped (non-finite) batches: the device skips the update and does not count the step; the host LR
    scheduler learns about it one step late through an asynchronous copy of the finite flag (no synchronisation inside the
    step) and then holds the schedule back by one tick, so schedule position == applied steps again from the next step on
    (the reference `continue`s before `scheduler.step()`, improved_diffusion_trainer.py:395-397,413)."""

    def __init__(self, unet: UNet, noise_scheduler: NoiseScheduler, optimizer: FusedAdamW, lr_scheduler=None, beta: float = 0.1,
                 clamp: float = 3.0, process_group=None):
        self.unet, self.ns, self.opt, self.lr_sched = unet, noise_scheduler, optimizer, lr_scheduler
        self.beta, self.clamp = beta, clamp
        self.pg = process_group
        self.buckets = 8
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # overlap="backward": bucketed all-reduce issued from inside backward (parallel.GradSync); "none": after backward
        self.overlap = os.environ.get("PSG_GRAD_OVERLAP", "backward")
        self.grad_sync = None
        self._flag_host = None      # pinned copy of the previous step's finite flag + the event that says it has landed
        self._flag_event = None
        self.skipped_steps = 0
        if self.world > 1 and self.overlap == "backward":
            reserve = int(os.environ.get("PSG_COMM_SMS", os.environ.get("NCCL_MAX_CTAS", "16")))
            lib = L.load()
            self.grad_sync = GradSync(process_group, bucket_bytes=int(os.environ.get("PSG_BUCKET_MB", "128")) << 20,
                                      reserve_sms=reserve, reserve_hook=lambda n: lib.psg_umma_reserve_sms(int(n)), prescaled=True,
                                      window_entries=int(os.environ.get("PSG_COMM_WINDOW", "4")))
        if self.world > 1:
            self.sync_replicas()

    def sync_replicas(self, src: int = 0) -> None:
        """Broadcast rank `src`'s flat parameters and AdamW state to every replica (start-up / after load_checkpoint)."""
        dev = next(self.unet.parameters()).device
        if dev.type != "cuda":
            return      # the U-Net has not been moved to its GPU yet; the owner calls sync_replicas() once it has
        eng = self.unet.engine()
        eng.prepare(dev)
        store = self.opt._ensure_state()
        for buf in (store.flat, self.opt._m, self.opt._v, self.opt.clip_state):
            dist.broadcast(buf, src=src, group=self.pg)
        eng.mark_params_dirty()

    def _previous_step_was_skipped(self) -> bool:
        if self._flag_event is None:
            return False
        self._flag_event.synchronize()      # recorded a whole step ago: never waits in steady state
        return float(self._flag_host[0]) == 0.0

    def __call__(self, latent: torch.Tensor, text_emb: torch.Tensor, timesteps: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """latent [B,8,27,27] fp32, text_emb [B,L,256] fp32 on the GPU.  Returns the (device, 0-dim) loss.
        timesteps / noise default to the reference's draws (improved_diffusion_trainer.py:366-373)."""
        eng = self.unet.engine()
        B = latent.shape[0]
        if timesteps is None:
            timesteps = torch.randint(0, self.ns.num_timesteps, (B,), device=latent.device)
        if noise is None:
            noise = torch.randn_like(latent)
        noisy = self.ns.add_noise(latent, noise, timesteps, clamp=self.clamp)          # clamp(+-3) fused (:363)
        pred, ctx = eng.forward(noisy, timesteps, text_emb, need_grad=True)
        loss, dpred = smooth_l1_fwd_bwd(pred, noise, beta=self.beta, grad_scale=1.0 / self.world)
        if self.grad_sync is not None:
            eng.backward(ctx, dpred, grad_sync=self.grad_sync)      # gradient buckets reduced while backward still runs
        else:
            eng.backward(ctx, dpred)
            if self.world > 1:
                allreduce_mean_(eng.store.grads, group=self.pg, prescaled=True, buckets=self.buckets)
        skipped_before = self._previous_step_was_skipped()
        self.opt.step()
        if self._flag_host is None:
            self._flag_host = torch.ones(1, dtype=torch.float32).pin_memory()
            self._flag_event = torch.cuda.Event()
        self._flag_host.copy_(self.opt.clip_state[2:3], non_blocking=True)
        self._flag_event.record()
        if skipped_before:
            self.skipped_steps += 1
        elif self.lr_sched is not None:
            self.lr_sched.step()
        return loss


def _get_device() -> torch.device:
    if not torch.cuda.is_available():
        raise L.PsgError("psg_b200 DiffusionTrainer needs a CUDA device (sm_100a); there is no CPU/MPS path")
    return torch.device("cuda", torch.cuda.current_device())


class DiffusionTrainer:
    """See module docstring.  `components` may provide: text_encoder(list[str]) -> [B,L,D], vae_encoder(images) ->
    (latent, mu, logvar), vae_decoder(latent, text_emb) -> images, data_loaders {'train','val','test'}."""

    def __init__(self, config: Dict[str, Any], vae_checkpoint_path: Optional[str] = None, experiment_name: str = "pokemon_diffusion",
                 *, components: Optional[Dict[str, Any]] = None, compute_dtype: torch.dtype = torch.bfloat16):
        self.config = config
        self.vae_checkpoint_path = vae_checkpoint_path
        self.experiment_name = experiment_name
        self.components = dict(components or {})
        self.compute_dtype = compute_dtype
        self.device = _get_device()
        self.setup_directories()
        self.setup_logging()
        self.setup_models()
        self.setup_data_loaders()
        self.setup_optimization()
        self.setup_scheduler()
        self.setup_monitoring()
        self.current_epoch = 0
        self.global_step = 0
        self.best_val_loss = float("inf")

    # ---- set-up (reference :128-326) -----------------------------------------------------------------------------
    def setup_directories(self):
        self.experiment_dir = Path(self.config.get("experiment_dir", "experiments")) / self.experiment_name
        self.checkpoint_dir = self.experiment_dir / "checkpoints"
        self.log_dir = self.experiment_dir / "logs"
        self.sample_dir = self.experiment_dir / "samples"
        for d in (self.experiment_dir, self.checkpoint_dir, self.log_dir, self.sample_dir):
            d.mkdir(parents=True, exist_ok=True)

    def setup_logging(self):
        self.logger = logging.getLogger(f"psg_b200.trainer.{self.experiment_name}")
        if not self.logger.handlers:
            self.logger.setLevel(logging.INFO)
            fmt = logging.Formatter("%(asctime)s - %(levelname)s - %(message)s")
            for h in (logging.FileHandler(self.log_dir / "diffusion_training.log"), logging.StreamHandler()):
                h.setFormatter(fmt)
                self.logger.addHandler(h)

    def _reference_component(self, name: str):
        """Frozen VAE / text encoder / dataset come from the reference package when it is importable."""
        try:
            import importlib
            mod = importlib.import_module("src.models" if name != "create_data_loaders" else "src.data")
            return getattr(mod, name)
        except Exception as e:  # pragma: no cover - depends on the host repo
            raise L.PsgError(f"component '{name}' was not injected and the reference package is not importable: {e}")

    def setup_models(self):
        mc = self.config.get("model", {})
        self.text_encoder = self.components.get("text_encoder")
        self.vae_encoder = self.components.get("vae_encoder")
        self.vae_decoder = self.components.get("vae_decoder")
        if self.text_encoder is None and "data_loaders" not in self.components:
            from .text_encoder import TextEncoder      # drop-in; needs the BERT files like the reference's
            self.text_encoder = TextEncoder(model_name=mc["bert_model"], hidden_dim=mc["text_embedding_dim"]).to(self.device).eval()
        if self.vae_encoder is None and self.vae_checkpoint_path is not None:
            from .vae import VAEDecoder, VAEEncoder      # drop-ins (same state_dicts as the reference's), on the CUDA kernels
            enc = VAEEncoder(input_channels=3, latent_dim=mc.get("latent_dim", 8), compute_dtype=torch.float32).to(self.device)
            dec = VAEDecoder(latent_dim=mc.get("latent_dim", 8), text_dim=mc["text_embedding_dim"], output_channels=3).to(self.device)
            ckpt = torch.load(self.vae_checkpoint_path, map_location=self.device)
            if "vae_state_dict" in ckpt:
                sd = ckpt["vae_state_dict"]
                enc.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=False)
                dec.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=False)
            if "text_encoder_state_dict" in ckpt and hasattr(self.text_encoder, "load_state_dict"):
                self.text_encoder.load_state_dict(ckpt["text_encoder_state_dict"], strict=False)
            for m in (enc, dec):
                for p in m.parameters():
                    p.requires_grad = False
                m.eval()
            self.vae_encoder, self.vae_decoder = enc, dec
        self.unet = self.components.get("unet")         # an already-built (e.g. warmed-up) drop-in U-Net may be injected
        if self.unet is None:
            self.unet = UNet(latent_dim=mc.get("latent_dim", 8), text_dim=mc.get("text_embedding_dim", 256),
                             time_emb_dim=mc.get("time_emb_dim", 128), num_heads=mc.get("num_heads", 4),
                             compute_dtype=self.compute_dtype).to(self.device)
        self.noise_scheduler = NoiseScheduler(num_timesteps=mc.get("num_timesteps", 1000), beta_start=mc.get("beta_start", 0.0001),
                                              beta_end=mc.get("beta_end", 0.02))
        self.logger.info(f"U-Net initialized with {sum(p.numel() for p in self.unet.parameters())} parameters")

    def setup_data_loaders(self):
        if "data_loaders" in self.components:
            self.data_loaders = self.components["data_loaders"]
            return
        dc = self.config["data"]
        uc = self.config.get("unet_optimization", {})
        tr, va, te = self._reference_component("create_data_loaders")(
            csv_path=dc["csv_path"], image_dir=dc["image_dir"], batch_size=uc.get("batch_size", dc["batch_size"]),
            val_split=dc["val_split"], test_split=dc["test_split"], image_size=dc["image_size"],
            num_workers=uc.get("num_workers", dc["num_workers"]), pin_memory=dc["pin_memory"])
        self.data_loaders = {"train": tr, "val": va, "test": te}
        if dc.get("latent_cache", False) and self.vae_encoder is not None and self.text_encoder is not None:
            # encode the training set ONCE; every epoch re-samples latents from the cached (mu, logvar) on the device (vae.LatentCache)
            from .vae import LatentCache
            bs = uc.get("batch_size", dc["batch_size"])
            self.data_loaders["train"] = LatentCache.build(tr, self.vae_encoder, self.text_encoder, self.device, bs, shuffle=True)
            self.logger.info(f"latent cache: {self.data_loaders['train'].mu.shape[0]} training images encoded once")

    def setup_optimization(self):
        uc = self.config.get("unet_optimization", {})
        oc = self.config.get("optimization", {})
        get = lambda k, d: uc.get(k, oc.get(k, d))   # noqa: E731  (README values as defaults)
        lr = get("learning_rate", 1e-4)
        self.max_grad_norm = get("max_grad_norm", 0.7)
        opt_type = get("optimizer", "adamw")
        self.scheduler_config = {"type": get("scheduler", "cosine"), "lr": lr}
        self.criterion = SmoothL1Loss(beta=0.1)
        if "optimizer" in self.components:
            self.optimizer = self.components["optimizer"]
            return
        self.optimizer = FusedAdamW(self.unet, lr=lr, betas=(get("beta1", 0.9), get("beta2", 0.999)), eps=1e-6,
                                    weight_decay=get("weight_decay", 1e-4), max_grad_norm=self.max_grad_norm,
                                    adamw=(opt_type == "adamw"))
        self.scheduler_config = {"type": get("scheduler", "cosine"), "lr": lr}
        self.criterion = SmoothL1Loss(beta=0.1)

    def setup_scheduler(self):
        lr = self.scheduler_config["lr"]
        if "lr_scheduler" in self.components:
            self.scheduler = self.components["lr_scheduler"]
        elif self.scheduler_config["type"] == "cosine":
            total = self.config.get("training", {}).get("diffusion_epochs", 1) * max(1, len(self.data_loaders["train"]))
            self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, max_lr=lr, total_steps=total, pct_start=0.1,
                                                                 anneal_strategy="cos")
        else:
            self.scheduler = torch.optim.lr_scheduler.ConstantLR(self.optimizer, factor=1.0)
        self._step = self.components.get("train_step") or TrainStep(self.unet, self.noise_scheduler, self.optimizer, self.scheduler)
        self.world = self._step.world
        self.rank = dist.get_rank() if self.world > 1 else 0

    def setup_monitoring(self):
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.log_dir)
        except Exception:  # tensorboard is optional
            self.writer = None

    # ---- batches -------------------------------------------------------------------------------------------------
    def _encode(self, batch):
        """batch -> (latent fp32 [B,8,27,27], text_emb fp32 [B,L,D]) on the device.  Accepts the reference's dict
        batches ({'image', 'full_description'}) or pre-encoded {'latent', 'text_emb'} batches."""
        if "latent" in batch:
            return batch["latent"].to(self.device, non_blocking=True), batch["text_emb"].to(self.device, non_blocking=True)
        with torch.no_grad():
            text_emb = self.text_encoder(batch["full_description"])
            latent, _, _ = self.vae_encoder(batch["image"].to(self.device))
        return latent.float(), text_emb.float()

    def train_step(self, latent: torch.Tensor, text_emb: torch.Tensor) -> torch.Tensor:
        """Public single-step API: host or device tensors in, device loss out."""
        self.unet.train()
        return self._step(latent.to(self.device, non_blocking=True), text_emb.to(self.device, non_blocking=True))

    # ---- epochs (reference :335-506) -----------------------------------------------------------------------------
    def train_epoch(self, epoch: int) -> Dict[str, float]:
        self.unet.train()
        losses = []
        log_every = self.config.get("training", {}).get("log_every", 50)
        for batch_idx, batch in enumerate(self.data_loaders["train"]):
            if self.world > 1 and not self.config.get("data", {}).get("presharded", False) and batch_idx % self.world != self.rank:
                continue        # data parallel over an unsharded (reference) loader: rank r takes batches r, r + world, ...
            latent, text_emb = self._encode(batch)
            loss = self._step(latent, text_emb)
            losses.append(loss)
            self.global_step += 1
            if self.writer is not None and batch_idx % log_every == 0:
                self.writer.add_scalar("Diffusion Train/Loss", loss.item(), self.global_step)
                self.writer.add_scalar("Diffusion Train/Learning_Rate", self.optimizer.param_groups[0]["lr"], self.global_step)
                self.writer.add_scalar("Diffusion Train/Gradient_Norm", self.optimizer.clip_state[0].item(), self.global_step)
        if not losses:
            self.logger.error("No valid batches processed!")
            return {"train_loss": float("inf")}
        stacked = torch.stack(losses)
        finite = torch.isfinite(stacked)
        K.check_kernel_timeouts()      # (synchronises, like the .item() below) a protocol fault inside a kernel must not pass silently
        if not bool(finite.any()):
            return {"train_loss": float("inf")}
        avg = stacked[finite].mean().item()
        self.logger.info(f"Epoch {epoch}: Average loss = {avg:.6f}, NaN batches = {int((~finite).sum())}, "
                         f"LR = {self.optimizer.param_groups[0]['lr']:.2e}")
        return {"train_loss": avg}

    @torch.no_grad()
    def validate_epoch(self, epoch: int, draws: Optional[Callable[[torch.Tensor], tuple]] = None) -> Dict[str, float]:
        """reference :447-506.  `draws(latent) -> (timesteps, noise)` overrides the per-batch random draws (parity tests
        inject the tensors the oracle uses); by default they come from the device generator in the reference's order."""
        self.unet.eval()
        losses = []
        for batch in self.data_loaders["val"]:
            latent, text_emb = self._encode(batch)
            if draws is not None:
                t, noise = draws(latent)
                t, noise = t.to(self.device), noise.to(self.device)
            else:
                t = torch.randint(0, self.noise_scheduler.num_timesteps, (latent.shape[0],), device=self.device)
                noise = torch.randn_like(latent)
            noisy = self.noise_scheduler.add_noise(latent, noise, t, clamp=3.0)
            pred = self.unet(noisy, t, text_emb)
            loss, _ = smooth_l1_fwd_bwd(pred, noise, beta=0.1, want_grad=False)
            losses.append(loss)
        if not losses:
            return {"val_loss": float("inf")}
        stacked = torch.stack(losses)
        finite = torch.isfinite(stacked)
        if not bool(finite.any()):
            return {"val_loss": float("inf")}
        avg = stacked[finite].mean().item()
        if self.writer is not None:
            self.writer.add_scalar("Diffusion Val/Loss", avg, epoch)
        return {"val_loss": avg}

    # ---- sampling (reference :508-569) ---------------------------------------------------------------------------
    @torch.no_grad()
    def ddpm_sample(self, text_emb: torch.Tensor, num_samples: int, fast_sampling: bool = True,
                    noise_fn: Optional[Callable[[tuple], torch.Tensor]] = None) -> torch.Tensor:
        from .sampler import ddpm_sample
        return ddpm_sample(self.unet, self.noise_scheduler, text_emb.to(self.device), num_samples, fast_sampling=fast_sampling,
                           latent_dim=self.config.get("model", {}).get("latent_dim", 8), noise_fn=noise_fn)

    @torch.no_grad()
    def generate_samples(self, epoch: int, num_samples: int = 8):
        if self.vae_decoder is None:
            raise L.PsgError("generate_samples needs a VAE decoder (components['vae_decoder'] or a VAE checkpoint)")
        self.unet.eval()
        batch = next(iter(self.data_loaders["val"]))
        _, text_emb = self._encode(batch)
        text_emb = text_emb[:num_samples]
        images = []
        for i in range(0, text_emb.shape[0], 4):
            te = text_emb[i:i + 4]
            lat = self.ddpm_sample(te, te.shape[0])
            img = torch.clamp((self.vae_decoder(lat, te) + 1.0) / 2.0, 0, 1)
            images.append(img)
            if self.writer is not None:
                for j, im in enumerate(img):
                    self.writer.add_image(f"Diffusion Generated/Sample_{i + j}", im.cpu(), epoch)
        self.logger.info(f"Generated {text_emb.shape[0]} samples for epoch {epoch}")
        return torch.cat(images) if images else None

    # ---- checkpoints (reference :617-655) ------------------------------------------------------------------------
    def save_checkpoint(self, epoch: int, is_best: bool = False):
        ckpt = {"epoch": epoch, "global_step": self.global_step,
                "unet_state_dict": {k: v.detach().clone() for k, v in self.unet.state_dict().items()},
                "optimizer_state_dict": self.optimizer.state_dict(), "scheduler_state_dict": self.scheduler.state_dict(),
                "best_val_loss": self.best_val_loss, "config": self.config}
        if is_best and self.rank == 0:      # replicas are identical: one writer
            torch.save(ckpt, self.checkpoint_dir / "diffusion_best_model.pth")
            self.logger.info(f"New best model saved at epoch {epoch}")

    def load_checkpoint(self, checkpoint_path: str):
        ckpt = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
        self.current_epoch = ckpt["epoch"]
        self.global_step = ckpt["global_step"]
        self.best_val_loss = ckpt["best_val_loss"]
        self.unet.load_state_dict(ckpt["unet_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if self.scheduler and ckpt.get("scheduler_state_dict"):
            self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        self.unet.engine().step_counter = int(self.global_step)      # dropout masks continue where the saved run stopped
        self.logger.info(f"Checkpoint loaded from {checkpoint_path}")

    def train(self):
        self.logger.info("Starting diffusion training...")
        tc = self.config.get("training", {})
        for epoch in range(self.current_epoch, tc.get("diffusion_epochs", 1)):
            self.current_epoch = epoch
            tm = self.train_epoch(epoch)
            if math.isinf(tm["train_loss"]):
                self.logger.error(f"Training failed at epoch {epoch}, stopping")
                break
            vm = self.validate_epoch(epoch)
            if self.vae_decoder is not None and epoch % tc.get("sample_every", 10) == 0:
                self.generate_samples(epoch)
            is_best = vm["val_loss"] < self.best_val_loss
            if is_best:
                self.best_val_loss = vm["val_loss"]
            if epoch % tc.get("save_every", 10) == 0 or is_best:
                self.save_checkpoint(epoch, is_best)
            self.logger.info(f"Epoch {epoch}: train_loss={tm['train_loss']:.4f}, val_loss={vm['val_loss']:.4f}")
        self.logger.info("diffusion training completed!")
        if self.writer is not None:
            self.writer.close()


ImprovedDiffusionTrainer = DiffusionTrainer
