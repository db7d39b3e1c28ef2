#!/usr/bin/env python
"""Benchmark of the B200-native latent-diffusion hot path (BASELINE.json metric, config 2 / 3).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1        # CPU arm (the unmodified reference from oracle/_ref)

A "step" is one optimisation step of the stage-2 trainer on synthetic inputs of the reference's shapes
(SURVEY.md 8d): q_sample -> U-Net forward -> SmoothL1 -> backward -> [NCCL all-reduce] -> global-norm clip -> AdamW ->
OneCycleLR, bf16 tensor-core compute with fp32 master weights, batch 256 per GPU (weak scaling), dropout on.
`value` is timed with inputs resident in HBM; `e2e` runs the same step through DiffusionTrainer.train_step with pinned
HOST inputs copied in and the loss read back every step.  Rank 0 prints exactly one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FWD_GFLOP_PER_SAMPLE = 77.46          # SURVEY.md 8d (2*MACs over convs, linears, MHA incl. cores)
TRAIN_GFLOP_PER_SAMPLE = 3 * FWD_GFLOP_PER_SAMPLE
METRIC = "unet_train_latent_samples_per_s"


def _umma_traffic():
    """DRAM read+write bytes per tcgen05 launch (mean over one step's launches), from the committed ncu launch list
    (tools/summarize_launches.py --json over `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); None if absent."""
    p = ROOT / "profiles" / "umma_traffic.json"
    try:
        return json.loads(p.read_text())["dram_bytes_per_launch"]
    except Exception:
        return None


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def _nvml(self):
        """In-process NVML sampling (nvidia_ml_py): no fork of a 40 GB process every 0.2 s next to a timed loop whose host side is
        only ~1.6x ahead of the GPU.  Returns False when NVML is unusable (the nvidia-smi loop below runs instead)."""
        if os.environ.get("PSG_CLOCKS") == "smi":
            return False
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)]
            get_reasons(h)
        except Exception:
            return False
        while not self._stop_evt.is_set():
            try:
                r = get_reasons(h)
                self.rows.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), str(N.nvmlDeviceGetPowerUsage(h) / 1e3)] +
                                 ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop_evt.wait(0.1)
        return True

    def run(self):
        if self._nvml():
            return
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        busy = [s for s in sm if mx and s > 0.3 * mx[0]] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref, shipped by oracle/build_ref.py) -- its own train_epoch, U-Net, scheduler
# ----------------------------------------------------------------------------------------------------------------------
CPU_BATCH = 8       # samples per CPU step: a bounded sample of the batch-256 workload (same graph, same optimiser)


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass


def _reference_trainer_namespace(ref, batch: int, n_batches: int, total_steps: int):
    """`self` for ImprovedDiffusionTrainer.train_epoch driven unbound (the full ctor needs a VAE checkpoint, BERT weights
    and the dataset, none of which exist offline -- SURVEY 8c).  The frozen encoders, which are NOT on the path, are
    identity stubs over pre-made synthetic latents / text embeddings; everything on the path is the reference's own:
    UNet(num_heads=4), NoiseScheduler, SmoothL1Loss(beta=0.1), AdamW(eps=1e-6), OneCycleLR, the 478-.item() norm loop,
    clip_grad_norm_(0.7) -- i.e. improved_diffusion_trainer.py:211-216,277-283,300,313-320,335-445 executed as written."""
    import logging
    import types
    import torch
    torch.manual_seed(0)
    unet = ref.UNet(latent_dim=8, text_dim=256, time_emb_dim=128, num_heads=4)
    opt = torch.optim.AdamW(unet.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-6)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=total_steps, pct_start=0.1, anneal_strategy="cos")
    g = torch.Generator().manual_seed(1234)
    loader = [{"image": torch.randn(batch, 8, 27, 27, generator=g), "full_description": torch.randn(batch, 32, 256, generator=g)}
              for _ in range(n_batches)]
    logger = logging.getLogger("psg_b200.bench.reference")
    logger.setLevel(logging.ERROR)
    fake = types.SimpleNamespace(unet=unet, data_loaders={"train": loader}, device=torch.device("cpu"),
                                 text_encoder=lambda d: d, vae_encoder=lambda im: (im, None, None),
                                 noise_scheduler=ref.NoiseScheduler(), optimizer=opt, scheduler=sched,
                                 criterion=torch.nn.SmoothL1Loss(beta=0.1), max_grad_norm=0.7, global_step=0,
                                 config={"training": {"log_every": 50}}, writer=_NullWriter(), logger=logger)
    fake.check_for_nans = types.MethodType(ref.ImprovedDiffusionTrainer.check_for_nans, fake)
    return fake


def cpu_reference_steps(steps: int, warmup: int, batch: int = CPU_BATCH):
    """Times the reference's train step on the host cores, fp32, `batch` samples per step, all host threads.
    kind "reference": ImprovedDiffusionTrainer.train_epoch itself over `steps` batches (after `warmup` batches);
    kind "port" (only when neither /root/reference nor oracle/_ref exists): the oracle restatement of the same step."""
    import torch
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    if ref_loader.available():
        from oracle import build_ref
        if ref_loader.REFERENCE_ROOT == build_ref.DST and not build_ref.verify():
            raise SystemExit("oracle/_ref does not match its manifest")
        os.environ.setdefault("TQDM_DISABLE", "1")      # (read when tqdm is first imported, i.e. by the reference module)
        ref = ref_loader.load()
        fake = _reference_trainer_namespace(ref, batch, warmup + steps, warmup + steps + 8)
        loader = fake.data_loaders["train"]
        if warmup:
            fake.data_loaders = {"train": loader[:warmup]}
            ref.ImprovedDiffusionTrainer.train_epoch(fake, 0)
        fake.data_loaders = {"train": loader[warmup:]}
        t0 = time.perf_counter()
        out = ref.ImprovedDiffusionTrainer.train_epoch(fake, 1)
        sec = (time.perf_counter() - t0) / steps
        assert fake.global_step == warmup + steps, f"reference train_epoch skipped batches ({fake.global_step} of {warmup + steps})"
        return {"samples_per_s": batch / sec, "ms_per_step": sec * 1e3, "cores": torch.get_num_threads(), "batch": batch, "steps": steps,
                "kind": "reference", "loss": out["train_loss"],
                "what": "unmodified reference: ImprovedDiffusionTrainer.train_epoch + UNet(num_heads=4) + NoiseScheduler + "
                        "AdamW/OneCycleLR from oracle/_ref (byte-identical copy of /root/reference/src), fp32"}
    from oracle import inputs, unet_oracle
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.unet import UNet
    torch.manual_seed(0)
    model = UNet()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.named_parameters()}
    sd = dict(params)
    sd["time_embed.emb_coeff"] = model.time_embed.emb_coeff
    del model
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-6)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=warmup + steps + 8, pct_start=0.1, anneal_strategy="cos")
    crit = torch.nn.SmoothL1Loss(beta=0.1)
    ns = NoiseScheduler()
    latent, text, _, _ = inputs.make_inputs(batch, 32, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        lat = torch.clamp(latent, -3.0, 3.0)
        t = torch.randint(0, 1000, (batch,))
        noise = torch.randn_like(lat)
        noisy = ns.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1) * lat + ns.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1) * noise
        opt.zero_grad()
        pred = unet_oracle.unet_forward(sd, noisy, t, text, num_heads=4)
        loss = crit(pred, noise)
        loss.backward()
        total = 0.0
        for p in params.values():
            total += p.grad.norm(2).item() ** 2
        torch.nn.utils.clip_grad_norm_(list(params.values()), max_norm=0.7)
        opt.step()
        sched.step()
        loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"samples_per_s": batch / sec, "ms_per_step": sec * 1e3, "cores": torch.get_num_threads(), "batch": batch, "steps": steps,
            "kind": "port", "what": "oracle port of the reference train step (oracle/_ref absent), fp32"}


def workload_config(B: int, Lt: int, heads: int, dropout: bool, world: int) -> dict:
    """`config` of both arms (the reference arm times a bounded sample of this workload on the host cores)."""
    return {"workload": f"U-Net diffusion train step (config 2), 27x27x8 latents, batch {B}/GPU, bf16 compute + fp32 master, "
                        f"1000-step cosine schedule, {Lt}x256 text emb, AdamW + clip 0.7 + OneCycleLR, dropout "
                        f"{'on' if dropout else 'off'}, heads {heads}",
            "global_batch": world * B, "parallelism": f"dp{world}",
            "l2": "per-step working set (>10 GB activations + 1.3 GB bf16 weights) far exceeds the 126 MB L2"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(max(1, args.steps), max(0, args.warmup))
    sample = f"{r['steps']} steps of {r['batch']} samples after {max(0, args.warmup)} warm-up ({r['what']})"
    line = {"impl": "reference", "metric": METRIC, "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, args.text_len, args.heads, not args.no_dropout, max(1, args.gpus)),
            "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loss": r.get("loss")}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm, config 4: DDPM sampling (improved_diffusion_trainer.py:508-569), prompt-sharded, no communication
# ----------------------------------------------------------------------------------------------------------------------
def run_sample_mode(args):
    """Every rank runs sampler.ddpm_sample(use_cuda_graph=True) over ITS 128 prompts (1024 / 8; weak scaling in the GPU count)
    for `--sample-steps` reverse steps (1000: fast_sampling=False; anything else: the 20-step fast schedule).  The timed
    region is the whole public call (x_T draw, graph capture excluded by one untimed 20-step call, 1000 x [graph replay +
    noise draw + reverse-step kernel]); rank 0 prints one JSON line with denoise steps/s per GPU and whole-job prompts/s."""
    import torch
    import torch.distributed as dist
    from pokemon_sprite_generator_b200 import _lib as L
    from pokemon_sprite_generator_b200 import parallel, sampler
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.unet import UNet
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().psg_check_device(), "psg_check_device")
    per_gpu = args.prompts // 8
    torch.manual_seed(0)
    unet = UNet(num_heads=args.heads, compute_dtype=torch.bfloat16).to(dev).eval()
    ns = NoiseScheduler().to(dev)
    g = torch.Generator(device="cpu").manual_seed(4321)
    all_text = torch.randn(world * per_gpu, args.text_len, 256, generator=g)
    text = parallel.shard_prompts(all_text).to(dev)          # this rank's slice; nothing is exchanged afterwards
    assert text.shape[0] == per_gpu
    full = args.sample_steps >= 1000
    torch.manual_seed(1000 + rank)
    sampler.ddpm_sample(unet, ns, text, per_gpu, fast_sampling=True, use_cuda_graph=True)      # warm-up (20 steps)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib = L.load()
    lib.psg_launch_count.restype = __import__("ctypes").c_longlong
    lib.psg_launch_count(1)
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x = sampler.ddpm_sample(unet, ns, text, per_gpu, fast_sampling=not full, use_cuda_graph=True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    n_steps = 1000 if full else 20
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    finite = bool(torch.isfinite(x).all())
    if rank == 0:
        steps_per_s = n_steps / (ms / 1e3)
        line = {"metric": "ddpm_denoise_steps_per_s_per_gpu", "value": steps_per_s, "unit": "steps/s/GPU", "n_gpus": world,
                "steps": n_steps, "warmup": 20, "ms_per_step": ms / n_steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"DDPM sampling (config 4), {n_steps} reverse steps, {per_gpu} prompts/GPU ({world * per_gpu} total), "
                                       f"27x27x8 latents, {args.text_len}x256 text emb, heads {args.heads}, no CFG, CUDA-graph U-Net forward",
                           "parallelism": f"prompt-sharded x{world}, no communication"},
                "prompts_per_s": world * per_gpu / (ms / 1e3), "total_s": ms / 1e3, "finite": finite,
                "tflops_per_gpu": per_gpu * FWD_GFLOP_PER_SAMPLE / 1e3 / (ms / n_steps / 1e3),
                "gpu_launches": int(lib.psg_launch_count(0)), "clocks": clk,
                "note": "launch count: kernels inside graph replays are counted once at capture, not per replay"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sprite_mode(args):
    """BASELINE config 5 from token ids on: every rank turns ITS 64 prompts (512 / 8; weak scaling) into sprites -- BERT-mini-shaped
    text encoder (random weights: there is no network for the checkpoint; tokenisation, a CPU string operation, is outside) -> 50
    posterior DDPM steps of the U-Net (CUDA-graph forward) -> VAE decoder to 215 x 215 -- reading host token ids and writing the
    images back to pinned host memory inside the timed region.  Rank 0 prints one JSON line (sprites/s whole job, and the split
    between text encoding, the sampling loop and the decoder)."""
    import torch
    import torch.distributed as dist
    from pokemon_sprite_generator_b200 import _lib as L
    from pokemon_sprite_generator_b200 import parallel, sampler
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler
    from pokemon_sprite_generator_b200.unet import UNet
    from pokemon_sprite_generator_b200.text_encoder import TextEncoder
    from pokemon_sprite_generator_b200.vae import VAEDecoder
    from transformers import BertConfig, BertModel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().psg_check_device(), "psg_check_device")
    per_gpu = args.sprites // 8
    torch.manual_seed(0)
    unet = UNet(num_heads=args.heads, compute_dtype=torch.bfloat16).to(dev).eval()
    dec = VAEDecoder().to(dev).eval()
    bert = BertModel(BertConfig(hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=1024))     # BERT-mini shape
    tenc = TextEncoder(hidden_dim=256, bert=bert).to(dev).eval()
    sched = LinearNoiseScheduler()
    g = torch.Generator(device="cpu").manual_seed(4321)
    all_ids = torch.randint(1000, 30000, (world * per_gpu, args.text_len), generator=g)
    host_ids = parallel.shard_prompts(all_ids).contiguous().pin_memory()
    host_img = torch.empty(per_gpu, 3, 215, 215).pin_memory()
    chunk = args.decode_chunk
    e_txt = torch.cuda.Event(enable_timing=True)

    def run(steps):
        text = tenc.encode_ids(host_ids.to(dev, non_blocking=True))
        e_txt.record()
        lat = sampler.posterior_sample(unet, sched, text, steps, use_cuda_graph=True)
        e_mid = torch.cuda.Event(enable_timing=True)
        e_mid.record()
        for i in range(0, per_gpu, chunk):       # the decoder's 215 x 215 activations: a chunk of prompts at a time
            img = torch.clamp((dec(lat[i:i + chunk], text[i:i + chunk]) + 1.0) / 2.0, 0, 1)
            host_img[i:i + chunk].copy_(img, non_blocking=True)
        return e_mid

    torch.manual_seed(1000 + rank)
    run(2)                                        # warm-up: graph capture, weight packing, workspaces
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib = L.load()
    lib.psg_launch_count.restype = __import__("ctypes").c_longlong
    lib.psg_launch_count(1)
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_mid = run(args.sprite_steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    t = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e_mid), e0.elapsed_time(e_txt)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_sample, ms_text = t.tolist()
    if rank == 0:
        line = {"metric": "text_to_sprite_images_per_s", "value": world * per_gpu / (ms / 1e3), "unit": "sprites/s", "n_gpus": world,
                "steps": 1, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"text-to-sprite (config 5): BERT-mini-shaped text encoder on {args.text_len} synthetic token ids -> "
                                       f"{args.sprite_steps} posterior DDPM steps (CUDA-graph U-Net forward) -> VAE decoder to 215x215, "
                                       f"{per_gpu} prompts/GPU ({world * per_gpu} total), heads {args.heads}, decoder chunk {chunk}, "
                                       "random-init weights",
                           "parallelism": f"prompt-sharded x{world}, no communication"},
                "e2e": {"value": world * per_gpu / (ms / 1e3), "unit": "sprites/s", "h2d_bytes_per_step": host_ids.numel() * 8,
                        "d2h_bytes_per_step": host_img.numel() * 4},
                "text_encoder_ms": ms_text, "sampling_ms": ms_sample - ms_text, "decoder_ms": ms - ms_sample, "finite": bool(torch.isfinite(host_img).all()),
                "image_mean": float(host_img.mean()), "gpu_launches": int(lib.psg_launch_count(0)), "clocks": clk,
                "note": "launch count: kernels inside graph replays are counted once at capture, not per replay"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE config 2: 256)")
    ap.add_argument("--text-len", type=int, default=32)
    ap.add_argument("--heads", type=int, default=4, help="trainer default (improved_diffusion_trainer.py:215)")
    ap.add_argument("--denoise-batch", type=int, default=128)
    ap.add_argument("--denoise-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--sprites", type=int, default=512, help="--mode sprite: global prompt batch at 8 GPUs (64 per GPU, weak scaling)")
    ap.add_argument("--sprite-steps", type=int, default=50, help="--mode sprite: posterior DDPM steps")
    ap.add_argument("--decode-chunk", type=int, default=16, help="--mode sprite: prompts per VAE-decoder call")
    ap.add_argument("--mode", default="train", choices=["train", "sample", "sprite"],
                    help="sample: BASELINE config 4 -- full DDPM sampling, prompts sharded over the GPUs with no communication")
    ap.add_argument("--prompts", type=int, default=1024, help="--mode sample: global prompt batch at 8 GPUs (128 per GPU, weak scaling)")
    ap.add_argument("--sample-steps", type=int, default=1000, help="--mode sample: reverse steps (1000 = fast_sampling=False)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.mode == "sample":
        return run_sample_mode(args)
    if args.mode == "sprite":
        return run_sprite_mode(args)

    import torch
    import torch.distributed as dist
    from pokemon_sprite_generator_b200 import _lib as L
    from pokemon_sprite_generator_b200 import gemm as G
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    from pokemon_sprite_generator_b200.trainer import DiffusionTrainer, FusedAdamW, TrainStep
    from pokemon_sprite_generator_b200.unet import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 must print exactly ONE line on stdout: NCCL writes its debug output (the "NCCL version ..." banner at
        # VERSION / WARN / INFO level) to stdout unless told otherwise
        # (and NCCL honours NCCL_DEBUG_FILE only above VERSION level, so VERSION is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # the gradient all-reduce runs inside backward (parallel.GradSync): keep its CTA count small and known, so the
        # persistent GEMM grids can leave exactly that many SMs free while buckets are in flight
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("PSG_COMM_SMS", "16"))
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().psg_check_device(), "psg_check_device")
    W = max(args.warmup, 3)
    Ksteps = max(args.steps, 1)
    B, Lt = args.batch, args.text_len

    torch.manual_seed(0)
    unet = UNet(num_heads=args.heads, compute_dtype=torch.bfloat16).to(dev)
    unet.train()
    eng = unet.engine()
    eng.dropout_enabled = not args.no_dropout
    ns = NoiseScheduler().to(dev)
    opt = FusedAdamW(unet, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-4, max_grad_norm=0.7)
    total_sched = 4 * (W + Ksteps) + 64
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=total_sched, pct_start=0.1, anneal_strategy="cos")
    step_fn = TrainStep(unet, ns, opt, sched)

    # synthetic inputs; 4 distinct batches, data seed 1234 + rank
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nb = 4
    host_lat = [torch.randn(B, 8, 27, 27, generator=g).clamp_(-3, 3).pin_memory() for _ in range(nb)]
    host_txt = [torch.randn(B, Lt, 256, generator=g).pin_memory() for _ in range(nb)]
    dev_lat = [t.to(dev) for t in host_lat]
    dev_txt = [t.to(dev) for t in host_txt]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step_fn(dev_lat[i % nb], dev_txt[i % nb])
    barrier()

    # ---- timed region: device-resident inputs ----
    lib = L.load()
    lib.psg_launch_count.restype = __import__("ctypes").c_longlong
    lib.psg_launch_count(1)
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(Ksteps):
        loss = step_fn(dev_lat[i % nb], dev_txt[i % nb])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = int(lib.psg_launch_count(0))
    # ---- roofline pass: the same K steps again with per-launch CUDA events on the tensor-core engine.  In the timed region above the
    # weight-gradient GEMMs run on a second stream next to the dX chain (engine._weight_stream), so a launch's event pair there brackets
    # two kernels sharing the SMs; here that stream is off and every launch has the device to itself (its duration is exclusive) ----
    eng.weight_stream_enabled = False
    G.PROFILE = []
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0.record()
    for i in range(Ksteps):
        step_fn(dev_lat[i % nb], dev_txt[i % nb])
    r1.record()
    barrier()
    serial_ms = r0.elapsed_time(r1)
    prof, G.PROFILE = G.PROFILE, None
    from pokemon_sprite_generator_b200 import ops as _K
    _K.check_kernel_timeouts()       # a bounded in-kernel wait that expired voids the measurement: fail loudly
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    ms_per_step = ms / Ksteps
    value = world * B * Ksteps / (ms / 1e3)
    final_loss = loss.item()

    # roofline of the dominant kernel (umma_gemm_kernel: every conv / linear fwd, dgrad, wgrad)
    um = [(s.elapsed_time(e), fl) for (s, e, fl, engn, *_) in prof if engn == "umma"]
    um_ms = sum(x[0] for x in um)
    um_flops = sum(x[1] for x in um)
    peaks = _peaks()
    achieved = um_flops / (um_ms / 1e3) / 1e12 if um_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "umma_gemm_kernel (tcgen05 implicit-GEMM conv + GEMM, fwd/dgrad/wgrad)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                "peak_source": f"{peaks['src']} sustained cuBLAS bf16", "traffic": _umma_traffic(),
                "algorithmic_bytes_note": "tensor-bound kernel: achieved/peak are TFLOP/s; traffic = mean DRAM bytes per launch (ncu)",
                "measured_in": "second pass of the same K steps, weight-gradient stream off (kernels serialised, durations exclusive)",
                "serial_ms_per_step": serial_ms / Ksteps,
                "launches_per_step": len(um) / Ksteps, "share_of_step": um_ms / serial_ms if serial_ms > 0 else None,
                "algorithmic_tflop_per_step": um_flops / Ksteps / 1e12,
                "step_model_tflops": B * TRAIN_GFLOP_PER_SAMPLE / 1e3 / (ms_per_step / 1e3)}

    # ---- HBM-bound kernel families: one extra (untimed) step with CUDA events around every C-ABI call ----
    # algorithmic bytes: GroupNorm fwd 2 N s, bwd 3 N s with N = 6,403,520 activations/sample, s = 2 B (bf16);
    # AdamW (7*4 + 2) B per parameter (fp32 p, g, m, v read, p, m, v written, bf16 shadow written); grad-norm 4 B per parameter
    hbm_kernels = None
    if rank == 0 and world == 1:
        L.CALL_PROFILE = []
        step_fn(dev_lat[0], dev_txt[0])
        torch.cuda.synchronize()
        calls, L.CALL_PROFILE = L.CALL_PROFILE, None
        tot = {}
        for name, s0, s1 in calls:
            tot[name] = tot.get(name, 0.0) + s0.elapsed_time(s1)
        n_act, n_par = 6403520 * B, sum(p.numel() for p in unet.parameters())
        fams = {"groupnorm_fwd": ("psg_groupnorm_fused_fwd", 2 * n_act * 2), "groupnorm_bwd": ("psg_groupnorm_fused_bwd_ws", 3 * n_act * 2),
                "adamw": ("psg_adam_step", 30 * n_par), "grad_sumsq": ("psg_sumsq", 4 * n_par)}
        hbm_kernels = {"peak_gbs": peaks["hbm_gbs"], "peak_source": f"{peaks['src']} copy bandwidth"}
        for fam, (cname, nbytes) in fams.items():
            if tot.get(cname):
                gbs = nbytes / (tot[cname] / 1e3) / 1e9
                hbm_kernels[fam] = {"ms_per_step": round(tot[cname], 3), "algorithmic_gb": round(nbytes / 1e9, 3), "achieved_gbs": round(gbs, 1),
                                    "frac": round(gbs / peaks["hbm_gbs"], 3) if peaks["hbm_gbs"] else None}

        # 43 of the 61 backward norms ADD into an existing gradient (x also feeds a residual connection: every ResBlock norm1 and
        # both norms of an attention block; template argument ACCUM = 1 in profiles/r02_launches_summary_v2.md), so they also read
        # dx: 4 N s instead of 3 N s.  N_accum = 729*1920 + 196*8960 + 49*17920 + 16*21760 = 4,382,080 activations per sample.
        if "groupnorm_bwd" in hbm_kernels:
            gb = hbm_kernels["groupnorm_bwd"]
            nbytes = (3 * n_act + 4382080 * B) * 2
            gbs = nbytes / (gb["ms_per_step"] / 1e3) / 1e9
            gb["with_accumulate_reads"] = {"algorithmic_gb": round(nbytes / 1e9, 3), "achieved_gbs": round(gbs, 1),
                                           "frac": round(gbs / peaks["hbm_gbs"], 3) if peaks["hbm_gbs"] else None,
                                           "note": "43 of 61 norms accumulate into dx (4 N s bytes); the 3 N s figure above is SURVEY 8d's"}

    # q_sample / SmoothL1 / DDPM reverse step at a roofline batch (at batch 256 they are L2-resident 10-20 us
    # launches): batch 8192 => 191 MB per tensor (> the 126 MB L2), each kernel timed alone, 20 launches, CUDA events.
    # algorithmic bytes per sample: q_sample 3 x 5832 x 4, SmoothL1 fwd+bwd 3 x 5832 x 4, DDPM step 4 x 5832 x 4
    if hbm_kernels is not None:
        try:
            from pokemon_sprite_generator_b200.losses import smooth_l1_fwd_bwd
            Bh = 8192
            xs = [torch.randn(Bh, 8, 27, 27, device=dev) for _ in range(3)]
            tt_h = torch.randint(0, 1000, (Bh,), device=dev)
            per = 5832 * 4
            small = {"q_sample": (lambda: ns.add_noise(xs[0], xs[1], tt_h, clamp=3.0), 3 * per * Bh),
                     "smooth_l1_fwd_bwd": (lambda: smooth_l1_fwd_bwd(xs[0], xs[1], beta=0.1), 3 * per * Bh),
                     "ddpm_step": (lambda: ns.ddpm_step(xs[0], xs[1], 500, xs[2]), 4 * per * Bh)}
            for fam, (fn, nbytes) in small.items():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                for _ in range(20):
                    fn()
                s1.record()
                torch.cuda.synchronize()
                msk = s0.elapsed_time(s1) / 20
                gbs = nbytes / (msk / 1e3) / 1e9
                hbm_kernels[fam] = {"ms_per_launch": round(msk, 4), "batch": Bh, "algorithmic_gb": round(nbytes / 1e9, 3),
                                    "achieved_gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 3) if peaks["hbm_gbs"] else None,
                                    "note": "includes the torch.empty of the output (public API call)"}
            del xs
        except Exception as ex:
            hbm_kernels["small_kernels_error"] = repr(ex)[:200]

    eng.weight_stream_enabled = True if os.environ.get("PSG_WGRAD_STREAM", "1") != "0" else False
    # ---- end-to-end through the public trainer API: pinned host inputs in, loss out, every step ----
    # built through the constructor a user calls (synthetic pre-encoded batches injected in place of the dataset / frozen
    # encoders, which are out of scope); it then adopts the warmed-up U-Net / optimiser so no second 13 GB model is built
    import tempfile
    cfg = {"experiment_dir": tempfile.mkdtemp(prefix="psg_bench_"), "model": {"latent_dim": 8, "text_embedding_dim": 256, "num_heads": args.heads},
           "training": {"diffusion_epochs": 1}, "unet_optimization": {"learning_rate": 1e-4, "weight_decay": 1e-4, "max_grad_norm": 0.7}}
    trainer = DiffusionTrainer(cfg, None, f"bench_rank{rank}", components={"data_loaders": {"train": [None] * total_sched, "val": [], "test": []},
                                                                          "unet": unet, "optimizer": opt, "lr_scheduler": sched, "train_step": step_fn})
    assert trainer.unet is unet and trainer._step is step_fn
    for i in range(2):
        trainer.train_step(host_lat[i % nb], host_txt[i % nb]).item()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(Ksteps):
        trainer.train_step(host_lat[i % nb], host_txt[i % nb]).item()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ksteps / (t.item() / 1e3)
    e2e = {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": host_lat[0].numel() * 4 + host_txt[0].numel() * 4,
           "d2h_bytes_per_step": 4, "wall_s": time.perf_counter() - t0}

    comm = None
    if getattr(step_fn, "grad_sync", None) is not None:
        gs = step_fn.grad_sync
        comm = {"overlap": "backward", "buckets": len(gs.bounds), "head_bucket_mb": round((gs.bounds[0][1] - gs.bounds[0][0]) * 4 / 2**20, 1),
                "bucket_mb": round(gs.bucket_elems * 4 / 2**20), "reserve_sms": gs.reserve_sms, "window_entries": gs.window_entries,
                "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), **gs.stats}
    elif world > 1:
        comm = {"overlap": "none", "buckets": step_fn.buckets, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")}

    # ---- DDPM denoise steps/s per GPU (second half of the BASELINE metric), eval mode, prompt-sharded ----
    denoise = None
    try:
        from pokemon_sprite_generator_b200.sampler import _GraphedUNet
        unet.eval()
        Bd = args.denoise_batch
        x = torch.randn(Bd, 8, 27, 27, device=dev)
        te = torch.randn(Bd, Lt, 256, device=dev)
        gun = _GraphedUNet(unet, x, torch.zeros(Bd, dtype=torch.long, device=dev), te)
        tt = 999
        for _ in range(2):
            x = ns.ddpm_step(x, gun(x, tt), tt, torch.randn_like(x))
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.denoise_steps):
            tt = 999 - i
            x = ns.ddpm_step(x, gun(x, tt), tt, torch.randn_like(x))
        e1.record()
        torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / args.denoise_steps
        denoise = {"steps_per_s_per_gpu": 1e3 / dms, "ms_per_step": dms, "batch": Bd, "cuda_graph": True,
                   "tflops": Bd * FWD_GFLOP_PER_SAMPLE / 1e3 / (dms / 1e3)}
        unet.train()
    except Exception as ex:  # the headline metric must still print
        denoise = {"error": repr(ex)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_steps(3, 1)
        cpu_baseline = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"],
                        "sample": f"3 steps of {r['batch']} samples after 1 warm-up ({r['what']})"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": Ksteps, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": workload_config(B, Lt, args.heads, not args.no_dropout, world),
                "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "denoise": denoise, "loss": final_loss, "comm": comm, "hbm_kernels": hbm_kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
