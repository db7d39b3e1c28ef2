"""Data-parallel equivalence check (run under torchrun, N >= 2): N ranks training on their shards of a global batch with the
overlapped bucketed all-reduce must follow the same trajectory as ONE process training on the whole batch.
fp32 parity mode, dropout off, fixed timesteps / noise; step 1 is GradSync's calibration pass, steps 2.. use the overlapped
schedule.  Prints the relative difference of the parameter update after `steps` steps (rank 0)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
from pokemon_sprite_generator_b200.unet import UNet

import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("PSG_DUMP_AFTER", "240")), exit=True)      # a hang prints where, and ends the run
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_MAX_CTAS", "16")
dist.init_process_group("nccl", device_id=dev)
per, steps = 2, 4
dtype = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "fp32") else torch.bfloat16


def make():
    torch.manual_seed(0)
    u = UNet(num_heads=4, compute_dtype=dtype)
    with torch.no_grad():       # O(1)-gain re-init of the 0.02-gain layers: gradients well above Adam's eps
        for _, p in u.named_parameters():
            if p.dim() >= 2 and float(p.std()) < 5e-3:
                p.mul_(20.0)
    return u.to(dev).eval()


g = torch.Generator().manual_seed(7)
B = per * world
latent = torch.randn(B, 8, 27, 27, generator=g).clamp_(-3, 3).to(dev)
text = torch.randn(B, 32, 256, generator=g).to(dev)
ts = [torch.randint(0, 1000, (B,), generator=g).to(dev) for _ in range(steps)]
noises = [torch.randn(B, 8, 27, 27, generator=g).to(dev) for _ in range(steps)]
sl = slice(rank * per, (rank + 1) * per)

unet = make()
p0 = {k: v.detach().clone() for k, v in unet.named_parameters()}
step = TrainStep(unet, NoiseScheduler().to(dev), FusedAdamW(unet, max_grad_norm=0.7))
assert step.world == world and step.grad_sync is not None
losses = []
for t, n in zip(ts, noises):
    loss = step(latent[sl], text[sl], timesteps=t[sl], noise=n[sl])
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)        # (the shards are equal: mean of the rank means = global mean)
    losses.append(loss.item() / world)
torch.cuda.synchronize()
stats = dict(step.grad_sync.stats)
# (constructed on every rank: TrainStep's constructor broadcasts rank 0's replica, a collective; only rank 0 then trains it)
ref = make()
rstep = TrainStep(ref, NoiseScheduler().to(dev), FusedAdamW(ref, max_grad_norm=0.7))
rstep.world, rstep.grad_sync = 1, None              # one process, whole batch
if rank == 0:
    rl = [rstep(latent, text, timesteps=t, noise=n).item() for t, n in zip(ts, noises)]
    num = den = 0.0
    for (k, a), (_, b) in zip(unet.named_parameters(), ref.named_parameters()):
        da, db = (a.detach() - p0[k]).double(), (b.detach() - p0[k]).double()
        num += float((da - db).pow(2).sum())
        den += float(db.pow(2).sum())
    print(f"dp_equivalence world={world} dtype={str(dtype).split('.')[-1]} per_rank_batch={per} steps={steps}: "
          f"losses dp={['%.6f' % v for v in losses]} single={['%.6f' % v for v in rl]} | relative update difference "
          f"{(num / den) ** 0.5:.3e} | gradsync {stats}", flush=True)
dist.barrier()
dist.destroy_process_group()
