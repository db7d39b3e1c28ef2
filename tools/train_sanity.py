"""End-to-end training sanity on one GPU: the fused bf16 train step (dropout on, AdamW + clip 0.7 + OneCycleLR) must drive the
SmoothL1 noise-prediction loss down on a small fixed set of synthetic latents.  python tools/train_sanity.py [steps] [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
from pokemon_sprite_generator_b200.unet import UNet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet(num_heads=4).to(dev).train()
opt = FusedAdamW(unet, lr=1e-4, weight_decay=1e-4, max_grad_norm=0.7)
sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=steps, pct_start=0.1, anneal_strategy="cos")
step = TrainStep(unet, NoiseScheduler().to(dev), opt, sched)
g = torch.Generator(device="cuda").manual_seed(1)
# structured "latents": a few smooth prototypes plus small noise, so that there is something to learn
protos = torch.nn.functional.interpolate(torch.randn(4, 8, 5, 5, device=dev, generator=g), size=27, mode="bilinear")
lat = (protos[torch.randint(0, 4, (4 * B,), device=dev, generator=g)] + 0.1 * torch.randn(4 * B, 8, 27, 27, device=dev, generator=g)).clamp_(-3, 3)
txt = torch.randn(4 * B, 32, 256, device=dev, generator=g)
window, hist = [], []
for i in range(steps):
    j = (i % 4) * B
    loss = step(lat[j:j + B], txt[j:j + B])
    window.append(loss)
    if (i + 1) % (steps // 10) == 0:
        m = torch.stack(window).mean().item()
        hist.append(m)
        window = []
        print(f"steps {i + 2 - steps // 10:4d}-{i + 1:4d}: mean loss {m:.4f}  lr {opt.param_groups[0]['lr']:.2e}  grad norm {opt.clip_state[0].item():.3f}", flush=True)
assert all(map(lambda v: v == v, hist)), "non-finite loss"
print(f"train_sanity: first window {hist[0]:.4f} -> last window {hist[-1]:.4f} ({'decreasing' if hist[-1] < hist[0] - 0.05 and all(b <= a + 2e-3 for a, b in zip(hist, hist[1:])) else 'NOT decreasing'})")
