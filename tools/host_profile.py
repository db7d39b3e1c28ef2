"""cProfile of the host side of the train step (how far is the Python enqueue ahead of the GPU?): top functions by own / cumulative time."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
from pokemon_sprite_generator_b200.unet import UNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet(num_heads=4).to(dev).train()
ns = NoiseScheduler().to(dev)
opt = FusedAdamW(unet, max_grad_norm=0.7)
step = TrainStep(unet, ns, opt)
lat = torch.randn(B, 8, 27, 27, device=dev).clamp_(-3, 3)
txt = torch.randn(B, 32, 256, device=dev)
for _ in range(8):
    step(lat, txt)
torch.cuda.synchronize()
# un-profiled host time per step with an idle GPU queue (sync before every step)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(lat, txt)
    ts.append(time.perf_counter() - t0)
print("host enqueue per step (GPU queue empty at start):", " ".join(f"{t * 1e3:.1f}" for t in ts), "ms")
pr = cProfile.Profile()
N = 5
torch.cuda.synchronize()
pr.enable()
for _ in range(N):
    step(lat, txt)
    torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(35)
