"""Per-step device time and cudaMalloc count of the first 30 train steps on a cold process (is the warm-up long enough?)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
from pokemon_sprite_generator_b200.trainer import FusedAdamW, TrainStep
from pokemon_sprite_generator_b200.unet import UNet

B = 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet(num_heads=4).to(dev).train()
ns = NoiseScheduler().to(dev)
opt = FusedAdamW(unet, max_grad_norm=0.7)
step = TrainStep(unet, ns, opt)
lat = torch.randn(B, 8, 27, 27, device=dev).clamp_(-3, 3)
txt = torch.randn(B, 32, 256, device=dev)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
rows = []
torch.cuda.synchronize()
for i in range(30):
    evs[i].record()
    t0 = time.perf_counter()
    step(lat, txt)
    host = time.perf_counter() - t0
    st = torch.cuda.memory_stats()
    rows.append((host, st.get("num_device_alloc", 0), st.get("num_device_free", 0), st.get("reserved_bytes.all.current", 0)))
evs[30].record()
torch.cuda.synchronize()
for i, (host, na, nf, res) in enumerate(rows):
    print(f"step {i:2d}: device {evs[i].elapsed_time(evs[i + 1]):8.2f} ms  host enqueue {host * 1e3:8.2f} ms  cudaMalloc {na:5d} cudaFree {nf:4d} reserved {res / 2**30:7.2f} GiB")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout)
