"""Per-kernel-family time breakdown of one training step, measured with CUDA events around every C-ABI call."""
import collections
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import _lib as L
from pokemon_sprite_generator_b200 import gemm as G
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
for _ in range(3):
    step(lat, txt)
torch.cuda.synchronize()
import os
NCU = os.environ.get("PSG_NCU") == "1"      # under ncu: --profile-from-start off, one step between start/stop
if not NCU:
    unet.engine().weight_stream_enabled = False     # per-call events need serialised kernels (see engine._weight_stream)
    L.CALL_PROFILE, G.PROFILE = [], []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if NCU:
    torch.cuda.profiler.start()
e0.record()
step(lat, txt)
e1.record()
torch.cuda.synchronize()
if NCU:
    torch.cuda.profiler.stop()
    print(f"step under profiler: {e0.elapsed_time(e1):.1f} ms")
    sys.exit(0)
calls, gemms = L.CALL_PROFILE, G.PROFILE
L.CALL_PROFILE, G.PROFILE = None, None
total = e0.elapsed_time(e1)
agg = collections.defaultdict(lambda: [0.0, 0])
for name, a, b in calls:
    agg[name][0] += a.elapsed_time(b)
    agg[name][1] += 1
MODE = {0: "K", 1: "MN", 2: "im2col", 3: "im2colT", 4: "dgrad", 5: "convwT"}
gagg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for a, b, fl, eng, (am, bm, M, N, Kd, sp), _sig in gemms:
    key = f"gemm:{eng}:{MODE[am]}x{MODE[bm]}"
    gagg[key][0] += a.elapsed_time(b)
    gagg[key][1] += 1
    gagg[key][2] += fl
print(f"step total {total:.2f} ms at batch {B}")
rows = [(v[0], k, v[1], None) for k, v in agg.items()] + [(v[0], k, v[1], v[2]) for k, v in gagg.items()]
for ms, k, n, fl in sorted(rows, reverse=True):
    extra = f"  {fl / ms / 1e9:8.1f} TFLOP/s" if fl else ""
    print(f"{ms:9.3f} ms  {100 * ms / total:5.1f}%  n={n:5d}  {k}{extra}")
print(f"sum of measured calls: {sum(r[0] for r in rows):.2f} ms")
if "--shapes" in sys.argv:
    per = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for a, b, fl, eng, key, sig in gemms:
        kk = (eng, MODE[key[0]] + "x" + MODE[key[1]]) + key[2:5] + (sig,)
        per[kk][0] += a.elapsed_time(b)
        per[kk][1] += 1
        per[kk][2] += fl
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:70]:
        print(f"{v[0]:8.3f} ms n={v[1]:3d} {v[2] / v[0] / 1e9:8.1f} TF/s  {k}")
