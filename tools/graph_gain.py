"""How much of the eager U-Net forward is launch gaps: eager forward vs CUDA-graph replay of the same forward (eval, no grad)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200.sampler import _GraphedUNet
from pokemon_sprite_generator_b200.unet import UNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet(num_heads=4).to(dev).eval()
x = torch.randn(B, 8, 27, 27, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
text = torch.randn(B, 32, 256, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    eager = timeit(lambda: unet(x, t, text))
    g = _GraphedUNet(unet, x, t, text)
    graph = timeit(lambda: g.graph.replay())
print(f"B={B}: eager forward {eager:.3f} ms, graph replay {graph:.3f} ms, gap {eager - graph:.3f} ms ({100 * (eager - graph) / eager:.1f}%)")
