"""GroupNorm(+SiLU) kernel micro-benchmark: achieved GB/s per U-Net shape, single-pass vs two-pass kernels."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import ops as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SHAPES = [(729, 320), (729, 640), (196, 640), (196, 1280), (49, 1280), (49, 2560), (16, 1280), (16, 2560)]
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for hw, c in SHAPES:
    x = torch.randn(B * hw, c, device=dev).bfloat16()
    dy = torch.randn(B * hw, c, device=dev).bfloat16()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    nbytes = x.numel() * 2
    t_f = timeit(lambda: K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True))
    t_b = timeit(lambda: K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False))
    t_f2 = timeit(lambda: K.groupnorm_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True))
    t_b2 = timeit(lambda: K.groupnorm_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False))
    print(f"HW={hw:4d} C={c:5d} {nbytes / 1e6:7.1f} MB | fused fwd {t_f * 1e3:7.1f} us {2 * nbytes / t_f / 1e6:7.0f} GB/s | "
          f"fused bwd {t_b * 1e3:7.1f} us {3 * nbytes / t_b / 1e6:7.0f} GB/s | 2-pass fwd {t_f2 * 1e3:7.1f} us {2 * nbytes / t_f2 / 1e6:7.0f} GB/s | "
          f"2-pass bwd {t_b2 * 1e3:7.1f} us {3 * nbytes / t_b2 / 1e6:7.0f} GB/s")
