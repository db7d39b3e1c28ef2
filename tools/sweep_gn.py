"""Sweep of the cluster-split GroupNorm tunables (fwd threads, bwd threads, slab bytes per CTA, largest cluster, vectors per
unit row) over the U-Net's GroupNorm shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import ops as K

B = 256
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = K.L.load()


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


CONFIGS = [(160, 160, 32768, 8, 0, 0), (160, 160, 32768, 8, 10, 16384), (160, 160, 32768, 8, 10, 32768), (160, 160, 32768, 8, 20, 16384),
           (160, 160, 32768, 8, 20, 32768), (320, 320, 32768, 8, 10, 16384), (320, 320, 32768, 8, 20, 32768), (320, 320, 65536, 8, 20, 65536),
           (160, 160, 65536, 8, 10, 65536), (160, 160, 16384, 8, 10, 8192)]
for hw, c in [(729, 320), (729, 640), (196, 640), (196, 1280), (49, 1280), (49, 2560), (16, 1280), (16, 2560)]:
    x = torch.randn(B * hw, c, device=dev).bfloat16()
    dy = torch.randn(B * hw, c, device=dev).bfloat16()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    nbytes = x.numel() * 2
    for cfg in CONFIGS:
        for i, v in enumerate(cfg):
            lib.psg_groupnorm_cluster_tune(i, v)
        try:
            t_f = timeit(lambda: K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True))
            t_b = timeit(lambda: K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False))
        except Exception as e:  # noqa: BLE001
            print(f"HW={hw} C={c} cfg={cfg}: {e}", flush=True)
            continue
        print(f"HW={hw:4d} C={c:5d} cfg={cfg} | fwd {t_f * 1e3:7.1f} us {2 * nbytes / t_f / 1e6:6.0f} GB/s | bwd {t_b * 1e3:7.1f} us "
              f"{3 * nbytes / t_b / 1e6:6.0f} GB/s", flush=True)
