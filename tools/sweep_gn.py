"""Sweep of the cluster-split GroupNorm tunables (fwd threads, bwd threads, fwd slab bytes per CTA, largest cluster, vectors per
unit row, bwd slab bytes per CTA, L2 prefetch one residency ahead) over the U-Net's GroupNorm shapes."""
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


D = (160, 160, 32768, 8, 0, 0)                      # the shipped plan (entries 0-5); entry 6 = prefetch (0 off, 1 auto, n resident CTAs)
CONFIGS = [D + (0,), D + (1,), D + (296,), D + (1184,),
           (160, 160, 32768, 8, 0, 16384, 0), (160, 160, 32768, 8, 0, 16384, 1),
           (160, 160, 16384, 8, 0, 16384, 1), (160, 160, 16384, 8, 0, 8192, 1),
           (160, 160, 32768, 8, 10, 16384, 1), (160, 160, 32768, 8, 20, 16384, 1), (160, 160, 32768, 8, 20, 32768, 1),
           (320, 320, 32768, 8, 0, 0, 1), (320, 320, 32768, 8, 0, 16384, 1), (320, 320, 65536, 8, 0, 32768, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "pf":          # prefetch distance only (resident CTAs assumed)
    CONFIGS = [D + (v,) for v in (0, 74, 148, 222, 296, 444, 592, 888)]
for hw, c in [(729, 320), (729, 640), (196, 640), (196, 1280), (49, 1280), (49, 2560), (16, 1280), (16, 2560)]:
    x = torch.randn(B * hw, c, device=dev).bfloat16()
    dy = torch.randn(B * hw, c, device=dev).bfloat16()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    nbytes = x.numel() * 2
    best_f = best_b = (1e9, None)
    for cfg in CONFIGS:
        for i, v in enumerate(cfg):
            lib.psg_groupnorm_cluster_tune(i, v)
        try:
            t_f = timeit(lambda: K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True))
            t_b = timeit(lambda: K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False))
        except Exception as e:  # noqa: BLE001
            print(f"HW={hw} C={c} cfg={cfg}: {e}", flush=True)
            continue
        best_f = min(best_f, (t_f, cfg))
        best_b = min(best_b, (t_b, cfg))
        print(f"HW={hw:4d} C={c:5d} cfg={cfg} | fwd {t_f * 1e3:7.1f} us {2 * nbytes / t_f / 1e6:6.0f} GB/s | bwd {t_b * 1e3:7.1f} us "
              f"{3 * nbytes / t_b / 1e6:6.0f} GB/s", flush=True)
    print(f"HW={hw:4d} C={c:5d} best fwd {best_f[0] * 1e3:7.1f} us {2 * nbytes / best_f[0] / 1e6:6.0f} GB/s {best_f[1]} | best bwd {best_b[0] * 1e3:7.1f} us "
          f"{3 * nbytes / best_b[0] / 1e6:6.0f} GB/s {best_b[1]}", flush=True)
