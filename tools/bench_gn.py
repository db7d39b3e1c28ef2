"""GroupNorm(+SiLU) kernel micro-benchmark: achieved GB/s per U-Net shape, cluster-split vs slab single-pass kernels (algorithmic bytes: fwd 2N, bwd 3N)."""
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


lib = K.L.load()
import ctypes
import os
lib.psg_groupnorm_stream_tune.restype = ctypes.c_longlong
if os.environ.get("PSG_GNS_TUNE"):          # streaming backward: "group bytes,rows per chunk,min tensor bytes"
    for i, v in enumerate(os.environ["PSG_GNS_TUNE"].split(",")):
        lib.psg_groupnorm_stream_tune(i, ctypes.c_longlong(int(v)))
ACC = os.environ.get("PSG_GN_BENCH_ACC") == "1"          # backward accumulates into dx (4 N bytes; the GB/s printed still count 3 N)
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""          # "cluster" / "stream": time only that family (for ncu); "fast": both, no slab
for hw, c in SHAPES:
    x = torch.randn(B * hw, c, device=dev).bfloat16()
    dy = torch.randn(B * hw, c, device=dev).bfloat16()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    nbytes = x.numel() * 2
    res = {}
    for mode, name in ((3, "cluster"), (2, "stream"), (1, "slab")):
        if ONLY and name != ONLY and not (ONLY == "fast" and name != "slab"):
            continue
        lib.psg_groupnorm_fused_mode(mode)
        t_f = timeit(lambda: K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True))
        t_b = timeit(lambda: K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, ACC))
        res[name] = (t_f, t_b)
    lib.psg_groupnorm_fused_mode(0)
    print(f"HW={hw:4d} C={c:5d} {nbytes / 1e6:7.1f} MB | " + " | ".join(
        f"{name} fwd {t_f * 1e3:7.1f} us {2 * nbytes / t_f / 1e6:6.0f} GB/s bwd {t_b * 1e3:7.1f} us {3 * nbytes / t_b / 1e6:6.0f} GB/s"
        for name, (t_f, t_b) in res.items()), flush=True)
