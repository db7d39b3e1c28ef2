"""Fused attention micro-benchmark on the U-Net's shapes: us per call and effective TFLOP/s (4*L_q*L_k*hd per head fwd)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import ops as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (lq, lk, c) in [(196, 196, 640), (196, 32, 640), (49, 49, 1280), (49, 32, 1280), (16, 16, 1280), (16, 32, 1280)]:
    hd = c // H
    q = torch.randn(B * lq, 3 * c, device=dev).bfloat16()
    kv = torch.randn(B * lk, 2 * c, device=dev).bfloat16()
    o = torch.empty(B * lq, c, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B * lq, c, device=dev).bfloat16()
    lse = torch.empty(B, H, lq, device=dev)
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    fl = 4.0 * B * H * lq * lk * hd
    for p in (0.0, 0.05):
        tf = timeit(lambda: K.attn_fused_fwd(q[:, :c], kv[:, :c], kv[:, c:], o, lse, B, H, lq, lk, hd, 7, p))
        tb = timeit(lambda: K.attn_fused_bwd(q[:, :c], kv[:, :c], kv[:, c:], o, do, lse, dq[:, :c], dkv[:, :c], dkv[:, c:], B, H, lq, lk, hd, 7, p))
        print(f"Lq={lq:4d} Lk={lk:4d} hd={hd:4d} p={p:.2f} | fwd {tf * 1e3:8.1f} us {fl / tf / 1e9:7.1f} TF/s | bwd {tb * 1e3:8.1f} us {2.5 * fl / tb / 1e9:7.1f} TF/s", flush=True)
