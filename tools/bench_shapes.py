"""tcgen05 engine on the exact GEMM shapes of one U-Net training step (CUDA events, L2 flushed between runs)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import gemm as G

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, flops, name, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:64s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)


def tn(M, N, K, bn=0, mt=0):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * M * N * K, f"TN     M={M} N={N} K={K} bn={bn} mt={mt}")


def nt(M, N, K, bn=0, mt=0):
    a = torch.randn(K, M, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev)
    timeit(lambda: G.run_gemm(G.mnmajor(a), G.mnmajor(b), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * M * N * K, f"NT     M={M} N={N} K={K} bn={bn} mt={mt}")


def conv(B, H, cin, cout, bn=0, mt=0):
    x = torch.randn(B, H, H, cin, device=dev).bfloat16(); w = torch.randn(cout, 9 * cin, device=dev).bfloat16()
    out = torch.empty(B * H * H, cout, device=dev, dtype=torch.bfloat16)
    timeit(lambda: G.run_gemm(G.im2col(x, 3, 1, 1), G.kmajor(w), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * B * H * H * cout * 9 * cin,
           f"fprop  B={B} H={H} {cin}->{cout} bn={bn} mt={mt}")


def wgrad(B, H, cin, cout, bn=0, mt=0):
    x = torch.randn(B, H, H, cin, device=dev).bfloat16(); dy = torch.randn(B * H * H, cout, device=dev).bfloat16()
    out = torch.empty(cout, 9 * cin, device=dev)
    timeit(lambda: G.run_gemm(G.mnmajor(dy), G.im2col_t(x, 3, 1, 1), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt),
           2.0 * B * H * H * cout * 9 * cin, f"wgrad  B={B} H={H} {cin}->{cout} bn={bn} mt={mt}")


def epi_sweep(M=50176, N=1280, K=640):
    from pokemon_sprite_generator_b200 import _lib as L
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    variants = {
        "plain": dict(),
        "bias": dict(bias=bias),
        "bias+gelu": dict(bias=bias, act=L.ACT_GELU),
        "bias+aux_out": dict(bias=bias, aux_out=pre),
        "bias+gelu+aux_out": dict(bias=bias, act=L.ACT_GELU, aux_out=pre),
        "bias+gelu+aux_out+drop": dict(bias=bias, act=L.ACT_GELU, aux_out=pre, drop_seed=123, drop_p=0.05),
        "bias+drop": dict(bias=bias, drop_seed=123, drop_p=0.05),
        "bias+residual": dict(bias=bias, residual=res, alpha=0.7),
        "aux_in(gelu')+drop": dict(aux_in=pre, aux_act=L.ACT_GELU, drop_seed=123, drop_p=0.05),
        "aux_in(gelu')": dict(aux_in=pre, aux_act=L.ACT_GELU),
    }
    from pokemon_sprite_generator_b200 import _lib as L2
    L2.load().psg_umma_debug(1)
    timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine="umma"), 2.0 * M * N * K, f"EPI {M}x{N}x{K} (no epilogue: mainloop alone)")
    timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine="umma", m_tiles=2), 2.0 * M * N * K, f"EPI {M}x{N}x{K} (no epilogue, mt=2)")
    L2.load().psg_umma_debug(0)
    for name, kw in variants.items():
        timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out, **kw), engine="umma"), 2.0 * M * N * K, f"EPI {M}x{N}x{K} {name}")


import os
if os.environ.get("PSG_PAIRS") is not None:
    from pokemon_sprite_generator_b200 import _lib as _L
    _L.load().psg_umma_pairs(int(os.environ["PSG_PAIRS"]))
B = 256
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "wgrad"):
    for mt in (1, 2):
        for bn in (128, 256):
            wgrad(B, 27, 320, 320, bn, mt); wgrad(B, 14, 640, 640, bn, mt); wgrad(B, 7, 1280, 1280, bn, mt); wgrad(B, 4, 1280, 1280, bn, mt)
if which in ("all", "fprop"):
    for mt in (1, 2):
        conv(B, 27, 320, 320, 160, mt); conv(B, 27, 320, 320, 256, mt); conv(B, 14, 640, 640, 256, mt); conv(B, 7, 1280, 1280, 256, mt)
        conv(B, 4, 1280, 1280, 256, mt)
if which in ("all", "tn"):
    for mt in (1, 2):
        tn(50176, 1280, 640, 256, mt); tn(50176, 640, 1280, 256, mt); tn(50176, 640, 640, 256, mt); tn(50176, 640, 640, 128, mt)
        tn(12544, 1280, 1280, 256, mt); tn(12544, 2560, 1280, 256, mt); tn(4096, 1280, 1280, 256, mt); tn(4096, 1280, 1280, 128, mt)
        tn(8192, 1280, 256, 256, mt)
if which in ("all", "nt"):
    for mt in (1, 2):
        nt(640, 640, 50176, 256, mt); nt(1280, 1280, 12544, 256, mt); nt(1280, 1280, 4096, 256, mt); nt(1280, 256, 8192, 256, mt)
        nt(640, 1280, 50176, 256, mt); nt(1280, 1280, 4096, 128, mt)
if which in ("all", "epi"):
    epi_sweep()
    epi_sweep(50176, 640, 640)
