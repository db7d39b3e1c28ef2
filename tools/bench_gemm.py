"""Micro-benchmark of the tcgen05 engine on the U-Net's GEMM shapes (CUDA events, L2 flushed between runs)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import gemm as G

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, flops, name, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:58s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s")


def tiled(M, N, K, bn=0, mt=0):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * M * N * K, f"tiled  M={M} N={N} K={K} bn={bn} mt={mt}")


def conv(B, H, cin, cout, bn=0, mt=0):
    x = torch.randn(B, H, H, cin, device=dev).bfloat16(); w = torch.randn(cout, 9 * cin, device=dev).bfloat16()
    out = torch.empty(B * H * H, cout, device=dev, dtype=torch.bfloat16)
    timeit(lambda: G.run_gemm(G.im2col(x, 3, 1, 1), G.kmajor(w), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * B * H * H * cout * 9 * cin,
           f"im2col B={B} H={H} {cin}->{cout} bn={bn} mt={mt}")


def wgrad(B, H, cin, cout, split=1, bn=128, mt=0):
    x = torch.randn(B, H, H, cin, device=dev).bfloat16(); dy = torch.randn(B * H * H, cout, device=dev).bfloat16()
    part = torch.empty(split, cout, 9 * cin, device=dev)
    timeit(lambda: G.run_gemm(G.mnmajor(dy), G.im2col_t(x, 3, 1, 1), G.Epilogue(out=part[0]), engine="umma", block_n=bn, m_tiles=mt),
           2.0 * B * H * H * cout * 9 * cin, f"wgrad  B={B} H={H} {cin}->{cout} split={split} bn={bn} mt={mt}")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
import os
if os.environ.get("PSG_UMMA_PAIRS"):        # 0 never, 1 the planner's rule, 2 always (where the tile shape allows)
    from pokemon_sprite_generator_b200 import _lib as _L
    _L.load().psg_umma_pairs(int(os.environ["PSG_UMMA_PAIRS"]))
    print("psg_umma_pairs =", os.environ["PSG_UMMA_PAIRS"])
B = 256
if which in ("all", "fprop"):
    for mt in (1, 2):
        conv(B, 27, 320, 320, 160, mt); conv(B, 27, 320, 320, 256, mt); conv(B, 27, 320, 1280, 256, mt)
        conv(B, 14, 640, 640, 256, mt); conv(B, 14, 640, 640, 160, mt); conv(B, 7, 1280, 1280, 256, mt); conv(B, 4, 1280, 1280, 256, mt)
        conv(B, 4, 1280, 1280, 128, mt)
        tiled(50176, 1280, 640, 256, mt); tiled(50176, 640, 1280, 256, mt); tiled(50176, 640, 640, 256, mt); tiled(50176, 640, 640, 128, mt)
        tiled(12544, 1280, 1280, 256, mt); tiled(4096, 1280, 1280, 256, mt); tiled(4096, 1280, 1280, 128, mt)
if which in ("all", "wgrad"):
    for mt in (1, 2):
        wgrad(B, 27, 320, 320, 3, 128, mt); wgrad(B, 27, 320, 320, 2, 256, mt); wgrad(B, 14, 640, 640, 1, 256, mt); wgrad(B, 14, 640, 640, 1, 128, mt)
        wgrad(B, 7, 1280, 1280, 1, 256, mt); wgrad(B, 4, 1280, 1280, 1, 256, mt); wgrad(B, 4, 1280, 1280, 2, 256, mt)
if which == "wgrad_sweep":
    # the conv weight-gradient shapes of a batch-256 step: (H, cin, cout); tile shape x pairing sweep
    from pokemon_sprite_generator_b200 import _lib as L
    shapes = [(14, 640, 640), (27, 320, 320), (14, 1280, 640), (27, 640, 320), (7, 1280, 1280), (4, 1280, 1280), (7, 2560, 1280)]
    for (H, cin, cout) in shapes:
        for pairs in (0, 1, 2):
            L.load().psg_umma_pairs(pairs)
            for bn, mt in ((256, 1), (256, 2), (128, 1), (128, 2)) if pairs != 1 else ((0, 0),):
                print(f"pairs={pairs} ", end="")
                wgrad(B, H, cin, cout, 1, bn, mt)
    L.load().psg_umma_pairs(1)
if which == "modes":
    # same L2-resident problem, same tile shape, through the three operand-layout modes: what does an MN-major operand cost?
    from pokemon_sprite_generator_b200 import _lib as L
    for (M, N, K) in ((3072, 3072, 4096), (3072, 3072, 16384), (1280, 11520, 12544)):
        a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
        at, bt = a.t().contiguous(), b.t().contiguous()
        out = torch.empty(M, N, device=dev)
        for pairs in (0, 2):
            L.load().psg_umma_pairs(pairs)
            for bn, mt in ((256, 1), (256, 2)):
                fl = 2.0 * M * N * K
                timeit(lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), fl, f"TN pairs={pairs} M={M} N={N} K={K} bn={bn} mt={mt}")
                timeit(lambda: G.run_gemm(G.kmajor(a), G.mnmajor(bt), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), fl, f"TT pairs={pairs} M={M} N={N} K={K} bn={bn} mt={mt}")
                timeit(lambda: G.run_gemm(G.mnmajor(at), G.mnmajor(bt), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), fl, f"NT pairs={pairs} M={M} N={N} K={K} bn={bn} mt={mt}")
    L.load().psg_umma_pairs(1)
if which == "wgrad_auto":
    # conv weight gradients of a batch-256 step, automatic tile shape
    for (H, cin, cout) in [(14, 640, 640), (27, 320, 320), (14, 1280, 640), (27, 640, 320), (7, 1280, 1280), (4, 1280, 1280), (7, 2560, 1280),
                           (4, 2560, 1280), (14, 320, 640), (7, 640, 1280), (14, 1920, 640), (27, 960, 320)]:
        wgrad(B, H, cin, cout, 1, 0, 0)
if which == "linear_sweep":
    # the Linear / attention-projection forward + dgrad shapes (TN, short K) of a batch-256 step: tile shape x pairing against the
    # planner's own choice (bn = mt = 0), plain bf16 epilogue
    from pokemon_sprite_generator_b200 import _lib as L
    shapes = [(50176, 640, 640), (50176, 1280, 640), (50176, 640, 1280), (50176, 1920, 640), (50176, 640, 1920), (12544, 1280, 1280),
              (12544, 2560, 1280), (12544, 1280, 2560), (12544, 3840, 1280), (4096, 1280, 1280), (8192, 1280, 2560), (8192, 2560, 1280)]
    for (M, N, K) in shapes:
        for pairs in (1, 2):
            L.load().psg_umma_pairs(pairs)
            for bn, mt in ((0, 0),) if pairs == 1 else ((256, 1), (256, 2), (128, 1), (128, 2)):
                print(f"pairs={pairs} ", end="")
                tiled(M, N, K, bn, mt)
        L.load().psg_umma_pairs(0)
        for bn, mt in ((256, 2), (160, 1), (160, 2), (128, 1), (128, 2)):
            if bn == 160 and N % 160:
                continue
            print("pairs=0 ", end="")
            tiled(M, N, K, bn, mt)
    L.load().psg_umma_pairs(1)
if which == "lwgrad_sweep":
    # Linear weight gradients (NT: dW[n_out][k_in] = sum_tokens dY[t, n_out] X[t, k_in]) of a batch-256 step: tile shape x pairing
    from pokemon_sprite_generator_b200 import _lib as L
    shapes = [(640, 640, 50176), (1280, 1280, 12544), (1280, 1280, 4096), (2560, 1280, 8192), (640, 1280, 50176), (1280, 2560, 12544),
              (3840, 1280, 12544), (1920, 640, 50176), (1280, 640, 50176), (2560, 1280, 12544), (2560, 1280, 4096), (1280, 2560, 4096)]
    for (M, N, K) in shapes:
        dy = torch.randn(K, M, device=dev).bfloat16(); x = torch.randn(K, N, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev)
        for pairs, combos in ((1, ((0, 0),)), (0, ((256, 1), (256, 2), (128, 1), (128, 2))), (2, ((256, 1), (256, 2), (128, 1), (128, 2)))):
            L.load().psg_umma_pairs(pairs)
            for bn, mt in combos:
                timeit(lambda: G.run_gemm(G.mnmajor(dy), G.mnmajor(x), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt), 2.0 * M * N * K,
                       f"pairs={pairs} lwgrad M={M} N={N} K={K} bn={bn} mt={mt}")
    L.load().psg_umma_pairs(1)
