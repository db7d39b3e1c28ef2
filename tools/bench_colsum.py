"""Bias-gradient column sums on the U-Net's shapes: the dedicated kernel (psg_colsum) vs the tcgen05 engine (dY^T x ones: an NT GEMM
whose A operand is dY, so the sums come out of TMA-fed tensor-core tiles) -- us per call and GB/s of dY read."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import gemm as G
from pokemon_sprite_generator_b200 import ops as K

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (M, N) in [(50176, 640), (50176, 1280), (50176, 1920), (50176, 2560), (12544, 1280), (12544, 2560), (12544, 3840), (12544, 5120),
               (4096, 1280), (4096, 2560), (4096, 5120), (186624, 320), (8192, 1280)]:
    dy = torch.randn(M, N, device=dev).bfloat16()
    gb = torch.empty(N, device=dev)
    ones = torch.zeros(M, 64, device=dev, dtype=torch.bfloat16); ones[:, 0] = 1
    scratch = torch.empty(N, 64, device=dev)
    t0 = timeit(lambda: K.colsum(dy, 1, None, gb))
    ref = gb.clone()
    t1 = timeit(lambda: G.run_gemm(G.mnmajor(dy), G.mnmajor(ones), G.Epilogue(out=scratch), engine="umma"))
    err = (scratch[:, 0] - ref).abs().max().item() / ref.abs().max().item()
    mb = M * N * 2 / 1e6
    print(f"M={M:6d} N={N:5d} {mb:7.1f} MB | colsum {t0 * 1e3:7.1f} us {mb / t0 / 1e3:6.2f} TB/s | gemm {t1 * 1e3:7.1f} us {mb / t1 / 1e3:6.2f} TB/s | rel diff {err:.1e}", flush=True)
