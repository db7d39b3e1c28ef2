"""One tcgen05 GEMM launch for ncu: python tools/ncu_one_gemm.py <variant> [M N K]  (variant: plain | ffn | res | wgrad | conv)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import _lib as L
from pokemon_sprite_generator_b200 import gemm as G

variant = sys.argv[1] if len(sys.argv) > 1 else "plain"
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (50176, 1280, 640)
dev = torch.device("cuda:0")
if variant in ("plain", "ffn", "res"):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    kw = {"plain": dict(), "ffn": dict(bias=bias, act=L.ACT_GELU, aux_out=pre, aux_act=L.ACT_GELU, drop_seed=1, drop_p=0.05),
          "res": dict(bias=bias, residual=res, alpha=0.7)}[variant]
    fn = lambda: G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out, **kw), engine="umma")
elif variant == "wgrad":
    Bn, H, cin, cout = 256, 14, 640, 640
    x = torch.randn(Bn, H, H, cin, device=dev).bfloat16(); dy = torch.randn(Bn * H * H, cout, device=dev).bfloat16()
    out = torch.empty(cout, 9 * cin, device=dev)
    import os
    L.load().psg_umma_pairs(int(os.environ.get("PSG_PAIRS", "1")))         # tile-shape overrides for A/B captures
    bn, mt = int(os.environ.get("PSG_BN", "0")), int(os.environ.get("PSG_MT", "0"))
    fn = lambda: G.run_gemm(G.mnmajor(dy), G.im2col_t(x, 3, 1, 1), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt)
else:
    Bn, H, cin, cout = 256, 14, 640, 640
    x = torch.randn(Bn, H, H, cin, device=dev).bfloat16(); w = torch.randn(cout, 9 * cin, device=dev).bfloat16()
    out = torch.empty(Bn * H * H, cout, device=dev, dtype=torch.bfloat16)
    fn = lambda: G.run_gemm(G.im2col(x, 3, 1, 1), G.kmajor(w), G.Epilogue(out=out), engine="umma")
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", variant)
