"""One fused-attention forward + backward for ncu (L=196 self-attention shape)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import ops as K
B, H, lq, lk, c = 256, 4, 196, 196, 640
hd = c // H
dev = torch.device("cuda:0")
q = torch.randn(B * lq, 3 * c, device=dev).bfloat16()
kv = torch.randn(B * lk, 2 * c, device=dev).bfloat16()
o = torch.empty(B * lq, c, device=dev, dtype=torch.bfloat16)
do = torch.randn(B * lq, c, device=dev).bfloat16()
lse = torch.empty(B, H, lq, device=dev)
dq = torch.empty_like(q); dkv = torch.empty_like(kv)
def run():
    K.attn_fused_fwd(q[:, :c], kv[:, :c], kv[:, c:], o, lse, B, H, lq, lk, hd, 7, 0.05)
    K.attn_fused_bwd(q[:, :c], kv[:, :c], kv[:, c:], o, do, lse, dq[:, :c], dkv[:, :c], dkv[:, c:], B, H, lq, lk, hd, 7, 0.05)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
