"""One GroupNorm(+SiLU) forward + backward per kernel family for ncu: python tools/ncu_gn.py [HW C [B]]  (default 729 320 256).
Launch order inside the profiled range: cluster fwd, cluster bwd (+param fold), slab fwd, slab bwd (+param fold)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pokemon_sprite_generator_b200 import ops as K

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 729
c = int(sys.argv[2]) if len(sys.argv) > 2 else 320
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda:0")
lib = K.L.load()
x = torch.randn(B * hw, c, device=dev).bfloat16()
dy = torch.randn(B * hw, c, device=dev).bfloat16()
y, dx = torch.empty_like(x), torch.empty_like(x)
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
stats = torch.empty(B, 32, 2, device=dev)
dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)


import os
if os.environ.get("PSG_GN_TUNE"):
    for i, v in enumerate(os.environ["PSG_GN_TUNE"].split(",")):
        lib.psg_groupnorm_cluster_tune(i, int(v))
MODES = (0,) if os.environ.get("PSG_GN_ONLY_CLUSTER") else (0, 1)


def run():
    for mode in MODES:
        lib.psg_groupnorm_fused_mode(mode)
        K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True)
        K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False)
    lib.psg_groupnorm_fused_mode(0)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
