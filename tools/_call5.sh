mkdir -p gpurun_out
cat > /tmp/ncu_gns.py <<'P'
import sys
sys.path.insert(0, '.')
import torch
from pokemon_sprite_generator_b200 import ops as K
hw, c, B = 729, 320, 256
dev = torch.device("cuda:0")
lib = K.L.load()
x = torch.randn(B * hw, c, device=dev).bfloat16(); dy = torch.randn(B * hw, c, device=dev).bfloat16()
y, dx = torch.empty_like(x), torch.empty_like(x)
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
stats = torch.empty(B, 32, 2, device=dev); dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
lib.psg_groupnorm_fused_mode(2)
K.groupnorm_fused_fwd(x, y, gamma, beta, stats, B, 32, 1e-5, True)
K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
K.groupnorm_fused_bwd(dy, x, dx, gamma, beta, stats, dg, db, B, 32, True, False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
P
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/gns_729_320 python /tmp/ncu_gns.py > gpurun_out/ncu_gns.log 2>&1; echo "ncu rc=$?"
