"""Ad-hoc diagnostic: per-parameter gradient error of the CUDA U-Net vs the golden reference gradients."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import inputs
from pokemon_sprite_generator_b200.unet import UNet
from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
from pokemon_sprite_generator_b200.losses import SmoothL1Loss

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
case_name = sys.argv[2] if len(sys.argv) > 2 else "init_h8_b2_l32"
gold = torch.load(Path(__file__).resolve().parents[1] / "tests/golden/unet_cases.pt")
case = gold["cases"][case_name]
torch.manual_seed(0)
m = UNet(compute_dtype=torch.float32 if mode == "fp32" else torch.bfloat16)
if case_name.startswith("amp"):
    m.load_state_dict(inputs.amplify_state_dict(m.state_dict()))
m = m.cuda().eval()
m.num_heads = case["heads"]
latent, text, t, noise = [a.cuda() for a in inputs.make_inputs(case["batch"], case["text_len"], case["seed"])]
noisy = NoiseScheduler().add_noise(latent, noise, t)
pred = m(noisy, t, text)
loss = SmoothL1Loss(beta=0.1)(pred, noise)
loss.backward()
rows = []
for k, p in m.named_parameters():
    ref = case["grad_norms"][k]
    rows.append((abs(p.grad.norm().item() - ref) / (ref + 1e-12), k, ref, p.grad.norm().item()))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f"{r[0]:.3e}  {r[1]:60s} ref={r[2]:.4e} mine={r[3]:.4e}")
print("median rel err", sorted(r[0] for r in rows)[len(rows) // 2])

# ---- isolate: recompute conv weight grads with torch (fp64, GPU) from the engine's own (x, dy) pairs ----
if "--isolate" in sys.argv:
    import torch.nn.functional as F
    eng = m.engine()
    rec = {}
    orig = eng._conv_bwd

    def hooked(x, cw, out, rowbias, residual, x_needs_grad, e):
        orig(x, cw, out, rowbias, residual, x_needs_grad, e)
        name = [k for k, mod in m.named_modules() if mod is cw.mod][0]
        if name in ("dec_block3.1.res_block.conv1", "middle_block.res_block.conv1", "enc_block0.0.res_block.conv1"):
            xn = x.nhwc().double().permute(0, 3, 1, 2)
            dyn = out.g().double().view(out.B, out.H, out.W, -1).permute(0, 3, 1, 2)
            ref = torch.nn.grad.conv2d_weight(xn, cw.mod.weight.shape, dyn, stride=cw.stride, padding=cw.pad)
            mine = eng.store.grad_of(cw.mod.weight).double()
            print(f"[isolate] {name}: |mine-ref|/|ref| = {((mine - ref).norm() / ref.norm()).item():.3e}  "
                  f"norm ratio-1 = {(mine.norm() / ref.norm()).item() - 1:.3e}  golden ratio-1 = "
                  f"{ref.norm().item() / case['grad_norms'][name + '.weight'] - 1:.3e}")
    eng._conv_bwd = hooked
    m.zero_grad(set_to_none=True)
    pred = m(noisy, t, text)
    SmoothL1Loss(beta=0.1)(pred, noise).backward()
