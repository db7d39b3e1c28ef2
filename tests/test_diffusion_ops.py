"""q_sample / SmoothL1 / DDPM-step kernels vs the golden vectors produced by the real reference
(oracle/make_golden.py) and vs the numpy oracle.  Integer gather + fp32 arithmetic: bit-exact (torch.equal)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "scheduler_tables.npz"


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _stub_unet(x, t, text):
    return x * 0.5 - 0.125


def test_tables_match_reference(cuda_device, gold):
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler, NoiseScheduler
    ns = NoiseScheduler()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(ns, k).numpy(), gold["cos_" + k]), k
    ls = LinearNoiseScheduler()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas",
              "posterior_variance"):
        assert np.array_equal(getattr(ls, k).numpy(), gold["lin_" + k]), k


def test_q_sample_bit_exact(cuda_device, gold):
    from oracle import diffusion_oracle as O
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler, NoiseScheduler
    x0 = torch.from_numpy(gold["qs_x0"]).cuda()
    eps = torch.from_numpy(gold["qs_eps"]).cuda()
    t = torch.from_numpy(gold["qs_t"]).cuda()
    ns = NoiseScheduler()
    out = ns.add_noise(x0, eps, t)
    assert torch.equal(out.cpu(), torch.from_numpy(gold["qs_cos"]))
    out_c = ns.add_noise(x0, eps, t, clamp=3.0)
    assert torch.equal(out_c.cpu(), torch.from_numpy(gold["qs_cos_clamped"]))
    assert torch.equal(LinearNoiseScheduler().add_noise(x0, eps, t).cpu(), torch.from_numpy(gold["qs_lin"]))
    # numpy oracle on the same tables
    ref = O.q_sample(gold["qs_x0"], gold["qs_eps"], gold["qs_t"], gold["cos_sqrt_alphas_cumprod"],
                     gold["cos_sqrt_one_minus_alphas_cumprod"])
    assert np.array_equal(out.cpu().numpy(), ref)


def test_q_sample_offset_views(cuda_device):
    """Contiguous views that start 4 bytes into their storage (a slice of a flat buffer): per-sample size is a multiple of 4 but the
    base is not 16-byte aligned, so the launcher must take the scalar path (a float4 access there is a misaligned-address fault).
    Bit-identical to the aligned call."""
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    ns = NoiseScheduler()
    B, n = 5, 8 * 27 * 27
    flat_x, flat_e = torch.randn(B * n + 3, device="cuda"), torch.randn(B * n + 3, device="cuda")
    t = torch.tensor([0, 1, 500, 998, 999], device="cuda")
    for off in (1, 2, 3):
        x, e = flat_x[off:off + B * n].view(B, 8, 27, 27), flat_e[off:off + B * n].view(B, 8, 27, 27)
        assert x.data_ptr() % 16 != 0 and x.is_contiguous()
        got = ns.add_noise(x, e, t, clamp=3.0)
        want = ns.add_noise(x.clone(), e.clone(), t, clamp=3.0)
        assert torch.equal(got, want)
    torch.cuda.synchronize()


def test_q_sample_edge_cases(cuda_device):
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    ns = NoiseScheduler()
    # empty batch
    e = ns.add_noise(torch.empty(0, 8, 27, 27, device="cuda"), torch.empty(0, 8, 27, 27, device="cuda"),
                     torch.empty(0, dtype=torch.long, device="cuda"))
    assert e.shape == (0, 8, 27, 27)
    # ragged per-sample size (not a multiple of 4) and int32 timesteps
    x = torch.randn(3, 7, device="cuda"); n = torch.randn(3, 7, device="cuda")
    t = torch.tensor([0, 999, 500], dtype=torch.int32, device="cuda")
    out = ns.add_noise(x, n, t)
    a = ns.sqrt_alphas_cumprod[t.long()].view(-1, 1); b = ns.sqrt_one_minus_alphas_cumprod[t.long()].view(-1, 1)
    assert torch.equal(out, a * x + b * n)
    # NaN/Inf -> reference fallback x0 + 0.1*noise (improved_diffusion_trainer.py:61-63)
    x[1, 3] = float("inf")
    out = ns.add_noise(x, n, t)
    assert torch.equal(out, x + 0.1 * n)


def test_q_sample_large_matches_eager(cuda_device):
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    ns = NoiseScheduler().to("cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(256, 8, 27, 27, device="cuda", generator=g); n = torch.randn(256, 8, 27, 27, device="cuda", generator=g)
    t = torch.randint(0, 1000, (256,), device="cuda", generator=g)
    ref = ns.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1) * x + ns.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1) * n
    assert torch.equal(ns.add_noise(x, n, t), ref)


def _drive_ddpm(sched, fast, seed, dev):
    torch.manual_seed(seed)
    x = torch.randn(2, 8, 27, 27).to(dev)
    steps = list(range(0, 1000, 50)) if fast else list(range(1000))
    for t in reversed(steps):
        pred = _stub_unet(x, None, None)
        z = torch.randn(2, 8, 27, 27).to(dev) if t > 0 else None
        x = sched.ddpm_step(x, pred, t, z)
    return x


def test_ddpm_loop_bit_exact(cuda_device, gold):
    from pokemon_sprite_generator_b200.scheduler import NoiseScheduler
    ns = NoiseScheduler()
    assert torch.equal(_drive_ddpm(ns, True, 99, "cuda").cpu(), torch.from_numpy(gold["ddpm_fast"]))
    assert torch.equal(_drive_ddpm(ns, False, 99, "cuda").cpu(), torch.from_numpy(gold["ddpm_full"]))


def test_posterior_step_bit_exact(cuda_device, gold):
    from pokemon_sprite_generator_b200.scheduler import LinearNoiseScheduler
    ls = LinearNoiseScheduler()
    x = torch.from_numpy(gold["post_x"]).cuda(); e = torch.from_numpy(gold["post_eps"]).cuda()
    assert torch.equal(ls.sample_previous_timestep(x, e, 0).cpu(), torch.from_numpy(gold["post_t0"]))
    torch.manual_seed(6)
    z = torch.randn(2, 8, 27, 27).cuda()
    assert torch.equal(ls.sample_previous_timestep(x, e, 999, noise=z).cpu(), torch.from_numpy(gold["post_t999"]))
    # the 50-step loop of FinalPokemonGenerator.forward (final_trainer.py:186-204)
    torch.manual_seed(123)
    lat = torch.randn(2, 8, 27, 27).cuda()
    for i in range(50):
        ts = max(0, 999 - i * 20)
        pred = _stub_unet(lat, None, None)
        lat = ls.sample_previous_timestep(lat, pred, ts, noise=torch.randn(2, 8, 27, 27).cuda())
    assert torch.equal(lat.cpu(), torch.from_numpy(gold["posterior_50"]))


@pytest.mark.parametrize("n", [1, 5, 11664, 256 * 5832])
def test_smooth_l1(cuda_device, n):
    from pokemon_sprite_generator_b200 import losses
    g = torch.Generator(device="cuda").manual_seed(n)
    pred = torch.randn(n, device="cuda", generator=g) * 0.2
    tgt = torch.randn(n, device="cuda", generator=g) * 0.2
    pred[0] = tgt[0]                       # exact zero difference
    loss, grad = losses.smooth_l1_fwd_bwd(pred, tgt, beta=0.1)
    p = pred.clone().requires_grad_(True)
    ref = torch.nn.functional.smooth_l1_loss(p, tgt, beta=0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert torch.allclose(grad, p.grad, rtol=1e-6, atol=1e-12)
    # autograd wrapper
    p2 = pred.clone().requires_grad_(True)
    l2 = losses.SmoothL1Loss(beta=0.1)(p2, tgt)
    (l2 * 3.0).backward()
    assert torch.allclose(p2.grad, 3.0 * p.grad, rtol=1e-6, atol=1e-12)
