"""GPU parity tests of the two GEMM / implicit-conv engines against plain PyTorch fp32 on the same inputs.

The tcgen05 engine takes bf16 operands and accumulates in fp32, so its oracle is an fp32 matmul / conv of the
bf16-rounded operands; tolerance = accumulation-order noise only (rtol 2e-3 of the output scale).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from pokemon_sprite_generator_b200 import _lib as L
    from pokemon_sprite_generator_b200 import gemm as G
    return L, G


def _close(out, ref, tol, what):
    out = out.float()
    scale = ref.abs().max().item() + 1e-12
    err = (out - ref).abs().max().item()
    print(f"[{what}] max_abs_err={err:.3e} scale={scale:.3e} rel={err / scale:.3e}")
    assert err <= tol * scale, f"{what}: err {err} > {tol} * {scale}"


def _check_timeout(L):
    flag = L.load().psg_umma_timeout_flag()
    assert flag == 0, "tcgen05 kernel hit an mbarrier timeout"


@pytest.mark.parametrize("engine,dtype", [("simt", torch.float32), ("simt", torch.bfloat16), ("umma", torch.bfloat16)])
@pytest.mark.parametrize("M,N,K,block_n", [
    (128, 64, 64, 64), (256, 256, 128, 128), (256, 256, 256, 256), (300, 320, 192, 160), (1000, 640, 1280, 0),
    (77, 1280, 640, 0), (4096, 96, 256, 0),
])
def test_gemm_tn(cuda_device, engine, dtype, M, N, K, block_n):
    L, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    b = torch.randn(N, K, device="cuda", generator=g).to(dtype)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=dtype)
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out), engine=engine, block_n=block_n if engine == "umma" else 0)
    torch.cuda.synchronize()
    if engine == "umma":
        _check_timeout(L)
    ref = a.float() @ b.float().t()
    tol = 1e-5 if dtype == torch.float32 else 6e-3   # bf16 output rounding dominates
    _close(out, ref, tol, f"tn {engine} {M}x{N}x{K}")


@pytest.mark.parametrize("engine", ["simt", "umma"])
def test_gemm_fp32_out_and_epilogue(cuda_device, engine):
    L, G = _mods()
    M, N, K, rpg = 392, 640, 320, 196
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    rowbias = torch.randn(M // rpg, N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(a), G.kmajor(b),
               G.Epilogue(out=out, bias=bias, rowbias=rowbias, rows_per_group=rpg, act=L.ACT_GELU, alpha=0.6,
                          residual=res, aux_out=pre), engine=engine)
    torch.cuda.synchronize()
    acc = a.float() @ b.float().t() + bias + rowbias.repeat_interleave(rpg, 0)
    ref = 0.6 * F.gelu(acc) + res.float()
    _close(pre, acc, 6e-3, f"epi-pre {engine}")
    _close(out, ref, 6e-3, f"epi-out {engine}")
    # backward-style epilogue: acc * gelu'(pre), fp32 accumulate into existing buffer
    out32 = torch.ones(M, N, device="cuda")
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out32, aux_in=pre, aux_act=L.ACT_GELU, accumulate=True), engine=engine)
    torch.cuda.synchronize()
    x = pre.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    ref2 = 1.0 + (a.float() @ b.float().t()) * x.grad
    _close(out32, ref2, 2e-3, f"epi-bwd {engine}")
    if engine == "umma":
        _check_timeout(L)


@pytest.mark.parametrize("engine", ["simt", "umma"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (320, 192, 1000), (640, 1280, 777), (96, 64, 4096)])
def test_gemm_nt(cuda_device, engine, M, N, K):
    """C[M,N] = A^T B with A stored [K,M], B stored [K,N] (wgrad layout)."""
    L, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(K, M, device="cuda", generator=g).bfloat16()
    b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    G.run_gemm(G.mnmajor(a), G.mnmajor(b), G.Epilogue(out=out), engine=engine)
    torch.cuda.synchronize()
    if engine == "umma":
        _check_timeout(L)
    ref = a.float().t() @ b.float()
    _close(out, ref, 1e-4, f"nt {engine} {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(256, 128, 64 * 1200), (640, 1280, 64 * 333), (128, 64, 64 * 2000), (384, 320, 64 * 257)])
def test_gemm_nt_stream_k(cuda_device, M, N, K):
    """Few tiles, long reduction: stream-K shares every tile between several CTAs (fp32 partials through the workspace);
    the result must match, twice in a row bit-identically (fixed summation order), and leave the flags clean."""
    L, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(11 + M)
    a = torch.randn(K, M, device="cuda", generator=g).bfloat16()
    b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    outs = []
    for _ in range(2):
        out = torch.full((M, N), float("nan"), device="cuda")
        G.run_gemm(G.mnmajor(a), G.mnmajor(b), G.Epilogue(out=out), engine="umma")
        torch.cuda.synchronize()
        _check_timeout(L)
        outs.append(out)
    ref = (a.double().t() @ b.double()).float()
    _close(outs[0], ref, 1e-4, f"nt stream-k {M}x{N}x{K}")
    assert torch.equal(outs[0], outs[1]), "stream-K result is not run-to-run deterministic"
    flags = next(iter(G._umma_ws.values()))[:1024].view(torch.int32)
    assert int(flags.abs().sum()) == 0, "stream-K flags not handed back"


def test_gemm_simt_split_k(cuda_device):
    L, G = _mods()
    M, N, K, S = 256, 128, 64 * 12, 4
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(K, N, device="cuda", generator=g)
    part = torch.full((S, M, N), float("nan"), device="cuda")
    G.run_gemm(G.kmajor(a), G.mnmajor(b), G.Epilogue(out=part[0]), engine="simt", split_k=S)
    torch.cuda.synchronize()
    _close(part.sum(0), a @ b, 1e-5, "simt split-k")


@pytest.mark.parametrize("M,N,K,bn,mt", [(128 * 37 + 5, 1280, 640, 0, 0), (12544, 1280, 11520, 256, 2), (4096, 1280, 2560, 256, 1),
                                         (3000, 320, 2880, 160, 1), (50176, 640, 640, 0, 0)])
def test_gemm_tn_stream_k(cuda_device, M, N, K, bn, mt):
    """Tile counts that do not divide the SM count: stream-K ranges cross tile boundaries with a fused bf16 epilogue."""
    L, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out, bias=bias, residual=res, alpha=0.7), engine="umma", block_n=bn, m_tiles=mt)
    torch.cuda.synchronize()
    _check_timeout(L)
    _close(out, 0.7 * (a.float() @ b.float().t() + bias) + res.float(), 6e-3, f"tn stream-k {M}x{N}x{K}")


@pytest.mark.parametrize("engine", ["simt", "umma"])
@pytest.mark.parametrize("act", ["gelu", "silu"])
def test_epilogue_saved_derivative(cuda_device, engine, act):
    """Forward epilogue stores act'(pre) * dropmask / (1-p) (aux_act set on the forward side); backward multiplies by it
    (PSG_ACT_MUL).  Same rule on both engines, same dropout mask in both passes."""
    L, G = _mods()
    M, N, K, p = 392, 640, 320, 0.25
    g = torch.Generator(device="cuda").manual_seed(17)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    code = L.ACT_GELU if act == "gelu" else L.ACT_SILU
    fn = F.gelu if act == "gelu" else F.silu
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out, bias=bias, act=code, aux_out=aux, aux_act=code, drop_seed=99, drop_p=p),
               engine=engine)
    torch.cuda.synchronize()
    x = (a.float() @ b.float().t() + bias).requires_grad_(True)
    y = fn(x)
    y.sum().backward()
    mask = (out.float() != 0) | (y.detach().abs() < 1e-3)
    frac = 1.0 - (out.float() != 0).float().mean().item()
    assert abs(frac - p) < 0.02, f"dropout rate {frac}"
    _close(out.float(), torch.where(out.float() != 0, y.detach() / (1 - p), torch.zeros_like(y)), 8e-3, f"saved-grad out {engine} {act}")
    keep = (out.float() != 0).float()
    sel = y.detach().abs() > 1e-3            # where the mask can be read off the output
    _close(aux.float()[sel], (x.grad * keep / (1 - p))[sel], 8e-3, f"saved-grad aux {engine} {act}")
    # backward: dX-like product times the saved derivative
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=o2, aux_in=aux, aux_act=L.ACT_MUL), engine=engine)
    torch.cuda.synchronize()
    _close(o2.float(), (a.float() @ b.float().t()) * aux.float(), 8e-3, f"saved-grad bwd {engine} {act}")
    if engine == "umma":
        _check_timeout(L)


def _conv_inputs(n, h, w, cin, cout, seed, dtype):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(dtype)            # NHWC
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).to(dtype)  # OIHW
    return x, wt


@pytest.mark.parametrize("engine,dtype", [("simt", torch.float32), ("umma", torch.bfloat16)])
@pytest.mark.parametrize("n,h,cin,cout,stride", [(2, 27, 64, 64, 1), (3, 14, 128, 320, 1), (2, 7, 320, 128, 1), (5, 4, 64, 160, 1),
                                                 (2, 27, 64, 128, 2), (2, 14, 128, 64, 2), (3, 7, 64, 64, 2)])
def test_conv_fprop(cuda_device, engine, dtype, n, h, cin, cout, stride):
    L, G = _mods()
    x, wt = _conv_inputs(n, h, h, cin, cout, 3 * h + cin + cout, dtype)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()   # [Cout][tap][Cin]
    a = G.im2col(x, 3, stride, 1)
    out = torch.full((a.rows, cout), float("nan"), device="cuda", dtype=dtype)
    bias = torch.randn(cout, device="cuda")
    G.run_gemm(a, G.kmajor(wp), G.Epilogue(out=out, bias=bias), engine=engine)
    torch.cuda.synchronize()
    if engine == "umma":
        _check_timeout(L)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, cout)
    _close(out, ref, 1e-5 if dtype == torch.float32 else 6e-3, f"fprop {engine} n{n} h{h} {cin}->{cout} s{stride}")


@pytest.mark.parametrize("engine,dtype", [("simt", torch.float32), ("umma", torch.bfloat16)])
@pytest.mark.parametrize("n,h,cin,cout", [(2, 27, 64, 64), (3, 14, 128, 320), (4, 4, 64, 192)])
def test_conv_dgrad_stride1(cuda_device, engine, dtype, n, h, cin, cout):
    L, G = _mods()
    x, wt = _conv_inputs(n, h, h, cin, cout, 17 + h, dtype)
    dy = torch.randn(n, h, h, cout, device="cuda").to(dtype)
    wd = wt.permute(1, 2, 3, 0).reshape(cin, 9 * cout).contiguous()   # [Cin][tap][Cout]
    a = G.im2col(dy, 3, 1, 1, flip=True)
    dx = torch.full((n * h * h, cin), float("nan"), device="cuda", dtype=dtype)
    G.run_gemm(a, G.kmajor(wd), G.Epilogue(out=dx), engine=engine)
    torch.cuda.synchronize()
    if engine == "umma":
        _check_timeout(L)
    ref = torch.nn.grad.conv2d_input((n, cin, h, h), wt.float(), dy.float().permute(0, 3, 1, 2), stride=1, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, cin)
    _close(dx, ref, 1e-5 if dtype == torch.float32 else 6e-3, f"dgrad {engine} h{h} {cin}<-{cout}")


@pytest.mark.parametrize("h,cin,cout", [(27, 64, 128), (14, 32, 64), (7, 64, 64)])
def test_conv_dgrad_stride2_gather(cuda_device, h, cin, cout):
    """General-stride dgrad gather (SIMT engine)."""
    L, G = _mods()
    n = 2
    _, wt = _conv_inputs(n, h, h, cin, cout, 23 + h, torch.float32)
    p = (h + 2 - 3) // 2 + 1
    dy = torch.randn(n, p, p, cout, device="cuda")
    wd = wt.permute(1, 2, 3, 0).reshape(cin, 9 * cout).contiguous()
    dx = torch.empty(n * h * h, cin, device="cuda")
    G.run_gemm(G.dgrad_gather(dy, h, h, 3, 2, 1), G.kmajor(wd), G.Epilogue(out=dx), engine="simt")
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((n, cin, h, h), wt, dy.permute(0, 3, 1, 2), stride=2, padding=1)
    _close(dx, ref.permute(0, 2, 3, 1).reshape(-1, cin), 1e-5, f"dgrad-s2 h{h}")


@pytest.mark.parametrize("engine,dtype", [("simt", torch.float32), ("umma", torch.bfloat16)])
@pytest.mark.parametrize("n,h,cin,cout,stride,split", [(2, 27, 64, 128, 1, 1), (3, 14, 128, 64, 1, 1), (8, 4, 320, 128, 1, 1),
                                                       (2, 27, 64, 64, 2, 1), (40, 27, 64, 128, 1, 1)])
def test_conv_wgrad(cuda_device, engine, dtype, n, h, cin, cout, stride, split):
    L, G = _mods()
    x, wt = _conv_inputs(n, h, h, cin, cout, 31 + h + cin, dtype)
    b = G.im2col_t(x, 3, stride, 1)
    p = (h + 2 - 3) // stride + 1
    dy = torch.randn(n * p * p, cout, device="cuda").to(dtype)
    part = torch.full((split, cout, 9 * cin), float("nan"), device="cuda")
    G.run_gemm(G.mnmajor(dy), b, G.Epilogue(out=part[0]), engine=engine, split_k=split)
    torch.cuda.synchronize()
    if engine == "umma":
        _check_timeout(L)
    dw = part.sum(0).reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                      dy.float().reshape(n, p, p, cout).permute(0, 3, 1, 2), stride=stride, padding=1)
    _close(dw, ref, 1e-4, f"wgrad {engine} h{h} {cin}->{cout} s{stride} split{split}")


@pytest.mark.parametrize("n,h,cin,cout", [(2, 14, 640, 640), (24, 14, 640, 640), (4, 8, 1280, 1280), (40, 8, 1280, 1280), (16, 27, 320, 320)])
def test_conv_wgrad_half_row_tiles(cuda_device, n, h, cin, cout):
    """Conv weight gradients whose last 256-row CTA tile keeps only its lower 128-row sub-tile (Cout = 640, 1280: the upper one is
    neither loaded, multiplied nor drained -- umma_gemm.cu live_subtiles), with and without shared (stream-K) tiles; bit-identical
    across repeats."""
    L, G = _mods()
    x, wt = _conv_inputs(n, h, h, cin, cout, 7 + h + cin, torch.bfloat16)
    dy = torch.randn(n * h * h, cout, device="cuda").bfloat16()
    outs = []
    for _ in range(2):
        dw = torch.full((cout, 9 * cin), float("nan"), device="cuda")
        G.run_gemm(G.mnmajor(dy), G.im2col_t(x, 3, 1, 1), G.Epilogue(out=dw), engine="umma")
        torch.cuda.synchronize()
        _check_timeout(L)
        outs.append(dw)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                      dy.float().reshape(n, h, h, cout).permute(0, 3, 1, 2), stride=1, padding=1)
    _close(outs[0].reshape(cout, 3, 3, cin).permute(0, 3, 1, 2), ref, 1e-4, f"wgrad half row tile h{h} {cin}->{cout}")
    assert torch.equal(outs[0], outs[1]), "wgrad is not run-to-run deterministic"


@pytest.mark.parametrize("M,N,K,bn,mt", [(392, 640, 1280, 0, 0), (12544, 1280, 2560, 256, 2), (1000, 320, 640, 0, 0), (300, 96, 128, 0, 0)])
def test_gemm_tt_weight_in_place(cuda_device, M, N, K, bn, mt):
    """dgrad of a Linear with the [N_w, K_w] weight read transposed in place (A K-major, B MN-major)."""
    L, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    dy = torch.randn(M, K, device="cuda", generator=g).bfloat16()                       # K = N_w
    w = (torch.randn(K, N, device="cuda", generator=g) / K ** 0.5).bfloat16()           # weight [N_w, K_w]: here [K, N]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(dy), G.mnmajor(w), G.Epilogue(out=out), engine="umma", block_n=bn, m_tiles=mt)
    torch.cuda.synchronize()
    _check_timeout(L)
    _close(out, dy.float() @ w.float(), 6e-3, f"tt {M}x{N}x{K}")


@pytest.mark.parametrize("n,h,cin,cout", [(2, 27, 320, 320), (3, 14, 640, 320), (8, 7, 1280, 1280), (5, 4, 128, 64), (2, 14, 64, 640)])
def test_conv_dgrad_weight_in_place(cuda_device, n, h, cin, cout):
    """Stride-1 dgrad with the fprop weight matrix [Cout][tap][Cin] read transposed in place (PSG_OP_CONVW_T)."""
    L, G = _mods()
    x, wt = _conv_inputs(n, h, h, cin, cout, 91 + h + cin, torch.bfloat16)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    dy = torch.randn(n, h, h, cout, device="cuda").bfloat16()
    dx = torch.full((n * h * h, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.im2col(dy, 3, 1, 1, flip=True), G.convw_t(wp, cin, 3), G.Epilogue(out=dx), engine="umma")
    torch.cuda.synchronize()
    _check_timeout(L)
    ref = torch.nn.grad.conv2d_input((n, cin, h, h), wt.float(), dy.float().permute(0, 3, 1, 2), stride=1, padding=1)
    _close(dx, ref.permute(0, 2, 3, 1).reshape(-1, cin), 6e-3, f"dgrad in place h{h} {cin}<-{cout}")


def test_strided_views(cuda_device):
    """Operands and outputs that are channel slices of wider (concat) buffers."""
    L, G = _mods()
    M, C0, C1, N = 392, 128, 64, 128
    cat = torch.randn(M, C0 + C1, device="cuda").bfloat16()
    w = torch.randn(N, C1, device="cuda").bfloat16()
    outbuf = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(cat[:, C0:]), G.kmajor(w), G.Epilogue(out=outbuf[:, N:]), engine="umma")
    torch.cuda.synchronize()
    _check_timeout(L)
    _close(outbuf[:, N:], cat[:, C0:].float() @ w.float().t(), 6e-3, "strided")
    assert outbuf[:, :N].abs().max().item() == 0.0


@pytest.mark.parametrize("mt", [1, 2])
@pytest.mark.parametrize("bn", [64, 128, 160, 256])
def test_umma_tile_shapes_tn(cuda_device, mt, bn):
    """Every CTA tile shape of the persistent tcgen05 kernel (128*mt x bn), multiple work items per CTA."""
    L, G = _mods()
    M, N, K = 128 * 150 * mt + 77, 640, 320          # > 148 tiles so that CTAs loop; ragged M; N not a multiple of 256
    g = torch.Generator(device="cuda").manual_seed(mt * 1000 + bn)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.kmajor(a), G.kmajor(b), G.Epilogue(out=out, bias=bias), engine="umma", block_n=bn, m_tiles=mt)
    torch.cuda.synchronize()
    _check_timeout(L)
    _close(out, a.float() @ b.float().t() + bias, 6e-3, f"tile {128 * mt}x{bn}")


@pytest.mark.parametrize("mt", [1, 2])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_umma_tile_shapes_wgrad(cuda_device, mt, bn):
    L, G = _mods()
    n, h, cin, cout, split = 6, 14, 320, 384, 1
    x, wt = _conv_inputs(n, h, h, cin, cout, 77 + bn + mt, torch.bfloat16)
    dy = torch.randn(n * h * h, cout, device="cuda").bfloat16()
    part = torch.full((split, cout, 9 * cin), float("nan"), device="cuda")
    G.run_gemm(G.mnmajor(dy), G.im2col_t(x, 3, 1, 1), G.Epilogue(out=part[0]), engine="umma", block_n=bn, m_tiles=mt)
    torch.cuda.synchronize()
    _check_timeout(L)
    dw = part.sum(0).reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.float().reshape(n, h, h, cout).permute(0, 3, 1, 2),
                                      stride=1, padding=1)
    _close(dw, ref, 1e-4, f"wgrad tile {128 * mt}x{bn}")


@pytest.mark.parametrize("mt", [1, 2])
def test_umma_tile_shapes_conv(cuda_device, mt):
    L, G = _mods()
    n, h, cin, cout = 40, 27, 64, 320
    x, wt = _conv_inputs(n, h, h, cin, cout, 5 + mt, torch.bfloat16)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.full((n * h * h, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    G.run_gemm(G.im2col(x, 3, 1, 1), G.kmajor(wp), G.Epilogue(out=out), engine="umma", m_tiles=mt)
    torch.cuda.synchronize()
    _check_timeout(L)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, cout)
    _close(out, ref, 6e-3, f"conv tile mt={mt}")
