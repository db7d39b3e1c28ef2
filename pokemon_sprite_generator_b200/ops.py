"""Thin tensor-level wrappers over the C ABI (one function per exported kernel family).

Every tensor argument is a CUDA tensor; token-major activations are 2-D views [rows, C] with unit column
stride and arbitrary row pitch (channel slices of concat buffers are fine).  Nothing here computes on the
host and nothing falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_ws_cache = {}


def workspace(device, floats: int, tag: str = "ws") -> torch.Tensor:
    """Grow-only fp32 scratch buffer per (device, tag); stream-ordered reuse."""
    key = (device, tag)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < floats:
        ws = torch.zeros(max(floats, 4096), dtype=torch.float32, device=device)
        _ws_cache[key] = ws
    return ws


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return t.stride(0)


# ---- layout -------------------------------------------------------------------------------------------------
def nchw_to_tokens(src: torch.Tensor, dst: torch.Tensor) -> None:
    b, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    L.call("psg_nchw_to_tokens", L.ptr(src), L.ptr(dst), C.c_longlong(_ld(dst)), C.c_int(b), C.c_int(c), C.c_int(h * w),
           C.c_int(L.dt(dst)), L.stream_ptr())


def tokens_to_nchw(src: torch.Tensor, dst: torch.Tensor) -> None:
    b, c, h, w = dst.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    L.call("psg_tokens_to_nchw", L.ptr(src), C.c_longlong(_ld(src)), L.ptr(dst), C.c_int(b), C.c_int(c), C.c_int(h * w),
           C.c_int(L.dt(src)), L.stream_ptr())


def copy_strided(src: torch.Tensor, dst: torch.Tensor, accumulate: bool = False) -> None:
    assert src.shape == dst.shape and src.dtype == dst.dtype
    L.call("psg_copy_strided", L.ptr(src), C.c_longlong(_ld(src)), L.ptr(dst), C.c_longlong(_ld(dst)), C.c_longlong(src.shape[0]),
           C.c_int(src.shape[1]), C.c_int(int(accumulate)), C.c_int(L.dt(src)), L.stream_ptr())


def colsum(x: torch.Tensor, groups: int, out_groups: torch.Tensor | None, out_total: torch.Tensor | None,
           acc_groups: bool = False, acc_total: bool = False, scale: float = 1.0) -> None:
    rows, c = x.shape
    assert rows % groups == 0
    rpg = rows // groups
    s = L.load().psg_colsum_slices(C.c_int(groups), C.c_int(rpg))
    # zero-initialised: [0] is the fused kernel's counter.  One buffer PER STREAM: backward runs bias column sums on its
    # second stream (engine._weight_stream) next to the main stream's conditioning column sums, and two kernels sharing the
    # counter / partial slots corrupt each other's sums (seen as a wrong init_conv.bias gradient in the batch-256 parity test)
    ws = workspace(x.device, 32 + groups * s * min(c, 4096), f"colsum@{L.raw_stream(x.device)}")
    L.call("psg_colsum", L.ptr(x), C.c_longlong(_ld(x)), C.c_int(groups), C.c_int(rpg), C.c_int(c), L.ptr(out_groups),
           C.c_longlong(out_groups.stride(0) if out_groups is not None else 0), C.c_int(int(acc_groups)), L.ptr(out_total),
           C.c_int(int(acc_total)), C.c_float(scale), L.ptr(ws), C.c_int(L.dt(x)), L.stream_ptr())


def upsample_fwd(x: torch.Tensor, y: torch.Tensor, b: int, ih: int, iw: int, oh: int, ow: int) -> None:
    L.call("psg_upsample_bilinear_fwd", L.ptr(x), C.c_longlong(_ld(x)), L.ptr(y), C.c_longlong(_ld(y)), C.c_int(b),
           C.c_int(x.shape[1]), C.c_int(ih), C.c_int(iw), C.c_int(oh), C.c_int(ow), C.c_int(L.dt(x)), L.stream_ptr())


def upsample_bwd(dy: torch.Tensor, dx: torch.Tensor, b: int, ih: int, iw: int, oh: int, ow: int, accumulate: bool) -> None:
    L.call("psg_upsample_bilinear_bwd", L.ptr(dy), C.c_longlong(_ld(dy)), L.ptr(dx), C.c_longlong(_ld(dx)), C.c_int(b),
           C.c_int(dx.shape[1]), C.c_int(ih), C.c_int(iw), C.c_int(oh), C.c_int(ow), C.c_int(int(accumulate)), C.c_int(L.dt(dx)),
           L.stream_ptr())


def dilate2(dy: torch.Tensor, out: torch.Tensor, b: int, p: int, q: int, h: int, w: int) -> None:
    L.call("psg_dilate2", L.ptr(dy), C.c_longlong(_ld(dy)), L.ptr(out), C.c_longlong(_ld(out)), C.c_int(b), C.c_int(dy.shape[1]),
           C.c_int(p), C.c_int(q), C.c_int(h), C.c_int(w), C.c_int(L.dt(dy)), L.stream_ptr())


def dgrad_s2_weights(wd: torch.Tensor, w00: torch.Tensor, w01: torch.Tensor, w10: torch.Tensor, w11: torch.Tensor, cin: int, cout: int) -> None:
    """wd [Cin, 9*Cout] (dgrad layout) -> the four parity-class weights of a 3x3 stride-2 pad-1 conv's dgrad."""
    L.call("psg_dgrad_s2_weights", L.ptr(wd), L.ptr(w00), L.ptr(w01), L.ptr(w10), L.ptr(w11), C.c_int(cin), C.c_int(cout), C.c_int(L.dt(wd)),
           L.stream_ptr())


def interleave2x2(c00, c01, c10, c11, out: torch.Tensor, b: int, p: int, q: int, h: int, w: int, accumulate: bool) -> None:
    L.call("psg_interleave2x2", L.ptr(c00), L.ptr(c01), L.ptr(c10), L.ptr(c11), L.ptr(out), C.c_longlong(_ld(out)), C.c_int(b),
           C.c_int(out.shape[1]), C.c_int(p), C.c_int(q), C.c_int(h), C.c_int(w), C.c_int(int(accumulate)), C.c_int(L.dt(out)), L.stream_ptr())


def dropout_scale(x: torch.Tensor, out: torch.Tensor, alpha: float, seed: int, drop_p: float) -> None:
    L.call("psg_dropout_scale", L.ptr(x), C.c_longlong(_ld(x)), L.ptr(out), C.c_longlong(_ld(out)), C.c_longlong(x.shape[0]),
           C.c_int(x.shape[1]), C.c_float(alpha), C.c_ulonglong(seed), C.c_float(drop_p), C.c_int(L.dt(x)), L.stream_ptr())


def timestep_embedding(t: torch.Tensor, coeff: torch.Tensor, out: torch.Tensor) -> None:
    assert t.dtype == torch.int64 and coeff.dtype == torch.float32 and out.is_contiguous()
    L.call("psg_timestep_embedding", L.ptr(t), L.ptr(coeff), L.ptr(out), C.c_int(t.shape[0]), C.c_int(coeff.shape[0]), L.stream_ptr())


def mean_pool(x: torch.Tensor, out: torch.Tensor) -> None:
    b, l, d = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    L.call("psg_mean_pool", L.ptr(x), L.ptr(out), C.c_int(b), C.c_int(l), C.c_int(d), L.stream_ptr())


# ---- weights ------------------------------------------------------------------------------------------------
def pack_conv_weight(w: torch.Tensor, wp: torch.Tensor | None, wd: torch.Tensor | None, cin_p: int = 0, cout_p: int = 0) -> None:
    """wp [cout_p, kk*cin_p], wd [cin_p, kk*cout_p]; padded entries (cin_p > Cin / cout_p > Cout) must be pre-zeroed."""
    cout, cin, kh, kw = w.shape
    ref = wp if wp is not None else wd
    L.call("psg_pack_conv_weight", L.ptr(w), L.ptr(wp), L.ptr(wd), C.c_int(cout), C.c_int(cin), C.c_int(kh * kw),
           C.c_int(cin_p or cin), C.c_int(cout_p or cout), C.c_int(L.dt(ref)), L.stream_ptr())


def conv_weights_transpose(src_base: torch.Tensor, dst_base: torch.Tensor, jobs) -> None:
    """jobs: list of (src element offset, dst element offset, Cout, Cin, kk); one launch for all of them."""
    assert src_base.dtype == torch.bfloat16 and dst_base.dtype == torch.bfloat16
    arr = (C.c_longlong * (5 * len(jobs)))(*[int(v) for j in jobs for v in j])
    L.call("psg_conv_weights_transpose", L.ptr(src_base), L.ptr(dst_base), arr, C.c_int(len(jobs)), L.stream_ptr())


def pack_linear_weight(w: torch.Tensor, wk: torch.Tensor | None, wt: torch.Tensor | None) -> None:
    n, k = w.shape
    ref = wk if wk is not None else wt
    L.call("psg_pack_linear_weight", L.ptr(w), L.ptr(wk), L.ptr(wt), C.c_int(n), C.c_int(k), C.c_int(L.dt(ref)), L.stream_ptr())


def wgrad_finalize(partial: torch.Tensor, splits: int, split_stride: int, grad: torch.Tensor, accumulate: bool = False,
                   cin_p: int = 0) -> None:
    cout, cin, kh, kw = grad.shape
    L.call("psg_wgrad_finalize", L.ptr(partial), C.c_int(splits), C.c_longlong(split_stride), L.ptr(grad), C.c_int(cout), C.c_int(cin),
           C.c_int(kh * kw), C.c_int(cin_p or cin), C.c_int(int(accumulate)), L.stream_ptr())


def sum_partials(partial: torch.Tensor, splits: int, split_stride: int, out: torch.Tensor, accumulate: bool = False) -> None:
    assert out.is_contiguous()
    L.call("psg_sum_partials", L.ptr(partial), C.c_int(splits), C.c_longlong(split_stride), L.ptr(out), C.c_longlong(out.numel()),
           C.c_int(int(accumulate)), L.stream_ptr())


# ---- GroupNorm ----------------------------------------------------------------------------------------------
def groupnorm_fwd(x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats: torch.Tensor, b: int,
                  groups: int, eps: float, silu: bool) -> None:
    rows, c = x.shape
    hw = rows // b
    s = L.load().psg_groupnorm_slices(C.c_int(b), C.c_int(hw))
    ws = workspace(x.device, b * s * groups * 2, "gn")
    L.call("psg_groupnorm_fwd", L.ptr(x), C.c_longlong(_ld(x)), L.ptr(y), C.c_longlong(_ld(y)), L.ptr(gamma), L.ptr(beta), L.ptr(stats),
           L.ptr(ws), C.c_int(b), C.c_int(hw), C.c_int(c), C.c_int(groups), C.c_float(eps), C.c_int(int(silu)), C.c_int(L.dt(x)),
           L.stream_ptr())


def groupnorm_bwd(dy: torch.Tensor, x: torch.Tensor, dx: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats: torch.Tensor,
                  dgamma: torch.Tensor, dbeta: torch.Tensor, b: int, groups: int, silu: bool, accumulate_dx: bool) -> None:
    rows, c = x.shape
    hw = rows // b
    s = L.load().psg_groupnorm_slices(C.c_int(b), C.c_int(hw))
    ws = workspace(x.device, b * s * c * 2, "gn")
    L.call("psg_groupnorm_bwd", L.ptr(dy), C.c_longlong(_ld(dy)), L.ptr(x), C.c_longlong(_ld(x)), L.ptr(dx), C.c_longlong(_ld(dx)),
           L.ptr(gamma), L.ptr(beta), L.ptr(stats), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), C.c_int(b), C.c_int(hw), C.c_int(c),
           C.c_int(groups), C.c_int(int(silu)), C.c_int(L.dt(x)), C.c_int(int(accumulate_dx)), C.c_int(0), L.stream_ptr())


_gn_fused_cache = {}


def groupnorm_fused_ok(b: int, hw: int, c: int, groups: int, dtype: torch.dtype) -> bool:
    """True when the single-pass (slab in shared memory) GroupNorm kernels handle this problem."""
    key = (b, hw, c, groups, dtype)
    ok = _gn_fused_cache.get(key)
    if ok is None:
        ok = dtype == torch.bfloat16 and bool(L.load().psg_groupnorm_fused_ok(C.c_int(b), C.c_int(hw), C.c_int(c), C.c_int(groups),
                                                                              C.c_int(L.DT_BF16)))
        _gn_fused_cache[key] = ok
    return ok


def groupnorm_fused_fwd(x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats: torch.Tensor, b: int,
                        groups: int, eps: float, silu: bool) -> None:
    rows, c = x.shape
    L.call("psg_groupnorm_fused_fwd", L.ptr(x), C.c_longlong(_ld(x)), L.ptr(y), C.c_longlong(_ld(y)), L.ptr(gamma), L.ptr(beta),
           L.ptr(stats), C.c_int(b), C.c_int(rows // b), C.c_int(c), C.c_int(groups), C.c_float(eps), C.c_int(int(silu)), L.stream_ptr())


def groupnorm_fused_bwd(dy: torch.Tensor, x: torch.Tensor, dx: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                        stats: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor, b: int, groups: int, silu: bool,
                        accumulate_dx: bool, dx_colsum: torch.Tensor | None = None, bias_total: torch.Tensor | None = None) -> None:
    """dx_colsum [b, C] (row pitch free) / bias_total [C]: optional sums over pixels of dx per (sample, channel) / per channel."""
    rows, c = x.shape
    lib = L.load()
    lib.psg_groupnorm_bwd_workspace_floats.restype = C.c_longlong
    need = int(lib.psg_groupnorm_bwd_workspace_floats(C.c_int(b), C.c_int(rows // b), C.c_int(c), C.c_int(groups)))
    ws = workspace(x.device, max(need, b * c * 3), "gn")
    L.call("psg_groupnorm_fused_bwd_ws", L.ptr(dy), C.c_longlong(_ld(dy)), L.ptr(x), C.c_longlong(_ld(x)), L.ptr(dx), C.c_longlong(_ld(dx)),
           L.ptr(gamma), L.ptr(beta), L.ptr(stats), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), C.c_longlong(ws.numel()), L.ptr(dx_colsum),
           C.c_longlong(dx_colsum.stride(0) if dx_colsum is not None else 0), L.ptr(bias_total), C.c_int(b), C.c_int(rows // b),
           C.c_int(c), C.c_int(groups), C.c_int(int(silu)), C.c_int(int(accumulate_dx)), C.c_int(0), L.stream_ptr())


# ---- attention ----------------------------------------------------------------------------------------------
def attn_fwd(q, k, v, o, lse, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0, drop_p: float = 0.0) -> None:
    L.call("psg_attn_fwd", L.ptr(q), C.c_longlong(_ld(q)), L.ptr(k), C.c_longlong(_ld(k)), L.ptr(v), C.c_longlong(_ld(v)), L.ptr(o),
           C.c_longlong(_ld(o)), L.ptr(lse), C.c_int(b), C.c_int(heads), C.c_int(lq), C.c_int(lk), C.c_int(hd),
           C.c_float(1.0 / (hd ** 0.5)), C.c_int(L.dt(q)), C.c_ulonglong(drop_seed), C.c_float(drop_p), L.stream_ptr())


def attn_bwd(q, k, v, o, do, lse, dq, dk, dv, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0,
             drop_p: float = 0.0) -> None:
    dsum = workspace(q.device, b * heads * lq, "attn_dsum")
    L.call("psg_attn_bwd", L.ptr(q), C.c_longlong(_ld(q)), L.ptr(k), C.c_longlong(_ld(k)), L.ptr(v), C.c_longlong(_ld(v)), L.ptr(o),
           C.c_longlong(_ld(o)), L.ptr(do), C.c_longlong(_ld(do)), L.ptr(lse), L.ptr(dsum), L.ptr(dq), C.c_longlong(_ld(dq)), L.ptr(dk),
           C.c_longlong(_ld(dk)), L.ptr(dv), C.c_longlong(_ld(dv)), C.c_int(b), C.c_int(heads), C.c_int(lq), C.c_int(lk), C.c_int(hd),
           C.c_float(1.0 / (hd ** 0.5)), C.c_int(L.dt(q)), C.c_ulonglong(drop_seed), C.c_float(drop_p), L.stream_ptr())


# ---- fused tensor-core attention (bf16): scores never leave the chip --------------------------------------------------
_attn_fused_cache = {}


def attn_fused_ok(b: int, heads: int, lq: int, lk: int, hd: int) -> bool:
    key = (b, heads, lq, lk, hd)
    ok = _attn_fused_cache.get(key)
    if ok is None:
        ok = bool(L.load().psg_attn_fused_ok(C.c_int(b), C.c_int(heads), C.c_int(lq), C.c_int(lk), C.c_int(hd)))
        _attn_fused_cache[key] = ok
    return ok


def attn_fused_fwd(q, k, v, o, lse, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0, drop_p: float = 0.0) -> None:
    L.call("psg_attn_fused_fwd", L.ptr(q), C.c_longlong(_ld(q)), L.ptr(k), C.c_longlong(_ld(k)), L.ptr(v), C.c_longlong(_ld(v)), L.ptr(o),
           C.c_longlong(_ld(o)), L.ptr(lse), C.c_int(b), C.c_int(heads), C.c_int(lq), C.c_int(lk), C.c_int(hd),
           C.c_float(1.0 / (hd ** 0.5)), C.c_ulonglong(drop_seed), C.c_float(drop_p), L.stream_ptr())


def attn_fused_bwd(q, k, v, o, do, lse, dq, dk, dv, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0,
                   drop_p: float = 0.0) -> None:
    delta = workspace(q.device, b * heads * lq, "attn_dsum")
    L.call("psg_attn_fused_bwd", L.ptr(q), C.c_longlong(_ld(q)), L.ptr(k), C.c_longlong(_ld(k)), L.ptr(v), C.c_longlong(_ld(v)), L.ptr(o),
           C.c_longlong(_ld(o)), L.ptr(do), C.c_longlong(_ld(do)), L.ptr(lse), L.ptr(delta), L.ptr(dq), C.c_longlong(_ld(dq)), L.ptr(dk),
           C.c_longlong(_ld(dk)), L.ptr(dv), C.c_longlong(_ld(dv)), C.c_int(b), C.c_int(heads), C.c_int(lq), C.c_int(lk), C.c_int(hd),
           C.c_float(1.0 / (hd ** 0.5)), C.c_ulonglong(drop_seed), C.c_float(drop_p), L.stream_ptr())


# ---- tensor-core attention (bf16): batched mma.sync GEMMs + row softmax ------------------------------------------
def _bmm(a, a_sb, a_sh, lda, ta, b, b_sb, b_sh, ldb, tb, c, c_sb, c_sh, ldc, batch, heads, m, n, k, alpha):
    L.call("psg_bmm_bf16", L.ptr(a), C.c_longlong(a_sb), C.c_longlong(a_sh), C.c_longlong(lda), C.c_int(ta), L.ptr(b),
           C.c_longlong(b_sb), C.c_longlong(b_sh), C.c_longlong(ldb), C.c_int(tb), L.ptr(c), C.c_longlong(c_sb), C.c_longlong(c_sh),
           C.c_longlong(ldc), C.c_int(int(c.dtype == torch.float32)), C.c_int(batch), C.c_int(heads), C.c_int(m), C.c_int(n),
           C.c_int(k), C.c_float(alpha), L.stream_ptr())


def attn_tc_fwd(q, k, v, o, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0, drop_p: float = 0.0):
    """Returns P (bf16 [b*heads*lq, lkp], un-dropped softmax) to be saved for backward."""
    lkp = (lk + 7) // 8 * 8
    rows = b * heads * lq
    scale = 1.0 / (hd ** 0.5)
    S = workspace(q.device, rows * lkp, "attn_s").narrow(0, 0, rows * lkp)
    _bmm(q, lq * _ld(q), hd, _ld(q), 0, k, lk * _ld(k), hd, _ld(k), 0, S, heads * lq * lkp, lq * lkp, lkp, b, heads, lq, lk, hd, scale)
    P = torch.empty(rows, lkp, dtype=torch.bfloat16, device=q.device)
    Pd = torch.empty_like(P) if drop_p > 0.0 else None
    L.call("psg_softmax_fwd", L.ptr(S), L.ptr(P), L.ptr(Pd), C.c_longlong(rows), C.c_int(lk), C.c_int(lkp), C.c_ulonglong(drop_seed),
           C.c_float(drop_p), L.stream_ptr())
    pv = Pd if Pd is not None else P
    _bmm(pv, heads * lq * lkp, lq * lkp, lkp, 0, v, lk * _ld(v), hd, _ld(v), 1, o, lq * _ld(o), hd, _ld(o), b, heads, lq, hd, lk, 1.0)
    return P


def attn_tc_bwd(q, k, v, do, P, dq, dk, dv, b: int, heads: int, lq: int, lk: int, hd: int, drop_seed: int = 0,
                drop_p: float = 0.0) -> None:
    lkp = P.shape[1]
    rows = b * heads * lq
    scale = 1.0 / (hd ** 0.5)
    bs, hs = heads * lq * lkp, lq * lkp
    dPd = workspace(q.device, rows * lkp, "attn_s").narrow(0, 0, rows * lkp)
    _bmm(do, lq * _ld(do), hd, _ld(do), 0, v, lk * _ld(v), hd, _ld(v), 0, dPd, bs, hs, lkp, b, heads, lq, lk, hd, 1.0)
    dS = torch.empty(rows, lkp, dtype=torch.bfloat16, device=q.device)
    Pd = torch.empty_like(dS) if drop_p > 0.0 else None
    L.call("psg_softmax_bwd", L.ptr(P), L.ptr(dPd), L.ptr(dS), L.ptr(Pd), C.c_longlong(rows), C.c_int(lk), C.c_int(lkp),
           C.c_ulonglong(drop_seed), C.c_float(drop_p), L.stream_ptr())
    pv = Pd if Pd is not None else P
    _bmm(pv, bs, hs, lkp, 1, do, lq * _ld(do), hd, _ld(do), 1, dv, lk * _ld(dv), hd, _ld(dv), b, heads, lk, hd, lq, 1.0)
    _bmm(dS, bs, hs, lkp, 0, k, lk * _ld(k), hd, _ld(k), 1, dq, lq * _ld(dq), hd, _ld(dq), b, heads, lq, hd, lk, scale)
    _bmm(dS, bs, hs, lkp, 1, q, lq * _ld(q), hd, _ld(q), 1, dk, lk * _ld(dk), hd, _ld(dk), b, heads, lk, hd, lq, scale)


# ---- optimiser ----------------------------------------------------------------------------------------------
def sumsq(x: torch.Tensor, out: torch.Tensor, accumulate: bool = False) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous()
    ws = workspace(x.device, 1024 + 8, "sumsq")
    L.call("psg_sumsq", L.ptr(x), C.c_longlong(x.numel()), L.ptr(out), C.c_int(int(accumulate)), L.ptr(ws), L.stream_ptr())


def clip_coef(sumsq_t: torch.Tensor, max_norm: float, state: torch.Tensor, count_steps: bool = False) -> None:
    """state[0..2] = total norm, clip coefficient, finite flag; count_steps: state[3] += 1 when the step will be applied."""
    if count_steps:
        assert state.numel() >= 4
        L.call("psg_clip_coef_count", L.ptr(sumsq_t), C.c_float(max_norm), L.ptr(state), L.stream_ptr())
    else:
        L.call("psg_clip_coef", L.ptr(sumsq_t), C.c_float(max_norm), L.ptr(state), L.stream_ptr())


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step: int, state: torch.Tensor | None,
               shadow: torch.Tensor | None = None, coupled_l2: bool = False) -> None:
    """shadow: optional bf16 buffer of p's size, rewritten with the updated parameters (what the tensor-core GEMMs read).
    step <= 0: bias corrections from the device-side applied-step counter state[3].  coupled_l2: torch.optim.Adam decay."""
    assert shadow is None or (shadow.dtype == torch.bfloat16 and shadow.numel() == p.numel())
    L.call("psg_adam_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), C.c_longlong(p.numel()), C.c_float(lr), C.c_double(beta1),
           C.c_double(beta2), C.c_float(eps), C.c_float(weight_decay), C.c_longlong(step), C.c_int(int(coupled_l2)), L.ptr(state),
           L.ptr(shadow), L.stream_ptr())


def cast_bf16(x: torch.Tensor, y: torch.Tensor) -> None:
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.numel() == y.numel() and x.is_contiguous() and y.is_contiguous()
    L.call("psg_cast_bf16", L.ptr(x), L.ptr(y), C.c_longlong(x.numel()), L.stream_ptr())


def scale_inplace(x: torch.Tensor, state: torch.Tensor | None, extra: float = 1.0) -> None:
    L.call("psg_scale_inplace", L.ptr(x), C.c_longlong(x.numel()), L.ptr(state), C.c_float(extra), L.stream_ptr())


def check_kernel_timeouts() -> None:
    """Raises if a bounded in-kernel wait (tcgen05 pipeline barrier, stream-K flag) expired since the last call: the kernels
    never hang the GPU on a protocol fault, they raise a device flag and fall through with incomplete results -- which must
    not go unnoticed.  Synchronises the device: call at log points, not per step."""
    lib = L.load()
    gemm, attn, gn = int(lib.psg_umma_timeout_flag()), int(lib.psg_attn_umma_timeout_flag()), int(lib.psg_groupnorm_timeout_flag())
    if gemm or attn or gn:
        raise L.PsgError(f"a bounded in-kernel wait expired (tcgen05 gemm={gemm}, attention={attn}, streaming GroupNorm backward={gn}): "
                         "results since the last check are not trustworthy")
