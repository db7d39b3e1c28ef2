"""Host-side description of GEMM / implicit-GEMM convolution problems for the two CUDA engines.

`Operand` wraps a torch tensor view as one of the access modes of csrc/gemm_desc.h; `run_gemm` fills a
PsgGemmDesc and dispatches to the tcgen05 engine (bf16) or the SIMT engine (fp32 parity mode, odd shapes).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L


@dataclass
class Operand:
    t: torch.Tensor
    mode: int
    ld: int
    rows: int          # logical rows of the [rows x K] operand
    k: int             # logical K
    conv: tuple = (0, 0, 0, 0, 0, 0, 1, 0, 1, 0)   # n,h,w,c,p,q,stride,pad,ksize,flip

    def fill(self, o: L.PsgOperand) -> None:
        o.ptr = self.t.data_ptr()
        o.mode = self.mode
        o.ld = self.ld
        (o.n, o.h, o.w, o.c, o.p, o.q, o.stride, o.pad, o.ksize, o.flip) = self.conv


def kmajor(t: torch.Tensor) -> Operand:
    """[rows, K] matrix with K contiguous (row pitch arbitrary)."""
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return Operand(t, L.OP_KMAJOR, t.stride(0), t.shape[0], t.shape[1])


def mnmajor(t: torch.Tensor) -> Operand:
    """Memory is [K, rows] with rows contiguous; logical operand is its transpose."""
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return Operand(t, L.OP_MNMAJOR, t.stride(0), t.shape[1], t.shape[0])


def _nhwc_geom(x: torch.Tensor):
    assert x.dim() == 4 and x.stride(3) == 1, (x.shape, x.stride())
    n, h, w, c = x.shape
    ld = x.stride(2)
    assert x.stride(1) == ld * w and (n == 1 or x.stride(0) == ld * w * h), "NHWC view must be pixel-pitched"
    return n, h, w, c, ld


def im2col(x: torch.Tensor, ksize: int, stride: int, pad: int, flip: bool = False, pad_hi: Optional[int] = None) -> Operand:
    """rows = conv output pixels (n,p,q); K = ksize^2 * C over NHWC `x`.  pad_hi: padding of the bottom / right border when it
    differs from `pad` (tcgen05 engine only; carried in bits 8..15 of the operand's `flip` field, see csrc/gemm_desc.h)."""
    n, h, w, c, ld = _nhwc_geom(x)
    hi = pad if pad_hi is None else pad_hi
    p = (h + pad + hi - ksize) // stride + 1
    q = (w + pad + hi - ksize) // stride + 1
    fl = int(flip) | ((hi + 1) << 8 if pad_hi is not None else 0)
    return Operand(x, L.OP_IM2COL, ld, n * p * q, ksize * ksize * c, (n, h, w, c, p, q, stride, pad, ksize, fl))


def im2col_t(x: torch.Tensor, ksize: int, stride: int, pad: int) -> Operand:
    """rows = (tap, c); K = conv output pixels -- the wgrad "B" operand."""
    n, h, w, c, ld = _nhwc_geom(x)
    p = (h + 2 * pad - ksize) // stride + 1
    q = (w + 2 * pad - ksize) // stride + 1
    return Operand(x, L.OP_IM2COL_T, ld, ksize * ksize * c, n * p * q, (n, h, w, c, p, q, stride, pad, ksize, 0))


def convw_t(w2: torch.Tensor, cin: int, ksize: int) -> Operand:
    """Conv weight stored [Cout, ksize^2 * Cin] (tap-major, channels innermost) read transposed: rows = cin,
    K = (tap, cout) -- the dgrad "B" operand, no transposed copy needed (tcgen05 engine only)."""
    assert w2.dim() == 2 and w2.stride(1) == 1 and w2.shape[1] == ksize * ksize * cin, (w2.shape, cin, ksize)
    cout = w2.shape[0]
    return Operand(w2, L.OP_CONVW_T, w2.stride(0), cin, ksize * ksize * cout, (cout, 0, 0, cin, 0, 0, 1, 0, ksize, 0))


def dgrad_gather(dy: torch.Tensor, in_h: int, in_w: int, ksize: int, stride: int, pad: int) -> Operand:
    """rows = conv input pixels; K = ksize^2 * Cout gathered from NHWC `dy` (general stride; SIMT engine only)."""
    n, h, w, c, ld = _nhwc_geom(dy)
    return Operand(dy, L.OP_DGRAD, ld, n * in_h * in_w, ksize * ksize * c, (n, h, w, c, in_h, in_w, stride, pad, ksize, 0))


@dataclass
class Epilogue:
    out: torch.Tensor                      # 2D view [M, >=N], unit column stride
    bias: Optional[torch.Tensor] = None    # fp32 [N]
    rowbias: Optional[torch.Tensor] = None  # fp32 [groups, N]
    rows_per_group: int = 1
    act: int = L.ACT_NONE
    alpha: float = 1.0
    residual: Optional[torch.Tensor] = None
    aux_out: Optional[torch.Tensor] = None
    aux_in: Optional[torch.Tensor] = None
    aux_act: int = L.ACT_NONE
    accumulate: bool = False
    drop_seed: int = 0
    drop_p: float = 0.0

    def fill(self, e: L.PsgEpilogue) -> None:
        out = self.out
        assert out.stride(-1) == 1
        e.out = out.data_ptr()
        e.ldc = out.stride(0) if out.dim() == 2 else out.stride(-2)
        e.out_dtype = L.dt(out)
        act_t = self.residual if self.residual is not None else (self.aux_out if self.aux_out is not None else self.aux_in)
        e.act_dtype = L.dt(act_t) if act_t is not None else L.dt(out)
        e.bias = self.bias.data_ptr() if self.bias is not None else None
        if self.bias is not None:
            assert self.bias.dtype == torch.float32 and self.bias.is_contiguous()
        if self.rowbias is not None:
            assert self.rowbias.dtype == torch.float32 and self.rowbias.stride(1) == 1
            e.rowbias = self.rowbias.data_ptr()
            e.ld_rowbias = self.rowbias.stride(0)
            e.rows_per_group = self.rows_per_group
        else:
            e.rowbias = None
            e.rows_per_group = 1
            e.ld_rowbias = 0
        e.act = self.act
        e.alpha = self.alpha
        if self.residual is not None:
            assert self.residual.stride(1) == 1
            e.residual = self.residual.data_ptr()
            e.ldr = self.residual.stride(0)
        else:
            e.residual = None
            e.ldr = 0
        aux = self.aux_out if self.aux_out is not None else self.aux_in
        e.aux_out = self.aux_out.data_ptr() if self.aux_out is not None else None
        e.aux_in = self.aux_in.data_ptr() if self.aux_in is not None else None
        e.ld_aux = aux.stride(0) if aux is not None else 0
        if self.aux_out is not None and self.aux_in is not None:
            assert self.aux_out.stride(0) == self.aux_in.stride(0)
        e.aux_act = self.aux_act
        e.accumulate = int(self.accumulate)
        e.drop_seed = self.drop_seed
        if self.drop_p > 0.0:
            e.drop_threshold = min(int(self.drop_p * 4294967296.0), 4294967295)
            e.drop_scale = 1.0 / (1.0 - self.drop_p)
        else:
            e.drop_threshold = 0
            e.drop_scale = 1.0


def _epi_signature(self) -> str:
    """what the epilogue fuses, for the per-shape profile (tools/profile_step.py --shapes)"""
    parts = []
    if self.bias is not None:
        parts.append("bias")
    if getattr(self, "rowbias", None) is not None:
        parts.append("rowbias")
    if self.act != L.ACT_NONE:
        parts.append({L.ACT_GELU: "gelu", L.ACT_SILU: "silu"}.get(self.act, f"act{self.act}"))
    if self.aux_out is not None:
        parts.append("aux_out")
    if self.aux_in is not None:
        parts.append("aux_in")
    if self.drop_p > 0.0:
        parts.append("drop")
    if self.residual is not None:
        parts.append("res")
    if self.accumulate:
        parts.append("acc")
    parts.append("f32" if self.out.dtype == torch.float32 else "bf16")
    return "+".join(parts)


Epilogue.signature = _epi_signature


def plan(a: Operand, b: Operand, block_n: int = 0, m_tiles: int = 0):
    """Tile shape the tcgen05 engine will use for this problem: (block_n, m_tiles, number of output tiles)."""
    d = L.PsgGemmDesc()
    a.fill(d.a)
    b.fill(d.b)
    d.M, d.N, d.K = a.rows, b.rows, a.k
    bn, mt = C.c_int(block_n), C.c_int(m_tiles)
    L.check(L.load().psg_umma_plan(C.byref(d), C.byref(bn), C.byref(mt)), "psg_umma_plan")
    bn, mt = bn.value, mt.value
    mtiles = -(-a.rows // (128 * mt))
    ntiles = -(-b.rows // bn)
    return bn, mt, mtiles * ntiles


_umma_ws = {}


def _ensure_umma_workspace(device) -> None:
    """Registers the stream-K partial-accumulator workspace of the tcgen05 engine (once per process/device)."""
    if device in _umma_ws:
        return
    lib = L.load()
    lib.psg_umma_workspace_bytes.restype = C.c_size_t
    n = lib.psg_umma_workspace_bytes()
    ws = torch.zeros(n, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):      # the library keys the registration by the calling thread's current device
        L.check(lib.psg_umma_set_workspace(C.c_void_p(ws.data_ptr()), C.c_size_t(n)), "psg_umma_set_workspace")
    _umma_ws[device] = ws


# Stream-K workspace lane of the launches that follow (0 or 1): GEMMs in flight together on two streams use different lanes.
LANE = 0

# When set to a list, every launch appends (start_event, end_event, algorithmic_flops, engine): bench.py uses it to time
# the tensor-core kernel on its own stream inside the timed region (roofline numerator and denominator).
PROFILE = None


def run_gemm(a: Operand, b: Operand, epi: Epilogue, *, engine: str = "auto", split_k: int = 1, block_n: int = 0,
             m_tiles: int = 0, algo_flops: Optional[float] = None) -> None:
    """C[M,N] = epilogue(sum_k A(m,k) B(n,k)).  engine: 'umma' (tcgen05, bf16), 'simt', or 'auto'.
    algo_flops: algorithmic FLOPs of this launch when they differ from 2*M*N*K (zero-inserted stride-2 dgrad)."""
    assert a.k == b.k, f"K mismatch {a.k} vs {b.k}"
    assert a.t.dtype == b.t.dtype, "operand dtypes differ"
    d = L.PsgGemmDesc()
    a.fill(d.a)
    b.fill(d.b)
    d.M, d.N, d.K = a.rows, b.rows, a.k
    d.in_dtype = L.dt(a.t)
    d.split_k = split_k
    epi.fill(d.epi)
    if engine == "auto":
        engine = "umma" if a.t.dtype == torch.bfloat16 else "simt"
    lib = L.load()
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if engine == "umma":
        _ensure_umma_workspace(a.t.device)
        L.check(lib.psg_umma_gemm_lane(C.byref(d), C.c_int(block_n), C.c_int(m_tiles), C.c_int(LANE), L.stream_ptr()), "psg_umma_gemm")
    elif engine == "simt":
        L.check(lib.psg_simt_gemm(C.byref(d), L.stream_ptr()), "psg_simt_gemm")
    else:
        raise ValueError(engine)
    if prof is not None:
        ev1.record()
        prof.append((ev0, ev1, float(algo_flops) if algo_flops is not None else 2.0 * a.rows * b.rows * a.k, engine,
                     (a.mode, b.mode, a.rows, b.rows, a.k, split_k), epi.signature()))
