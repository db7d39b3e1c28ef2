"""Execution engine of the U-Net: an explicit forward / backward schedule of C-ABI kernel launches.

This is the host side of the hot path: it walks the reference's graph (UNet.forward,
src/models/unet.py:428-509; ResBlock.forward :112-132; CrossAttentionBlock.forward :206-260) and issues one kernel
per fused op on the current CUDA stream.  Activations are token-major ("NHWC") in the compute dtype; channel
concatenation is a strided view of a wider buffer, never a copy of both halves.  A reverse-mode tape of closures
gives the backward pass; parameter gradients land in one flat fp32 buffer laid out like the flat parameter buffer
(so the gradient all-reduce, the global-norm clip and fused AdamW each are a single pass over contiguous memory).

No op here is computed by PyTorch: torch provides memory (torch.empty), streams and autograd glue only.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import Callable, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import gemm as G
from . import ops as K
from .unet import LEVELS

# gradient buckets: "main" (default) joins the weight stream into the main stream before a bucket's all-reduce; "side" issues the
# all-reduce with the weight stream current instead (no stall of the dX chain).  Measured equal at 2 GPUs (86.2 vs 86.1-86.6 ms,
# profiles/r02_bench_bucket_issue_ab.txt), so the simpler, longer-validated ordering stays the default.
_BUCKET_ISSUE_MAIN = os.environ.get("PSG_BUCKET_ISSUE", "main") == "main"
_WGRAD_STREAM = os.environ.get("PSG_WGRAD_STREAM", "1") != "0"  # A/B switch: 0 = weight gradients on the main stream
_DGRAD_S2 = os.environ.get("PSG_DGRAD_S2", "1") != "0"     # A/B switch: 0 = zero-inserted stride-1 dgrad for the downsample convs

NUM_SMS = 148
ALIGN = 64  # elements; keeps every parameter 256-byte aligned inside the flat buffers


class Act:
    """A token-major activation [M = B*H*W, C] (possibly a channel slice of a wider buffer) plus its gradient."""

    __slots__ = ("t", "B", "H", "W", "grad", "parent", "col", "pre", "pre_act", "drop", "sink", "colsum_done")

    def __init__(self, t: torch.Tensor, B: int, H: int, W: int, parent: "Act" = None, col: int = 0):
        self.t, self.B, self.H, self.W = t, B, H, W
        self.grad: Optional[torch.Tensor] = None
        self.parent, self.col = parent, col
        self.pre: Optional[torch.Tensor] = None   # when this is drop(act(pre)): the saved local derivative act'(pre) * dropmask
        self.pre_act = L.ACT_NONE
        self.drop = None                          # (unused: the dropout mask is folded into `pre`)
        self.sink = None                          # (rowbias Act, bias grad) fed by the column sums of this conv output's gradient
        self.colsum_done = False                  # ... already produced by the consumer's fused GroupNorm backward

    @property
    def M(self) -> int:
        return self.t.shape[0]

    @property
    def C(self) -> int:
        return self.t.shape[1]

    def slice(self, col: int, width: int) -> "Act":
        return Act(self.t[:, col:col + width], self.B, self.H, self.W, parent=self, col=col)

    def g(self) -> torch.Tensor:
        if self.parent is not None:
            return self.parent.grad[:, self.col:self.col + self.C]
        return self.grad

    def nhwc(self) -> torch.Tensor:
        ld = self.t.stride(0)
        return self.t.as_strided((self.B, self.H, self.W, self.C), (self.H * self.W * ld, self.W * ld, ld, 1))

    @staticmethod
    def nhwc_of(t: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        ld = t.stride(0)
        return t.as_strided((B, H, W, t.shape[1]), (H * W * ld, W * ld, ld, 1))


class ConvW:
    def __init__(self, conv: nn.Conv2d, pad_channels: bool):
        self.mod = conv
        self.cout, self.cin, self.k = conv.out_channels, conv.in_channels, conv.kernel_size[0]
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        self.wp = self.wd = self.bias_p = None
        # Edge layers (init_conv Cin=8, final_conv Cout=8) do not fill a 64-wide tensor-core tile.  In bf16 mode their
        # channels are zero-padded to 64 in the packed weights / activations so they run on the tcgen05 engine too
        # (8x the FLOPs of a 0.04%-of-the-model layer, instead of a 100x slower CUDA-core GEMM).
        self.cin_p = (self.cin + 63) // 64 * 64 if pad_channels else self.cin
        self.cout_p = (self.cout + 63) // 64 * 64 if pad_channels else self.cout
        self.padded = (self.cin_p != self.cin) or (self.cout_p != self.cout)
        # tensor-core mode, tileable channels: the weight lives tap-major in the flat buffers and the GEMMs read its bf16
        # shadow in place (fprop: K-major [Cout, taps*Cin]; dgrad: the same matrix read transposed; wgrad writes the
        # gradient matrix directly) -- no packing pass, no transposed copy, no finalize pass
        self.in_place = pad_channels and not self.padded and self.k > 1


class LinW:
    """A (row-slice of a) Linear weight [N, K] with optional bias; w/b/gw/gb are views of the flat buffers."""

    def __init__(self, weight: nn.Parameter, bias: Optional[nn.Parameter], rows: Optional[tuple] = None):
        self.weight, self.bias, self.rows = weight, bias, rows
        self.wk = self.wt = None

    def views(self, store: "ParamStore"):
        w, gw = self.weight.data, store.grad_of(self.weight)
        b = self.bias.data if self.bias is not None else None
        gb = store.grad_of(self.bias) if self.bias is not None else None
        if self.rows is not None:
            a, e = self.rows
            w, gw = w[a:e], gw[a:e]
            if b is not None:
                b, gb = b[a:e], gb[a:e]
        return w, b, gw, gb


class FlatLinW(LinW):
    """A Linear whose weight [N, K] (and bias [N]) is a contiguous run of several parameters inside the flat buffers:
    the 17 ResBlock time_proj (resp. text_proj) layers evaluated as ONE GEMM (reference unet.py:83,86,119-124)."""

    def __init__(self, first_w: nn.Parameter, first_b: nn.Parameter, n: int, k: int):
        super().__init__(first_w, first_b)
        self.n, self.k = n, k

    def views(self, store: "ParamStore"):
        ow, ob = store.offset_of(self.weight), store.offset_of(self.bias)
        n, k = self.n, self.k
        store.touch(ow, n * k)
        store.touch(ob, n)
        return (store.flat[ow:ow + n * k].view(n, k), store.flat[ob:ob + n], store.grads[ow:ow + n * k].view(n, k),
                store.grads[ob:ob + n])


class NormW:
    def __init__(self, gn: nn.GroupNorm):
        self.mod, self.groups, self.eps = gn, gn.num_groups, gn.eps


class ParamStore:
    """Flat fp32 parameter / gradient buffers; every nn.Parameter's .data is a view into `flat`."""

    def __init__(self, module: nn.Module, front: Optional[List[nn.Parameter]] = None,
                 tap_major: Optional[List[nn.Parameter]] = None, shadow: bool = False):
        """`front`: parameters laid out first, contiguously and in the given order (the ResBlock conditioning
        projections, so that all time_proj / text_proj weights form one [sum C, K] matrix each); every other parameter
        follows in registration order.
        `tap_major`: conv weights stored physically as [Cout][kh][kw][Cin] (channels-last memory format): `.data` keeps
        the logical OIHW shape through a permuted view, so state_dict / optimisers / autograd see the reference layout
        while the implicit-GEMM kernels read (and write gradients to) the buffer as a [Cout, taps*Cin] matrix in place.
        `shadow`: keep a bf16 copy of the flat buffer (same offsets) for the tensor-core GEMMs."""
        self.module = module
        self.tap_major = {id(p) for p in (tap_major or [])}
        self.want_shadow = shadow
        self.shadow: Optional[torch.Tensor] = None
        named = list(module.named_parameters())
        name_of = {id(p): n for n, p in named}
        front = list(front or [])
        front_ids = {id(p) for p in front}
        self.named = [(name_of[id(p)], p) for p in front] + [(n, p) for n, p in named if id(p) not in front_ids]
        self.offsets = {}
        off = 0
        for name, p in self.named:
            self.offsets[name] = off
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self._name_of = name_of
        self.flat: Optional[torch.Tensor] = None
        self.grads: Optional[torch.Tensor] = None
        self._by_id = {}
        self.generation = 0
        self.touch_log: Optional[list] = None

    def ensure_flat(self, device) -> bool:
        """(Re)builds the flat buffer if any parameter was moved/replaced.  Returns True when rebuilt."""
        ok = self.flat is not None and self.flat.device == device
        if ok:
            base = self.flat.data_ptr()
            for name, p in (self.named[0], self.named[-1], self.named[len(self.named) // 2]):
                if p.data_ptr() != base + 4 * self.offsets[name] or p.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            return False
        flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for name, p in self.named:
                view = self._view(flat, p, self.offsets[name])
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
        self.flat = flat
        self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=device) if self.want_shadow else None
        self.grads = torch.zeros(self.total, dtype=torch.float32, device=device)
        self._index_grads()
        self.generation += 1
        return True

    def _view(self, buf: torch.Tensor, p: nn.Parameter, off: int) -> torch.Tensor:
        """The view of `buf` that is parameter p (logical shape; tap-major conv weights through a permute)."""
        n = p.numel()
        if id(p) in self.tap_major:
            co, ci, kh, kw = p.shape
            return buf[off:off + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return buf[off:off + n].view(p.shape)

    def matrix(self, buf: torch.Tensor, p: nn.Parameter) -> torch.Tensor:
        """Physical [rows, cols] matrix of p inside `buf` (flat / grads / shadow): [Cout, taps*Cin] for tap-major conv
        weights, [shape[0], rest] otherwise."""
        off = self.offset_of(p)
        return buf[off:off + p.numel()].view(p.shape[0], -1)

    def _index_grads(self):
        self._by_id = {}
        for name, p in self.named:
            self._by_id[id(p)] = self._view(self.grads, p, self.offsets[name])

    def fresh_grads(self):
        """Detach the current gradient buffer (it may be referenced by p.grad) and start a new one."""
        self.grads = torch.zeros(self.total, dtype=torch.float32, device=self.flat.device)
        self._index_grads()

    def grad_of(self, p: nn.Parameter) -> torch.Tensor:
        if self.touch_log is not None:
            self.touch_log.append((self.offsets[self._name_of[id(p)]], p.numel()))
        return self._by_id[id(p)]

    def touch(self, offset: int, numel: int) -> None:
        """Records that the gradient range [offset, offset + numel) is about to be written (backward only: the bucketed
        all-reduce in parallel.GradSync learns from this log when each bucket of the flat buffer is final)."""
        if self.touch_log is not None:
            self.touch_log.append((offset, numel))

    def offset_of(self, p: nn.Parameter) -> int:
        return self.offsets[self._name_of[id(p)]]

    def grad_views(self) -> List[torch.Tensor]:
        return [self._by_id[id(p)] for _, p in self.named]

    def version(self) -> int:
        return sum(p._version for _, p in self.named)


class UNetEngine:
    def __init__(self, unet: nn.Module, compute_dtype: torch.dtype = torch.bfloat16):
        if compute_dtype not in (torch.bfloat16, torch.float32):
            raise L.PsgError(f"compute_dtype must be bfloat16 or float32, got {compute_dtype}")
        L.load()
        self.unet = unet
        self.dtype = compute_dtype
        self.bf16 = compute_dtype == torch.bfloat16
        self._packed_version = None
        self._packed_generation = None
        self._edges_dirty = False
        self._wd_all = None
        self._wt_all = None
        self._build_descriptors()
        self.step_counter = 0
        self.dropout_enabled = True      # honoured only in train() mode
        self.cond_on_tensor_cores = True  # bf16 mode: all-blocks conditioning GEMMs on the tcgen05 engine (see _cond_all)
        # Dropout RNG contract: masks are a stateless hash of (seed, step_counter, site, element).  `seed` defaults (lazily, at
        # the first forward) to torch.initial_seed() mixed with the data-parallel rank, so torch.manual_seed(...) selects the
        # mask sequence and replicas draw different masks; `step_counter` advances once per forward and is restored from
        # `global_step` by DiffusionTrainer.load_checkpoint, so a resumed run continues the sequence instead of replaying it.
        self.seed = None
        self.launches = 0
        # backward: weight / bias gradients run on a second stream (see _weight_stream)
        self.weight_stream_enabled = _WGRAD_STREAM     # bench.py turns it off for its per-kernel (serialised) timing pass
        self.wstream = None               # the stream in use during the current backward (None: everything on the main stream)
        self._wstream_obj = None
        self._wbusy = {}                  # data_ptr of a gradient tensor the second stream reads -> event after that read

    # ------------------------------------------------------------------------------------------------------------
    # descriptors
    # ------------------------------------------------------------------------------------------------------------
    def _build_descriptors(self):
        u = self.unet
        self.convs: List[ConvW] = []
        self.lins: List[LinW] = []

        def conv(m):
            c = ConvW(m, pad_channels=self.bf16)
            self.convs.append(c)
            return c

        def lin(w, b, rows=None):
            l = LinW(w, b, rows)
            self.lins.append(l)
            return l

        res_blocks = []

        def res(rb):
            # time_proj / text_proj of all ResBlocks are evaluated up front as two GEMMs over their concatenated weights
            # (FlatLinW); each block's conditioning is a column slice [cond_off, cond_off + cout) of that result.
            d = {"norm1": NormW(rb.norm1), "conv1": conv(rb.conv1), "norm2": NormW(rb.norm2), "conv2": conv(rb.conv2),
                 "skip": None, "cin": rb.in_channels, "cout": rb.out_channels,
                 "cond_off": sum(r.out_channels for r in res_blocks)}
            res_blocks.append(rb)
            if isinstance(rb.skip_conv, nn.Conv2d):
                sk = rb.skip_conv
                d["skip"] = lin(sk.weight, sk.bias)      # 1x1 conv == linear over tokens ([Cout, Cin, 1, 1] viewed [Cout, Cin])
            return d

        def attn(ab):
            c = ab.channels
            sa, ca = ab.self_attn, ab.cross_attn
            return {"c": c, "norm1": NormW(ab.norm1), "norm2": NormW(ab.norm2),
                    "qkv": lin(sa.in_proj_weight, sa.in_proj_bias), "so": lin(sa.out_proj.weight, sa.out_proj.bias),
                    "cq": lin(ca.in_proj_weight, ca.in_proj_bias, (0, c)), "ckv": lin(ca.in_proj_weight, ca.in_proj_bias, (c, 3 * c)),
                    "co": lin(ca.out_proj.weight, ca.out_proj.bias), "tp": lin(ab.text_proj.weight, ab.text_proj.bias),
                    "f1": lin(ab.ffn[0].weight, ab.ffn[0].bias), "f2": lin(ab.ffn[3].weight, ab.ffn[3].bias),
                    "p_attn": ab.ATTN_DROPOUT, "p_ffn": ab.FFN_DROPOUT}

        def block(b):
            return {"res": res(b.res_block), "attn": attn(b.attn_block) if b.has_attention else None}

        te = u.time_embed.time_mlp
        self.d_time = [lin(te[0].weight, te[0].bias), lin(te[2].weight, te[2].bias), lin(te[4].weight, te[4].bias)]
        self.d_init = conv(u.init_conv)
        self.d_enc = [[block(b) for b in getattr(u, f"enc_block{l}")] for l in range(4)]
        self.d_down = [None] + [conv(getattr(u, f"downsample{l}")) for l in (1, 2, 3)]
        self.d_mid = block(u.middle_block)
        self.d_dec = {l: [block(b) for b in getattr(u, f"dec_block{l}")] for l in (3, 2, 1, 0)}
        self.d_up = {l: conv(getattr(u, f"upsample{l}")[1]) for l in (3, 2, 1)}
        self.d_final_norm = NormW(u.final_conv[0])
        self.d_final = conv(u.final_conv[2])
        # flat layout: [time_proj weights | text_proj weights | time_proj biases | text_proj biases | everything else]
        front = ([rb.time_proj.weight for rb in res_blocks] + [rb.text_proj.weight for rb in res_blocks] +
                 [rb.time_proj.bias for rb in res_blocks] + [rb.text_proj.bias for rb in res_blocks])
        for p in front:
            assert p.numel() % ALIGN == 0, "conditioning projections must tile the flat buffer without padding"
        self.store = ParamStore(u, front=front, tap_major=[c.mod.weight for c in self.convs if c.in_place], shadow=self.bf16)
        self.cond_width = sum(rb.out_channels for rb in res_blocks)
        rb0 = res_blocks[0]
        self.d_cond_time = FlatLinW(rb0.time_proj.weight, rb0.time_proj.bias, self.cond_width, rb0.time_proj.in_features)
        self.d_cond_text = FlatLinW(rb0.text_proj.weight, rb0.text_proj.bias, self.cond_width, rb0.text_proj.in_features)

    # ------------------------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------------------------
    def prepare(self, device) -> None:
        """Flatten parameters if needed and refresh the kernel-side weight copies if they changed.
        bf16 mode: one cast pass flat fp32 -> bf16 shadow (skipped when the fused AdamW already wrote it) plus a re-pack
        of the two channel-padded edge convs; fp32 parity mode: re-pack every conv (fprop / dgrad layouts)."""
        store = self.store
        rebuilt = store.ensure_flat(device)
        ver = store.version()
        stale = rebuilt or self._packed_version != ver or self._packed_generation != store.generation
        if not stale and not self._edges_dirty:
            return
        if self.bf16 and stale:
            K.cast_bf16(store.flat, store.shadow)
        if self.bf16:
            self._refresh_dgrad_weights(device, stale)
        for c in self.convs:
            if c.in_place:
                if stale:
                    c.wp = store.matrix(store.shadow, c.mod.weight)
                continue
            w = c.mod.weight.data
            if c.wp is None or c.wp.device != device:
                c.wp = torch.zeros(c.cout_p, c.k * c.k * c.cin_p, dtype=self.dtype, device=device)
                c.wd = torch.zeros(c.cin_p, c.k * c.k * c.cout_p, dtype=self.dtype, device=device)
                c.bias_p = torch.zeros(1, c.cout_p, dtype=torch.float32, device=device) if c.padded else None
            K.pack_conv_weight(w, c.wp, c.wd, c.cin_p, c.cout_p)
            if c.bias_p is not None:
                K.copy_strided(c.mod.bias.data.view(1, c.cout), c.bias_p[:, :c.cout])
        if self.bf16 and stale:
            for l in self.lins:
                wk = store.matrix(store.shadow, l.weight)
                if l.rows is not None:
                    a, e = l.rows
                    wk = wk[a:e]
                l.wk = wk
            for fl in (self.d_cond_time, self.d_cond_text):
                ow = store.offset_of(fl.weight)
                fl.wk = store.shadow[ow:ow + fl.n * fl.k].view(fl.n, fl.k)
        self._edges_dirty = False
        self._packed_version = store.version()
        self._packed_generation = store.generation

    def _refresh_dgrad_weights(self, device, rebuild_views: bool) -> None:
        """dgrad reads a K-major weight like fprop does: all in-place conv weights [Cout][tap][Cin] (bf16 shadow) are
        transposed to [Cin][tap][Cout] in ONE launch per parameter update (0.4 ms; reading the fprop matrix transposed in
        place through an MN-major UMMA operand was measured 10-20 % slower per dgrad GEMM, 2.5 ms per step)."""
        convs = [c for c in self.convs if c.in_place]
        if not convs:
            return
        total = sum(c.mod.weight.numel() for c in convs)
        if self._wd_all is None or self._wd_all.device != device or self._wd_all.numel() != total:
            self._wd_all = torch.empty(total, dtype=torch.bfloat16, device=device)
            rebuild_views = True
        jobs, off = [], 0
        for c in convs:
            n = c.mod.weight.numel()
            jobs.append((self.store.offset_of(c.mod.weight), off, c.cout, c.cin, c.k * c.k))
            if rebuild_views:
                c.wd = self._wd_all[off:off + n].view(c.cin, c.k * c.k * c.cout)
            off += n
        K.conv_weights_transpose(self.store.shadow, self._wd_all, jobs)
        # stride-2 convs (the three downsample layers): dgrad by output parity needs four small class weights (see _conv_bwd)
        for c in convs:
            if c.stride == 2 and c.k == 3 and c.pad == 1:
                if getattr(c, "wd_s2", None) is None or c.wd_s2[0].device != device:
                    c.wd_s2 = [torch.empty(c.cin, (1 if i == 0 else 4) * c.cout, dtype=torch.bfloat16, device=device) for i in range(4)]
                K.dgrad_s2_weights(c.wd, *c.wd_s2, c.cin, c.cout)
        # Linear weights [N, K] likewise: dgrad (dX = dY W) reads a K-major [K, N] copy instead of the MN-major in-place
        # operand (a Linear is a 1-tap conv for the transpose kernel; row slices of a packed in_proj are column slices here)
        weights = {}
        for l in self.lins:
            if isinstance(l, FlatLinW):
                continue
            n, k = l.weight.shape[0], l.weight.numel() // l.weight.shape[0]
            if n % 64 == 0 and k % 64 == 0:
                weights.setdefault(id(l.weight), l.weight)
        total = sum(w.numel() for w in weights.values())
        if total == 0:
            return
        if self._wt_all is None or self._wt_all.device != device or self._wt_all.numel() != total:
            self._wt_all = torch.empty(total, dtype=torch.bfloat16, device=device)
            rebuild_views = True
        jobs, off, views = [], 0, {}
        for wid, w in weights.items():
            n, k = w.shape[0], w.numel() // w.shape[0]
            jobs.append((self.store.offset_of(w), off, n, k, 1))
            views[wid] = self._wt_all[off:off + n * k].view(k, n)
            off += n * k
        if rebuild_views:
            for l in self.lins:
                full = views.get(id(l.weight)) if not isinstance(l, FlatLinW) else None
                l.wt = None if full is None else (full if l.rows is None else full[:, l.rows[0]:l.rows[1]])
        for i in range(0, len(jobs), 64):
            K.conv_weights_transpose(self.store.shadow, self._wt_all, jobs[i:i + 64])

    def mark_params_dirty(self) -> None:
        """Call after updating the flat parameter buffer outside of torch: forces a full refresh of the kernel-side copies."""
        self._packed_version = None

    def mark_shadow_fresh(self) -> None:
        """Call after a fused optimiser step that rewrote the bf16 shadow itself: only the edge convs need re-packing."""
        self._edges_dirty = True

    # ------------------------------------------------------------------------------------------------------------
    # primitive ops (forward + tape entry)
    # ------------------------------------------------------------------------------------------------------------
    def _new(self, M, C_, B, H, W, dtype=None) -> Act:
        return Act(torch.empty(M, C_, dtype=dtype or self.dtype, device=self.device), B, H, W)

    def _grad_target(self, a: Act):
        """Returns (tensor, accumulate) for adding a gradient contribution to `a`."""
        assert a.parent is None, "gradients of slices are produced through their parent buffer"
        if a.grad is None:
            a.grad = torch.empty(a.M, a.C, dtype=a.t.dtype, device=a.t.device)
            return a.grad, False
        if self._wbusy:
            # a residual's gradient may BE the dY buffer of a layer whose weight gradient is still being formed on the second
            # stream (_pass_grad hands the buffer over): the in-place accumulation that follows waits for that read
            ev = self._wbusy.pop(a.grad.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        return a.grad, True

    @contextlib.contextmanager
    def _weight_stream(self, dy: torch.Tensor, *reads: torch.Tensor):
        """Backward's weight / bias gradients (wgrad GEMMs, bias column sums) are needed only by the optimizer (or the gradient
        all-reduce), not by the dX chain, so they are enqueued on a second stream: their CTAs fill the SMs the critical-path
        kernels leave idle (persistent-grid tails, stream-K imbalance, the small 4x4-level GEMMs) and the column sums co-reside
        with the GEMM CTAs.  The block runs after everything enqueued so far on the main stream; `dy` and `reads` are the
        tensors it reads (kept alive for the second stream; an in-place accumulation into `dy` later on the main stream waits
        for the read: _grad_target).  backward() joins the streams at the end, GradSync before each bucket goes out."""
        side = self.wstream
        if side is None:
            yield
            return
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        for t in (dy,) + reads:
            t.record_stream(side)
        G.LANE = 1                      # the second stream-K workspace: these GEMMs overlap the main stream's
        try:
            with torch.cuda.stream(side):
                yield
        finally:
            G.LANE = 0
        ev = torch.cuda.Event()
        ev.record(side)
        self._wbusy[dy.data_ptr()] = ev

    @contextlib.contextmanager
    def _bucket_issue_ctx(self):
        """The weight-gradient stream, caught up with the main stream, as the current stream: what is enqueued inside is ordered
        after everything launched so far on either stream (GradSync issues a bucket's all-reduce here)."""
        side = self.wstream
        if side is None:
            yield
            return
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            yield

    def _join_weight_stream(self) -> None:
        if self.wstream is not None:
            torch.cuda.current_stream().wait_stream(self.wstream)
            self._wbusy.clear()

    def _pass_grad(self, dy: torch.Tensor, a: Act) -> None:
        """grad(a) += dy for a residual connection (out = f(..) + a).  The first contribution to a stand-alone activation
        takes dy's buffer itself instead of a copy: dy is dead once its producer's backward entry has been enqueued, and
        later contributions accumulate into it in place, in stream order."""
        if a.parent is None and a.grad is None and dy.is_contiguous() and dy.shape == a.t.shape and dy.dtype == a.t.dtype:
            a.grad = dy
            return
        tgt, acc = self._grad_target(a)
        K.copy_strided(dy, tgt, accumulate=acc)

    def _engine_for(self, t: torch.Tensor, edge: bool = False) -> str:
        return "umma" if (t.dtype == torch.bfloat16 and not edge) else "simt"

    def _seed(self, site: int) -> int:
        if self.seed is None:
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()
            except Exception:  # pragma: no cover
                rank = 0
            self.seed = (torch.initial_seed() ^ (rank * 0x9E3779B97F4A7C15) ^ 0x5EED) & 0xFFFFFFFFFFFFFFFF
        return ((self.seed * 0x9E3779B1 + self.step_counter) * 0x85EBCA77 + site * 0xC2B2AE3D) & 0xFFFFFFFFFFFFFFFF

    # ---- convolution -------------------------------------------------------------------------------------------
    def conv(self, x: Act, cw: ConvW, *, out: Act = None, rowbias: Act = None, residual: Act = None, x_needs_grad=True) -> Act:
        P = (x.H + 2 * cw.pad - cw.k) // cw.stride + 1
        Q = (x.W + 2 * cw.pad - cw.k) // cw.stride + 1
        if out is None:
            out = self._new(x.B * P * Q, cw.cout_p, x.B, P, Q)
        assert x.C == cw.cin_p and out.C == cw.cout_p, (x.C, cw.cin_p, out.C, cw.cout_p)
        eng = self._engine_for(x.t)
        bias = cw.bias_p[0] if cw.bias_p is not None else cw.mod.bias.data
        a = G.im2col(x.nhwc(), cw.k, cw.stride, cw.pad)
        epi = G.Epilogue(out=out.t, bias=bias, rowbias=rowbias.t if rowbias is not None else None, rows_per_group=P * Q,
                         residual=residual.t if residual is not None else None)
        G.run_gemm(a, G.kmajor(cw.wp), epi, engine=eng)
        if self.taping:
            if rowbias is not None and cw.cout_p == cw.cout:
                out.sink = (rowbias, cw)
            self.tape.append(lambda: self._conv_bwd(x, cw, out, rowbias, residual, x_needs_grad, eng))
        return out

    def _rowbias_target(self, rowbias: Act):
        if rowbias.parent is not None:      # a column slice of the all-blocks conditioning buffer: written once
            par = rowbias.parent
            if par.grad is None:
                par.grad = torch.empty(par.M, par.C, dtype=par.t.dtype, device=par.t.device)
            return rowbias.g(), False
        return self._grad_target(rowbias)

    def _conv_bwd(self, x: Act, cw: ConvW, out: Act, rowbias: Act, residual: Act, x_needs_grad: bool, eng: str):
        dy = out.g()
        gb = self.store.grad_of(cw.mod.bias)
        dy_real = dy[:, :cw.cout] if cw.cout_p != cw.cout else dy
        plain_colsum = False
        if out.colsum_done:
            pass        # bias / conditioning gradients came out of the consumer GroupNorm's backward
        elif rowbias is not None:
            tgt, acc = self._rowbias_target(rowbias)
            K.colsum(dy_real, x.B, tgt, gb, acc_groups=acc)
        else:
            plain_colsum = True
        if residual is not None:
            self._pass_grad(dy, residual)
        # wgrad: dW[co][tap][ci] = sum_pix dY[pix, co] * im2col(X)[pix, (tap, ci)]
        gw = self.store.grad_of(cw.mod.weight)
        kk = cw.k * cw.k
        ncols = kk * cw.cin_p
        with self._weight_stream(dy, x.t) if eng == "umma" else contextlib.nullcontext():
            if plain_colsum:
                K.colsum(dy_real, 1, None, gb)
            a_op, b_op = G.mnmajor(dy), G.im2col_t(x.nhwc(), cw.k, cw.stride, cw.pad)
            if cw.in_place:     # the gradient buffer IS the [Cout, taps*Cin] matrix the GEMM produces (stream-K: no split pass)
                G.run_gemm(a_op, b_op, G.Epilogue(out=self.store.matrix(self.store.grads, cw.mod.weight)), engine=eng)
            else:
                nel = cw.cout_p * ncols
                tag = "wgrad2" if G.LANE else "wgrad"      # the second stream's own scratch
                part = K.workspace(self.device, nel, tag).narrow(0, 0, nel).view(1, cw.cout_p, ncols)
                G.run_gemm(a_op, b_op, G.Epilogue(out=part[0]), engine=eng)
                K.wgrad_finalize(part, 1, nel, gw, cin_p=cw.cin_p)
        if not x_needs_grad:
            return
        # dgrad
        tgt, acc = self._grad_target(x) if x.parent is None else (x.g(), False)
        res = tgt if acc else None
        dy4 = Act.nhwc_of(dy, out.B, out.H, out.W)
        wd_op = G.kmajor(cw.wd)      # (G.convw_t(cw.wp, ...) reads the fprop matrix transposed in place: same result, slower MMA)
        if cw.stride == 1:
            a = G.im2col(dy4, cw.k, 1, cw.pad, flip=True)
            self._dgrad_gemm(a, wd_op, tgt, res, x, eng)
        elif eng == "simt":
            a = G.dgrad_gather(dy4, x.H, x.W, cw.k, cw.stride, cw.pad)
            self._dgrad_gemm(a, G.kmajor(cw.wd), tgt, res, x, eng)
        elif getattr(cw, "wd_s2", None) is not None and x.pre is None and _DGRAD_S2:
            # stride-2 dgrad by output parity: dX[2a+pi, 2b+pj] only receives the taps of matching parity, so each of the four classes
            # is a 2x2 (1x1 for even/even) stride-1 convolution over dY whose window may run one row / column past the end
            # (pad_hi = 1: zero-filled by TMA); 13 taps over the 14x14 grid instead of the 36 a zero-inserted convolution over the
            # 27x27 grid executes (2.6x fewer MMA FLOPs), then one interleave pass writes (or accumulates into) dX
            P, Q = out.H, out.W
            cls = []
            for i, wc in enumerate(cw.wd_s2):
                ci = torch.empty(x.B * P * Q, cw.cin_p, dtype=dy.dtype, device=dy.device)
                a = G.kmajor(dy) if i == 0 else G.im2col(dy4, 2, 1, 0, pad_hi=1)
                G.run_gemm(a, G.kmajor(wc), G.Epilogue(out=ci), engine=eng,
                           algo_flops=(2.0 * dy.shape[0] * cw.cin_p * cw.k * cw.k * cw.cout_p) if i == 0 else 0.0)
                cls.append(ci)
            K.interleave2x2(cls[0], cls[1], cls[2], cls[3], tgt, x.B, P, Q, x.H, x.W, acc)
        else:
            # stride-2 dgrad on the tensor-core engine: zero-insert dY to the input grid, then a stride-1 flipped conv
            dil = torch.empty(x.B * x.H * x.W, cw.cout_p, dtype=dy.dtype, device=dy.device)
            K.dilate2(dy, dil, x.B, out.H, out.W, x.H, x.W)
            a = G.im2col(Act.nhwc_of(dil, x.B, x.H, x.W), cw.k, 1, cw.pad, flip=True)
            self._dgrad_gemm(a, wd_op, tgt, res, x, eng, algo_flops=2.0 * dy.shape[0] * cw.cin_p * cw.k * cw.k * cw.cout_p)

    def _dgrad_gemm(self, a, b, tgt, res, x: Act, eng: str, alpha: float = 1.0, algo_flops=None):
        """tgt (+)= alpha * (A B^T) [* act'(pre) * dropmask]  -- gradient w.r.t. the pre-activation when x carries one
        (the forward epilogue saved that local derivative in x.pre)."""
        epi = G.Epilogue(out=tgt, residual=res, alpha=alpha)
        if x.pre is not None:
            epi.aux_in, epi.aux_act = x.pre, x.pre_act
        G.run_gemm(a, b, epi, engine=eng, algo_flops=algo_flops)

    # ---- linear ------------------------------------------------------------------------------------------------
    def linear(self, x: Act, lw: LinW, *, out: Act = None, act=L.ACT_NONE, alpha: float = 1.0, residual: Act = None,
               into: Act = None, drop: tuple = None, x_needs_grad: bool = True, B=None, H=None, W=None) -> Act:
        """y = alpha * drop(act(x W^T + b)) + residual.   `into`: accumulate onto an existing activation in place
        (y = into + x W^T + b), whose gradient then simply passes through."""
        w, b, gw, gb = lw.views(self.store)
        w2 = w.view(w.shape[0], -1)
        N = w2.shape[0]
        fp32 = x.t.dtype == torch.float32
        if into is not None:
            out = into
        elif out is None:
            out = self._new(x.M, N, B or x.B, H or x.H, W or x.W, dtype=x.t.dtype)
        eng = self._engine_for(x.t, edge=(N % 32 != 0))
        bmat = w2 if fp32 else lw.wk
        epi = G.Epilogue(out=out.t, bias=b, act=act, alpha=alpha)
        res_t = into.t if into is not None else (residual.t if residual is not None else None)
        epi.residual = res_t
        if act != L.ACT_NONE and self.taping:
            out.pre = torch.empty(x.M, N, dtype=x.t.dtype, device=self.device)
            out.pre_act = L.ACT_MUL
            epi.aux_out = out.pre
            epi.aux_act = act          # forward side: store act'(pre) * dropmask, so that backward is one multiply
        if drop is not None:
            epi.drop_seed, epi.drop_p = drop
        G.run_gemm(G.kmajor(x.t), G.kmajor(bmat), epi, engine=eng)
        if self.taping:
            self.tape.append(lambda: self._linear_bwd(x, lw, out, act, alpha, residual, into, drop, x_needs_grad, eng))
        return out

    def _linear_bwd(self, x: Act, lw: LinW, out: Act, act, alpha, residual: Act, into: Act, drop, x_needs_grad, eng):
        w, b, gw, gb = lw.views(self.store)
        w2, gw2 = w.view(w.shape[0], -1), gw.view(gw.shape[0], -1)
        N, Kd = w2.shape
        fp32 = x.t.dtype == torch.float32
        dy = out.g()            # for act != NONE this already is the gradient w.r.t. the pre-activation
        if residual is not None:
            self._pass_grad(dy, residual)
        scale = alpha
        if act == L.ACT_NONE and drop is not None:
            dpre = torch.empty(dy.shape, dtype=dy.dtype, device=dy.device)
            K.dropout_scale(dy, dpre, alpha, drop[0], drop[1])
            dy, scale = dpre, 1.0
        with self._weight_stream(dy, x.t) if eng == "umma" else contextlib.nullcontext():
            if gb is not None:
                K.colsum(dy, 1, None, gb, scale=scale)
            # wgrad: dW[n][k] = scale * sum_m dY[m, n] X[m, k]
            G.run_gemm(G.mnmajor(dy), G.mnmajor(x.t), G.Epilogue(out=gw2, alpha=scale), engine=eng)
        if not x_needs_grad:
            return
        tgt, acc = self._grad_target(x) if x.parent is None else (x.g(), False)
        if not fp32 and eng == "umma" and lw.wt is not None:
            bop = G.kmajor(lw.wt)                              # K-major [K, N] copy (see _refresh_dgrad_weights)
        else:
            bop = G.mnmajor(w2) if fp32 else G.mnmajor(lw.wk)      # the [N, K] weight read transposed in place
        if eng == "simt" and x.pre is None and scale == 1.0 and tgt.dtype == torch.float32 and tgt.is_contiguous() \
                and N >= 2048 and x.M * Kd <= 256 * 512:
            # skinny dgrad with a long reduction (the all-blocks conditioning projection): split-K over the CUDA-core engine
            split = min(32, N // 512)
            while True:     # fixed point of the engine's own slicing rule (16-aligned k slices)
                per = -(-(-(-N // split)) // 16) * 16
                eff = -(-N // per)
                if eff == split:
                    break
                split = eff
            part = K.workspace(self.device, split * x.M * Kd, "wgrad").narrow(0, 0, split * x.M * Kd).view(split, x.M, Kd)
            G.run_gemm(G.kmajor(dy), bop, G.Epilogue(out=part[0]), engine=eng, split_k=split)
            K.sum_partials(part, split, x.M * Kd, tgt, accumulate=acc)
            return
        self._dgrad_gemm(G.kmajor(dy), bop, tgt, tgt if acc else None, x, eng, alpha=scale)

    # ---- GroupNorm ---------------------------------------------------------------------------------------------
    def groupnorm(self, x: Act, nw: NormW, silu: bool, out: Act = None) -> Act:
        if out is None:
            out = self._new(x.M, x.C, x.B, x.H, x.W)
        stats = torch.empty(x.B, nw.groups, 2, dtype=torch.float32, device=self.device)
        gamma, beta = nw.mod.weight.data, nw.mod.bias.data
        fused = K.groupnorm_fused_ok(x.B, x.H * x.W, x.C, nw.groups, x.t.dtype)
        if fused:
            K.groupnorm_fused_fwd(x.t, out.t, gamma, beta, stats, x.B, nw.groups, nw.eps, silu)
        else:
            K.groupnorm_fwd(x.t, out.t, gamma, beta, stats, x.B, nw.groups, nw.eps, silu)
        if self.taping:
            def bwd():
                tgt, acc = self._grad_target(x)
                gw, gb = self.store.grad_of(nw.mod.weight), self.store.grad_of(nw.mod.bias)
                if not fused:
                    K.groupnorm_bwd(out.g(), x.t, tgt, gamma, beta, stats, gw, gb, x.B, nw.groups, silu, acc)
                    return
                colsum = total = None
                if x.sink is not None and not acc:
                    # x is a conv output consumed only here: its gradient's column sums are the conv's bias gradient and
                    # the gradient of its broadcast conditioning; the kernel has them for free
                    rowbias, cw = x.sink
                    colsum, racc = self._rowbias_target(rowbias)
                    if racc:
                        colsum = None
                    else:
                        total = self.store.grad_of(cw.mod.bias)
                        x.colsum_done = True
                K.groupnorm_fused_bwd(out.g(), x.t, tgt, gamma, beta, stats, gw, gb, x.B, nw.groups, silu, acc, colsum, total)
            self.tape.append(bwd)
        return out

    # ---- attention core ----------------------------------------------------------------------------------------
    def attention(self, q: Act, k: Act, v: Act, lq: int, lk: int, heads: int, drop_p: float, site: int) -> Act:
        """q/k/v are channel slices of projection outputs; gradients are written into the parents' grad buffers."""
        c = q.C
        hd = c // heads
        o = self._new(q.M, c, q.B, q.H, q.W)
        p = drop_p if (self.training and self.dropout_enabled) else 0.0
        seed = self._seed(site)
        tensor_core = q.t.dtype == torch.bfloat16 and lk <= 1024
        fused = q.t.dtype == torch.bfloat16 and K.attn_fused_ok(q.B, heads, lq, lk, hd)
        if fused:
            lse = torch.empty(q.B, heads, lq, dtype=torch.float32, device=self.device) if self.taping else None
            K.attn_fused_fwd(q.t, k.t, v.t, o.t, lse, q.B, heads, lq, lk, hd, seed, p)
        elif tensor_core:
            P = K.attn_tc_fwd(q.t, k.t, v.t, o.t, q.B, heads, lq, lk, hd, seed, p)
            lse = None
        else:
            lse = torch.empty(q.B, heads, lq, dtype=torch.float32, device=self.device) if self.taping else None
            K.attn_fwd(q.t, k.t, v.t, o.t, lse, q.B, heads, lq, lk, hd, seed, p)
        if self.taping:
            def bwd():
                for a in (q, k, v):
                    par = a.parent if a.parent is not None else a
                    if par.grad is None:
                        par.grad = torch.empty(par.M, par.C, dtype=par.t.dtype, device=self.device)
                if fused:
                    K.attn_fused_bwd(q.t, k.t, v.t, o.t, o.g(), lse, q.g(), k.g(), v.g(), q.B, heads, lq, lk, hd, seed, p)
                elif tensor_core:
                    K.attn_tc_bwd(q.t, k.t, v.t, o.g(), P, q.g(), k.g(), v.g(), q.B, heads, lq, lk, hd, seed, p)
                else:
                    K.attn_bwd(q.t, k.t, v.t, o.t, o.g(), lse, q.g(), k.g(), v.g(), q.B, heads, lq, lk, hd, seed, p)
            self.tape.append(bwd)
        return o

    # ---- resize / concat ---------------------------------------------------------------------------------------
    def upsample(self, x: Act, size: int) -> Act:
        out = self._new(x.B * size * size, x.C, x.B, size, size)
        K.upsample_fwd(x.t, out.t, x.B, x.H, x.W, size, size)
        if self.taping:
            def bwd():
                tgt, acc = self._grad_target(x)
                K.upsample_bwd(out.g(), tgt, x.B, x.H, x.W, size, size, acc)
            self.tape.append(bwd)
        return out

    def copy_into(self, src: Act, dst_slice: Act) -> None:
        K.copy_strided(src.t, dst_slice.t)
        if self.taping:
            def bwd():
                tgt, acc = self._grad_target(src)
                K.copy_strided(dst_slice.g(), tgt, accumulate=acc)
            self.tape.append(bwd)

    # ------------------------------------------------------------------------------------------------------------
    # blocks
    # ------------------------------------------------------------------------------------------------------------
    def _cond_all(self, temb: Act, pooled: Act) -> Act:
        """time_proj(time_emb) + text_proj(text_pooled) of every ResBlock at once: [B, sum Cout] (unet.py:119-124)."""
        if not (self.bf16 and self.cond_on_tensor_cores):
            cond = self.linear(temb, self.d_cond_time)
            self.linear(pooled, self.d_cond_text, into=cond, x_needs_grad=False)
            return cond
        # bf16 mode: the two [B, K] x [sum Cout, K]^T products (and their three backward GEMMs) run on the tcgen05 engine over
        # bf16 copies of the [B, K] inputs and the bf16 weight shadow, fp32 accumulation and fp32 output (the result is the
        # conv epilogues' broadcast row bias); on the CUDA-core engine they were 1.2 ms of the step at 5 TFLOP/s
        store, B, N = self.store, temb.M, self.cond_width
        lt, lx = self.d_cond_time, self.d_cond_text
        _, b_t, gw_t, gb_t = lt.views(store)
        _, b_x, gw_x, gb_x = lx.views(store)
        if temb.t.dtype == torch.bfloat16:
            t16 = temb.t
        else:
            t16 = torch.empty(B, lt.k, dtype=torch.bfloat16, device=self.device)
            K.cast_bf16(temb.t, t16)
        p16 = torch.empty(B, lx.k, dtype=torch.bfloat16, device=self.device)
        K.cast_bf16(pooled.t, p16)
        cond = Act(torch.empty(B, N, dtype=torch.float32, device=self.device), B, 1, 1)
        G.run_gemm(G.kmajor(t16), G.kmajor(lt.wk), G.Epilogue(out=cond.t, bias=b_t), engine="umma")
        G.run_gemm(G.kmajor(p16), G.kmajor(lx.wk), G.Epilogue(out=cond.t, bias=b_x, accumulate=True), engine="umma")
        if self.taping:
            def bwd():
                _, _, gw_t, gb_t = lt.views(store)          # (also logs the gradient ranges for parallel.GradSync)
                _, _, gw_x, gb_x = lx.views(store)
                dy = cond.g()
                dy16 = torch.empty(B, N, dtype=torch.bfloat16, device=self.device)
                K.cast_bf16(dy, dy16)
                K.colsum(dy, 1, None, gb_t)
                K.copy_strided(gb_t.view(1, N), gb_x.view(1, N))      # both biases are added to the same sum: same gradient
                G.run_gemm(G.mnmajor(dy16), G.mnmajor(t16), G.Epilogue(out=gw_t), engine="umma")
                G.run_gemm(G.mnmajor(dy16), G.mnmajor(p16), G.Epilogue(out=gw_x), engine="umma")
                tgt, acc = self._grad_target(temb)
                epi = G.Epilogue(out=tgt, accumulate=acc) if tgt.dtype == torch.float32 else G.Epilogue(out=tgt, residual=tgt if acc else None)
                G.run_gemm(G.kmajor(dy16), G.mnmajor(lt.wk), epi, engine="umma")
            self.tape.append(bwd)
        return cond

    def _res_block(self, d, x: Act, cond_all: Act, out: Act = None) -> Act:
        a1 = self.groupnorm(x, d["norm1"], silu=True)
        cond = cond_all.slice(d["cond_off"], d["cout"])     # the [B, Cout] broadcast bias of conv1
        h1 = self.conv(a1, d["conv1"], rowbias=cond)
        a2 = self.groupnorm(h1, d["norm2"], silu=True)
        if d["skip"] is None:
            return self.conv(a2, d["conv2"], residual=x, out=out)
        y = self.conv(a2, d["conv2"], out=out)
        self.linear(x, d["skip"], into=y)
        return y

    def _attn_block(self, d, x: Act, text_tok: Act, lt: int, site: int, out: Act = None) -> Act:
        c, B, HW = d["c"], x.B, x.H * x.W
        train_drop = self.training and self.dropout_enabled
        n1 = self.groupnorm(x, d["norm1"], silu=False)
        qkv = self.linear(n1, d["qkv"])
        o = self.attention(qkv.slice(0, c), qkv.slice(c, c), qkv.slice(2 * c, c), HW, HW, self.unet.num_heads, d["p_attn"], site)
        x1 = self.linear(o, d["so"], alpha=0.7, residual=x)
        n2 = self.groupnorm(x1, d["norm2"], silu=False)
        q = self.linear(n2, d["cq"])
        tp = self.linear(text_tok, d["tp"], x_needs_grad=False, B=B, H=lt, W=1)
        kv = self.linear(tp, d["ckv"])
        o2 = self.attention(q, kv.slice(0, c), kv.slice(c, c), HW, lt, self.unet.num_heads, d["p_attn"], site + 1)
        x2 = self.linear(o2, d["co"], alpha=0.8, residual=x1)
        pf = d["p_ffn"]
        f = self.linear(x2, d["f1"], act=L.ACT_GELU, drop=(self._seed(site + 2), pf) if train_drop else None)
        return self.linear(f, d["f2"], alpha=0.6, residual=x2, drop=(self._seed(site + 3), pf) if train_drop else None, out=out)

    def _block(self, d, x: Act, cond_all, text_tok, lt, site, out: Act = None) -> Act:
        if d["attn"] is None:
            return self._res_block(d["res"], x, cond_all, out=out)
        h = self._res_block(d["res"], x, cond_all)
        return self._attn_block(d["attn"], h, text_tok, lt, site, out=out)     # `out`: the FFN down-projection writes the concat half itself

    # ------------------------------------------------------------------------------------------------------------
    # whole network
    # ------------------------------------------------------------------------------------------------------------
    def forward(self, noisy_latent: torch.Tensor, timesteps: torch.Tensor, text_emb: torch.Tensor, need_grad: bool):
        dev = noisy_latent.device
        self.device = dev
        self.prepare(dev)
        self.training = self.unet.training
        self.taping = need_grad
        self.tape: List[Callable] = []
        self.step_counter += 1
        B, lat_c, Hh, Ww = noisy_latent.shape
        if (Hh, Ww) != (LEVELS[0][1], LEVELS[0][1]):
            raise L.PsgError(f"UNet expects {LEVELS[0][1]}x{LEVELS[0][1]} latents, got {Hh}x{Ww}")
        if text_emb.dim() != 3 or text_emb.shape[0] != B or text_emb.shape[2] != self.unet.text_dim:
            raise L.PsgError(f"text_emb must be [B, L, {self.unet.text_dim}], got {tuple(text_emb.shape)}")
        lt = text_emb.shape[1]
        x_in = noisy_latent.detach().contiguous().float()
        text = text_emb.detach().contiguous().float()
        t = timesteps.detach().to(device=dev, dtype=torch.int64).contiguous()

        # ---- conditioning path (M = B rows; fp32 on the CUDA-core engine in parity mode, bf16 tensor-core GEMMs otherwise) ----
        u = self.unet
        coeff = u.time_embed.emb_coeff
        if coeff.device != dev:
            coeff = coeff.to(dev)
        sin = Act(torch.empty(B, 2 * coeff.shape[0], dtype=torch.float32, device=dev), B, 1, 1)
        K.timestep_embedding(t, coeff.float().contiguous(), sin.t)
        if self.bf16 and self.cond_on_tensor_cores:
            # the time MLP ([B, 128] -> 512 -> 512 -> 128) also runs on the tcgen05 engine over a bf16 copy of the sinusoidal
            # embedding (three ~100 us latency-bound CUDA-core launches each way otherwise)
            sin16 = Act(torch.empty(B, sin.C, dtype=torch.bfloat16, device=dev), B, 1, 1)
            K.cast_bf16(sin.t, sin16.t)
            sin = sin16
        h = self.linear(sin, self.d_time[0], act=L.ACT_SILU, x_needs_grad=False)
        h = self.linear(h, self.d_time[1], act=L.ACT_SILU)
        temb = self.linear(h, self.d_time[2])
        pooled = Act(torch.empty(B, text.shape[2], dtype=torch.float32, device=dev), B, 1, 1)
        K.mean_pool(text, pooled.t)
        cond_all = self._cond_all(temb, pooled)
        text_tok = Act(torch.empty(B * lt, text.shape[2], dtype=self.dtype, device=dev), B, lt, 1)
        K.nchw_to_tokens(text.view(B * lt, text.shape[2], 1, 1), text_tok.t)

        # ---- encoder ----
        s0 = LEVELS[0][1]
        cin_p = self.d_init.cin_p
        lat = Act((torch.zeros if cin_p != lat_c else torch.empty)(B * s0 * s0, cin_p, dtype=self.dtype, device=dev), B, s0, s0)
        K.nchw_to_tokens(x_in, lat.t[:, :lat_c])
        x = self.conv(lat, self.d_init, x_needs_grad=False)
        skips = []
        site = 0
        for lvl, (ch, size, _) in enumerate(LEVELS):
            if lvl > 0:
                x = self.conv(x, self.d_down[lvl])
            for blk in self.d_enc[lvl]:
                x = self._block(blk, x, cond_all, text_tok, lt, site)
                site += 4
            skips.append(x)
        # ---- decoder: cat([x, skip]) is one [M, 2C] buffer.  Its x half is written by x's PRODUCER (the middle block, the
        # previous decoder block or the up-sampling conv write straight into it, and read their gradient out of the concat
        # gradient's slice: no strided copy either way); the skip half is one strided copy per block ----
        def new_cat(lvl):
            ch, size, _ = LEVELS[lvl]
            return self._new(B * size * size, 2 * ch, B, size, size)

        cat = new_cat(3)
        self._block(self.d_mid, x, cond_all, text_tok, lt, site, out=cat.slice(0, LEVELS[3][0]))
        site += 4
        for lvl in (3, 2, 1, 0):
            ch, size, _ = LEVELS[lvl]
            skip = skips.pop()
            blocks = self.d_dec[lvl]
            for i, blk in enumerate(blocks):
                self.copy_into(skip, cat.slice(ch, ch))
                if i + 1 < len(blocks):
                    nxt = new_cat(lvl)
                    self._block(blk, cat, cond_all, text_tok, lt, site, out=nxt.slice(0, ch))
                    cat = nxt
                else:
                    x = self._block(blk, cat, cond_all, text_tok, lt, site)
                site += 4
            if lvl > 0:
                x = self.upsample(x, LEVELS[lvl - 1][1])
                cat = new_cat(lvl - 1)
                self.conv(x, self.d_up[lvl], out=cat.slice(0, LEVELS[lvl - 1][0]))
        a = self.groupnorm(x, self.d_final_norm, silu=True)
        y = self.conv(a, self.d_final)
        out = torch.empty(B, lat_c, s0, s0, dtype=torch.float32, device=dev)
        K.tokens_to_nchw(y.t[:, :lat_c], out)
        tape, self.tape = self.tape, []
        self.taping = False
        return out, (tape, y)

    def backward(self, ctx, dout: torch.Tensor, grad_sync=None) -> None:
        """Runs the tape; parameter gradients are written (not accumulated) into self.store.grads.
        `grad_sync` (parallel.GradSync): data-parallel gradient all-reduce, bucket by bucket as the tape finalises the
        flat gradient buffer (tail first), overlapped with the rest of the backward pass."""
        tape, y = ctx
        self.taping = False
        if self.weight_stream_enabled and self.bf16 and dout.is_cuda:
            if self._wstream_obj is None or self._wstream_obj.device != dout.device:
                self._wstream_obj = torch.cuda.Stream(device=dout.device)
            self.wstream = self._wstream_obj
        else:
            self.wstream = None
        lat_c = dout.shape[1]
        y.grad = (torch.zeros if y.C != lat_c else torch.empty)(y.M, y.C, dtype=y.t.dtype, device=y.t.device)
        K.nchw_to_tokens(dout.detach().contiguous().float(), y.grad[:, :lat_c])
        if grad_sync is None:
            try:
                while tape:
                    tape.pop()()
            finally:
                self._join_weight_stream()
            return
        grad_sync.begin(self.store, len(tape))
        # A bucket's all-reduce must follow the kernels of BOTH streams: either the weight stream is joined into the main stream
        # before every bucket (~20 per backward), or the collective is issued with the weight stream current after that stream has
        # waited for the main one (PSG_BUCKET_ISSUE=side).
        if _BUCKET_ISSUE_MAIN:
            grad_sync.before_issue, grad_sync.issue_ctx = self._join_weight_stream, None
        else:
            grad_sync.before_issue, grad_sync.issue_ctx = None, self._bucket_issue_ctx
        try:
            done = 0
            while tape:
                tape.pop()()
                done += 1
                grad_sync.after_entry(done)
        finally:
            grad_sync.finish()
            self._join_weight_stream()

    # ------------------------------------------------------------------------------------------------------------
    # autograd glue
    # ------------------------------------------------------------------------------------------------------------
    def autograd_forward(self, noisy_latent, timesteps, text_emb):
        params = [p for _, p in self.store.named]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if not need:
            out, _ = self.forward(noisy_latent, timesteps, text_emb, need_grad=False)
            return out
        return _UNetFunction.apply(self, noisy_latent, timesteps, text_emb, *params)


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: UNetEngine, x, t, text, *params):
        out, run = eng.forward(x, t, text, need_grad=True)
        ctx.eng, ctx.run = eng, run
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        eng: UNetEngine = ctx.eng
        store = eng.store
        # If a previous backward's views are still installed as p.grad, writing into the same buffer would corrupt
        # autograd's accumulation: start a fresh gradient buffer in that (rare: zero_grad(set_to_none=False)) case.
        p0 = ctx.params[0]
        if p0.grad is not None and p0.grad.data_ptr() == store.grad_of(p0).data_ptr():
            store.fresh_grads()
        eng.backward(ctx.run, dout)
        ctx.run = None
        grads = [store.grad_of(p) if p.requires_grad else None for p in ctx.params]
        return (None, None, None, None, *grads)
