"""Text encoder of the latent-diffusion pipeline: token ids -> [B, L, hidden_dim] conditioning sequence, the
BERT encoder + projection + LayerNorm of `src/models/text_encoder.py:TextEncoder` (:137-163) computed by this library's sm_100a
kernels through the C ABI -- tcgen05 GEMMs with fused bias / GELU / residual epilogues, a key-length-masked attention core, LayerNorm
and embedding-sum kernels.  Inference only: stage 2 and sampling run the text encoder frozen under `no_grad`
(src/training/improved_diffusion_trainer.py:350-353).

Drop-in contract: same constructor arguments (`model_name`, `hidden_dim`, `finetune_strategy`), the same child modules in the same
order (`bert` -- a `transformers.BertModel` that HOLDS the weights --, `projection`, `layer_norm`), hence the same state_dict keys, the
same `requires_grad` pattern per fine-tuning strategy, and `forward(list[str]) -> Tensor`.  Tokenisation stays with
`transformers.BertTokenizer` (string processing, not GPU work); `encode_ids(input_ids, attention_mask)` is the entry point below it.
Without network access `BertModel.from_pretrained` cannot fetch weights: pass `bert=` (a constructed `BertModel`, e.g. from a
`BertConfig` or a checkpoint) and optionally `tokenizer=`.

Padding: the reference tokenises with `padding=True`, i.e. an attention mask whose zeros are a suffix; masked keys get zero weight
inside BERT (`psg_attn_fwd_keylen`), padded positions still produce outputs, exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import gemm as G
from . import ops as K


class TextEncoder(nn.Module):
    def __init__(self, model_name: str = "google-bert/bert-base-uncased", hidden_dim: int = 768, finetune_strategy: str = "minimal", *,
                 bert: Optional[nn.Module] = None, tokenizer=None, compute_dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.finetune_strategy = finetune_strategy
        self.compute_dtype = compute_dtype
        if bert is None:
            from transformers import BertModel, BertTokenizer      # needs the model files (hub cache or a local directory)
            self.tokenizer = BertTokenizer.from_pretrained(model_name)
            self.bert = BertModel.from_pretrained(model_name)
        else:
            self.tokenizer = tokenizer
            self.bert = bert
        self._apply_finetune_strategy()
        self.bert_hidden_size = self.bert.config.hidden_size
        self.projection = nn.Linear(self.bert_hidden_size, hidden_dim) if self.bert_hidden_size != hidden_dim else nn.Identity()
        self.layer_norm = nn.LayerNorm(hidden_dim)
        for p in list(self.projection.parameters()) + list(self.layer_norm.parameters()):
            p.requires_grad = True
        self._packed: Dict[str, torch.Tensor] = {}
        self._packed_key = None

    def _apply_finetune_strategy(self):
        """reference :60-112: which BERT parameters stay trainable (the flags only; this module never back-propagates)."""
        n = len(self.bert.encoder.layer)
        keep = {"none": 0, "minimal": 2, "partial": 4, "full": n}
        if self.finetune_strategy not in keep:
            raise ValueError(f"Unknown finetune_strategy: {self.finetune_strategy}")
        full = self.finetune_strategy == "full"
        for p in self.bert.parameters():
            p.requires_grad = full
        if self.finetune_strategy in ("minimal", "partial"):
            for i in range(max(0, n - keep[self.finetune_strategy]), n):
                for p in self.bert.encoder.layer[i].parameters():
                    p.requires_grad = True
            if getattr(self.bert, "pooler", None) is not None:
                for p in self.bert.pooler.parameters():
                    p.requires_grad = True

    # ---- kernel-side weights: fused QKV matrix per layer, bf16 copies in bf16 mode --------------------------------------------
    def _pack(self, device):
        key = (str(device), self.compute_dtype, tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if key == self._packed_key:
            return
        dt = self.compute_dtype

        def mat(w: torch.Tensor) -> torch.Tensor:
            w = w.detach().float().contiguous()
            if dt == torch.bfloat16:
                out = torch.empty(w.shape, dtype=dt, device=device)
                K.pack_linear_weight(w, out, None)
                return out
            return w

        pk = {}
        for i, layer in enumerate(self.bert.encoder.layer):
            a = layer.attention.self
            pk[f"qkv_w{i}"] = mat(torch.cat([a.query.weight, a.key.weight, a.value.weight], 0))
            pk[f"qkv_b{i}"] = torch.cat([a.query.bias, a.key.bias, a.value.bias], 0).detach().float().contiguous()
            pk[f"ao_w{i}"] = mat(layer.attention.output.dense.weight)
            pk[f"up_w{i}"] = mat(layer.intermediate.dense.weight)
            pk[f"dn_w{i}"] = mat(layer.output.dense.weight)
        if isinstance(self.projection, nn.Linear):
            pk["proj_w"] = mat(self.projection.weight)
        self._packed, self._packed_key = pk, key

    def _linear(self, x, w, bias, act=L.ACT_NONE, residual=None):
        out = torch.empty(x.shape[0], w.shape[0], dtype=x.dtype, device=x.device)
        eng = "umma" if x.dtype == torch.bfloat16 else "simt"
        G.run_gemm(G.kmajor(x), G.kmajor(w), G.Epilogue(out=out, bias=bias, act=act, residual=residual), engine=eng)
        return out

    @staticmethod
    def _layernorm(x, ln: nn.LayerNorm, out_dtype):
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        L.call("psg_layernorm", L.ptr(x), C.c_longlong(x.stride(0)), L.ptr(y), C.c_longlong(y.stride(0)), L.ptr(ln.weight.data),
               L.ptr(ln.bias.data), C.c_longlong(x.shape[0]), C.c_int(x.shape[1]), C.c_float(ln.eps), C.c_int(L.dt(x)), C.c_int(L.dt(y)),
               L.stream_ptr())
        return y

    @torch.no_grad()
    def encode_ids(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                   token_type_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """BertModel(...).last_hidden_state -> projection -> LayerNorm, fp32 [B, L, hidden_dim]."""
        if not input_ids.is_cuda:
            raise L.PsgError("TextEncoder runs on CUDA only (there is no CPU fallback)")
        cfg = self.bert.config
        if getattr(cfg, "hidden_act", "gelu") != "gelu" or getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise L.PsgError("TextEncoder kernels implement the standard BERT encoder (erf GELU, absolute positions)")
        dev, dt = input_ids.device, self.compute_dtype
        self._pack(dev)
        B, Lt = input_ids.shape
        D, H = cfg.hidden_size, cfg.num_attention_heads
        hd = D // H
        rows = B * Lt
        ids = input_ids.to(torch.long).contiguous()
        tt = token_type_ids.to(torch.long).contiguous() if token_type_ids is not None else None
        if attention_mask is not None:
            m = attention_mask.to(dev)
            klen = m.sum(1).to(torch.int32).contiguous()
            # a suffix mask is the only shape the tokenizer produces (padding=True pads on the right); anything else is refused
            if not bool((m.to(torch.bool) == (torch.arange(Lt, device=dev)[None, :] < klen[:, None])).all()):
                raise L.PsgError("TextEncoder: attention_mask must be a prefix of ones per sample (right padding)")
        else:
            klen = None
        emb = self.bert.embeddings
        x32 = torch.empty(rows, D, dtype=torch.float32, device=dev)
        L.call("psg_bert_embed", L.ptr(ids), L.ptr(tt), L.ptr(emb.word_embeddings.weight.data), L.ptr(emb.position_embeddings.weight.data),
               L.ptr(emb.token_type_embeddings.weight.data), L.ptr(x32), C.c_longlong(rows), C.c_int(Lt), C.c_int(D),
               C.c_int(cfg.vocab_size), L.stream_ptr())
        x = self._layernorm(x32, emb.LayerNorm, dt)
        pk = self._packed
        for i, layer in enumerate(self.bert.encoder.layer):
            qkv = self._linear(x, pk[f"qkv_w{i}"], pk[f"qkv_b{i}"])
            ctx = torch.empty(rows, D, dtype=dt, device=dev)
            L.call("psg_attn_fwd_keylen", L.ptr(qkv[:, :D]), C.c_longlong(qkv.stride(0)), L.ptr(qkv[:, D:2 * D]), C.c_longlong(qkv.stride(0)),
                   L.ptr(qkv[:, 2 * D:]), C.c_longlong(qkv.stride(0)), L.ptr(ctx), C.c_longlong(ctx.stride(0)), L.ptr(None), L.ptr(klen),
                   C.c_int(B), C.c_int(H), C.c_int(Lt), C.c_int(Lt), C.c_int(hd), C.c_float(1.0 / (hd ** 0.5)), C.c_int(L.dt(qkv)),
                   C.c_ulonglong(0), C.c_float(0.0), L.stream_ptr())
            ao = layer.attention.output
            h1 = self._layernorm(self._linear(ctx, pk[f"ao_w{i}"], ao.dense.bias.data, residual=x), ao.LayerNorm, dt)
            up = self._linear(h1, pk[f"up_w{i}"], layer.intermediate.dense.bias.data, act=L.ACT_GELU)
            x = self._layernorm(self._linear(up, pk[f"dn_w{i}"], layer.output.dense.bias.data, residual=h1), layer.output.LayerNorm, dt)
        if isinstance(self.projection, nn.Linear):
            x = self._linear(x, pk["proj_w"], self.projection.bias.data)
        y = self._layernorm(x, self.layer_norm, torch.float32)
        return y.view(B, Lt, -1)

    def forward(self, text_list):
        if self.tokenizer is None:
            raise L.PsgError("TextEncoder.forward(list[str]) needs a tokenizer (pass tokenizer= or use encode_ids)")
        inputs = self.tokenizer(text_list, return_tensors="pt", padding=True, truncation=True, max_length=256)
        dev = next(self.bert.parameters()).device
        return self.encode_ids(inputs["input_ids"].to(dev), inputs["attention_mask"].to(dev),
                               inputs["token_type_ids"].to(dev) if "token_type_ids" in inputs else None)
