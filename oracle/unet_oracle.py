"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference U-Net forward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
path (pokemon_sprite_generator_b200/) never does.

A functional re-statement of `UNet.forward` (reference src/models/unet.py:428-509) over a plain state_dict
with the reference's 479 keys, written with torch.nn.functional on CPU tensors.  It follows:
  TimestepEmbedding.forward      unet.py:36-55
  ResBlock.forward               unet.py:100-132
  CrossAttentionBlock.forward    unet.py:195-260   (nn.MultiheadAttention semantics restated explicitly)
  UNet.forward                   unet.py:428-509
Pinned against the real reference by oracle/make_golden.py -> tests/golden/unet_*.pt (tests/test_oracle.py).
Dropout is not modelled: parity is defined in eval mode.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LEVEL_SIZES = {0: (27, 27), 1: (14, 14), 2: (7, 7), 3: (4, 4)}


def _groups(c: int) -> int:
    g = min(32, c)
    while c % g != 0 and g > 1:
        g -= 1
    return g


def timestep_embedding(sd, t: torch.Tensor) -> torch.Tensor:
    emb = t.float().unsqueeze(-1) * sd["time_embed.emb_coeff"].unsqueeze(0)
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    h = F.silu(F.linear(emb, sd["time_embed.time_mlp.0.weight"], sd["time_embed.time_mlp.0.bias"]))
    h = F.silu(F.linear(h, sd["time_embed.time_mlp.2.weight"], sd["time_embed.time_mlp.2.bias"]))
    return F.linear(h, sd["time_embed.time_mlp.4.weight"], sd["time_embed.time_mlp.4.bias"])


def res_block(sd, pre: str, x, temb, pooled):
    cin = x.shape[1]
    h = F.silu(F.group_norm(x, _groups(cin), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps=1e-5))
    h = F.conv2d(h, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    h = h + F.linear(temb, sd[pre + "time_proj.weight"], sd[pre + "time_proj.bias"])[:, :, None, None]
    h = h + F.linear(pooled, sd[pre + "text_proj.weight"], sd[pre + "text_proj.bias"])[:, :, None, None]
    cout = h.shape[1]
    h = F.silu(F.group_norm(h, _groups(cout), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps=1e-5))
    h = F.conv2d(h, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    if pre + "skip_conv.weight" in sd:
        x = F.conv2d(x, sd[pre + "skip_conv.weight"], sd[pre + "skip_conv.bias"])
    return h + x


def mha(sd, pre: str, q_in, kv_in, heads: int):
    """nn.MultiheadAttention(batch_first=True) forward, eval mode, no masks."""
    c = q_in.shape[-1]
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    q = F.linear(q_in, w[:c], b[:c])
    k = F.linear(kv_in, w[c:2 * c], b[c:2 * c])
    v = F.linear(kv_in, w[2 * c:], b[2 * c:])
    bsz, lq, _ = q.shape
    lk = k.shape[1]
    hd = c // heads
    q = q.view(bsz, lq, heads, hd).transpose(1, 2)
    k = k.view(bsz, lk, heads, hd).transpose(1, 2)
    v = v.view(bsz, lk, heads, hd).transpose(1, 2)
    p = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(bsz, lq, c)
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def attn_block(sd, pre: str, x, text, heads: int):
    bsz, c, hh, ww = x.shape
    tok = x.reshape(bsz, c, hh * ww).permute(0, 2, 1)
    g = max(1, _groups(c))

    def gn(t, name):
        return F.group_norm(t.permute(0, 2, 1), g, sd[pre + name + ".weight"], sd[pre + name + ".bias"], eps=1e-6).permute(0, 2, 1)

    n1 = gn(tok, "norm1")
    tok = tok + 0.7 * mha(sd, pre + "self_attn.", n1, n1, heads)
    n2 = gn(tok, "norm2")
    tp = F.linear(text, sd[pre + "text_proj.weight"], sd[pre + "text_proj.bias"])
    tok = tok + 0.8 * mha(sd, pre + "cross_attn.", n2, tp, heads)
    f = F.linear(F.gelu(F.linear(tok, sd[pre + "ffn.0.weight"], sd[pre + "ffn.0.bias"])), sd[pre + "ffn.3.weight"],
                 sd[pre + "ffn.3.bias"])
    tok = tok + 0.6 * f
    return tok.permute(0, 2, 1).reshape(bsz, c, hh, ww)


def unet_block(sd, pre: str, x, temb, pooled, text, heads: int):
    x = res_block(sd, pre + "res_block.", x, temb, pooled)
    if pre + "attn_block.norm1.weight" in sd:
        x = attn_block(sd, pre + "attn_block.", x, text, heads)
    return x


def unet_forward(sd, noisy_latent: torch.Tensor, timesteps: torch.Tensor, text_emb: torch.Tensor, num_heads: int = 8):
    temb = timestep_embedding(sd, timesteps)
    pooled = text_emb.mean(dim=1)
    x = F.conv2d(noisy_latent, sd["init_conv.weight"], sd["init_conv.bias"], padding=1)
    skips = []
    for lvl in range(4):
        if lvl > 0:
            x = F.conv2d(x, sd[f"downsample{lvl}.weight"], sd[f"downsample{lvl}.bias"], stride=2, padding=1)
        for i in range(2):
            x = unet_block(sd, f"enc_block{lvl}.{i}.", x, temb, pooled, text_emb, num_heads)
        skips.append(x)
    x = unet_block(sd, "middle_block.", x, temb, pooled, text_emb, num_heads)
    for lvl in (3, 2, 1, 0):
        skip = skips.pop()
        for i in range(2):
            x = unet_block(sd, f"dec_block{lvl}.{i}.", torch.cat([x, skip], dim=1), temb, pooled, text_emb, num_heads)
        if lvl > 0:
            x = F.interpolate(x, size=LEVEL_SIZES[lvl - 1], mode="bilinear", align_corners=False)
            x = F.conv2d(x, sd[f"upsample{lvl}.1.weight"], sd[f"upsample{lvl}.1.bias"], padding=1)
    x = F.silu(F.group_norm(x, 32, sd["final_conv.0.weight"], sd["final_conv.0.bias"], eps=1e-5))
    return F.conv2d(x, sd["final_conv.2.weight"], sd["final_conv.2.bias"], padding=1)
