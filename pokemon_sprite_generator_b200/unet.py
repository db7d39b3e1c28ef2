"""Drop-in replacement for the reference U-Net denoiser (`src/models/unet.py`).

`UNet(latent_dim=8, text_dim=256, time_emb_dim=128, num_heads=8)` is an `nn.Module` with the reference's
constructor, `forward(noisy_latent, timesteps, text_emb)` signature and 479-entry state_dict (same keys, shapes,
fp32).  The sub-modules below are *parameter containers only*: they are built from the same torch.nn classes,
in the same registration order, and re-initialised by the same rules as the reference
(`UNet._initialize_weights`, unet.py:405-426; CrossAttentionBlock ctor inits, unet.py:176-193), so that under
the same `torch.manual_seed` the weights are bit-identical to the reference's (checked against
tests/golden/unet_cases.pt).  None of their `forward`s is ever called: all compute goes through
`engine.UNetEngine`, i.e. the hand-written CUDA kernels behind the C ABI.  There is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ._lib import PsgError

# (name, channels, spatial size, has attention) per resolution level -- unet.py:329-353
LEVELS = ((320, 27, False), (640, 14, True), (1280, 7, True), (1280, 4, True))


def _largest_group_count(channels: int, cap: int = 32) -> int:
    g = min(cap, channels)
    while channels % g != 0 and g > 1:
        g -= 1
    return g


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are never executed
        raise PsgError(f"{type(self).__name__} is a parameter container; run the whole UNet (CUDA engine) instead")


class TimestepEmbedding(_Container):
    """Parameters of the sinusoidal timestep embedding + MLP (reference unet.py:12-55)."""

    def __init__(self, embedding_dim: int = 128, max_time: int = 1000):
        super().__init__()
        self.embedding_dim, self.max_time = embedding_dim, max_time
        half = embedding_dim // 2
        self.register_buffer("emb_coeff", torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))))
        hidden = embedding_dim * 4
        self.time_mlp = nn.Sequential(nn.Linear(embedding_dim, hidden), nn.SiLU(), nn.Linear(hidden, hidden), nn.SiLU(),
                                      nn.Linear(hidden, embedding_dim))


class ResBlock(_Container):
    """norm1, conv1, time_proj, text_proj, norm2, conv2, dropout, skip_conv (reference unet.py:58-98)."""

    def __init__(self, in_channels: int, out_channels: int, time_emb_dim: int = 128, text_emb_dim: int = 256, dropout: float = 0.0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(_largest_group_count(in_channels), in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.time_proj = nn.Linear(time_emb_dim, out_channels)
        self.text_proj = nn.Linear(text_emb_dim, out_channels)
        self.norm2 = nn.GroupNorm(_largest_group_count(out_channels), out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        self.dropout = nn.Dropout(dropout)
        self.skip_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else nn.Identity()


class CrossAttentionBlock(_Container):
    """norm1/2 (eps 1e-6), self_attn, cross_attn, text_proj, ffn (reference unet.py:135-193)."""

    ATTN_DROPOUT = 0.05
    FFN_DROPOUT = 0.05

    def __init__(self, channels: int, text_dim: int, num_heads: int = 8):
        super().__init__()
        if channels % num_heads != 0:
            raise AssertionError(f"channels ({channels}) must be divisible by num_heads ({num_heads})")
        self.channels, self.text_dim, self.num_heads, self.head_dim = channels, text_dim, num_heads, channels // num_heads
        groups = max(1, _largest_group_count(channels))
        self.norm1 = nn.GroupNorm(groups, channels, eps=1e-6)
        self.norm2 = nn.GroupNorm(groups, channels, eps=1e-6)
        self.self_attn = nn.MultiheadAttention(channels, num_heads, dropout=self.ATTN_DROPOUT, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(channels, num_heads, dropout=self.ATTN_DROPOUT, batch_first=True)
        self.text_proj = nn.Linear(text_dim, channels)
        _small_xavier(self.text_proj)
        self.ffn = nn.Sequential(nn.Linear(channels, channels * 2), nn.GELU(), nn.Dropout(self.FFN_DROPOUT),
                                 nn.Linear(channels * 2, channels), nn.Dropout(self.FFN_DROPOUT))
        for layer in self.ffn:
            if isinstance(layer, nn.Linear):
                _small_xavier(layer)


def _small_xavier(linear: nn.Linear, gain: float = 0.02) -> None:
    nn.init.xavier_uniform_(linear.weight, gain=gain)
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


class UNetBlock(_Container):
    def __init__(self, in_channels: int, out_channels: int, time_emb_dim: int = 128, text_emb_dim: int = 256,
                 has_attention: bool = True, num_heads: int = 8):
        super().__init__()
        self.has_attention = has_attention
        self.res_block = ResBlock(in_channels, out_channels, time_emb_dim, text_emb_dim)
        if has_attention:
            self.attn_block = CrossAttentionBlock(out_channels, text_emb_dim, num_heads)


class UNet(nn.Module):
    """Text-conditioned U-Net noise predictor, [B, 8, 27, 27] -> [B, 8, 27, 27], computed by sm_100a kernels.

    Extra (keyword-only, non-reference) arguments:
      compute_dtype  torch.bfloat16 (default; tcgen05 tensor cores, fp32 accumulate) or torch.float32
                     (fp32 parity mode: every GEMM/conv runs in the fp32 CUDA-core engine).
    Gradients w.r.t. `noisy_latent` / `text_emb` are not produced (the reference trainer never needs them:
    both are computed under no_grad, improved_diffusion_trainer.py:351-374).
    """

    def __init__(self, latent_dim: int = 8, text_dim: int = 256, time_emb_dim: int = 128, num_heads: int = 8, *,
                 compute_dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.latent_dim, self.text_dim, self.time_emb_dim, self.num_heads = latent_dim, text_dim, time_emb_dim, num_heads
        self.compute_dtype = compute_dtype
        self.time_embed = TimestepEmbedding(time_emb_dim)
        self.text_pool = nn.AdaptiveAvgPool1d(1)

        def blocks(cin, cout, attn):
            return nn.ModuleList([UNetBlock(cin, cout, time_emb_dim, text_dim, has_attention=attn, num_heads=num_heads)
                                  for _ in range(2)])

        c0 = LEVELS[0][0]
        self.init_conv = nn.Conv2d(latent_dim, c0, kernel_size=3, padding=1)
        prev = c0
        for lvl, (ch, _, attn) in enumerate(LEVELS):
            if lvl > 0:
                setattr(self, f"downsample{lvl}", nn.Conv2d(prev, ch, kernel_size=3, stride=2, padding=1))
            setattr(self, f"enc_block{lvl}", blocks(ch, ch, attn))
            prev = ch
        self.middle_block = UNetBlock(prev, prev, time_emb_dim, text_dim, has_attention=True, num_heads=num_heads)
        for lvl in (3, 2, 1, 0):
            ch, _, attn = LEVELS[lvl]
            setattr(self, f"dec_block{lvl}", blocks(ch + ch, ch, attn))
            if lvl > 0:
                nxt, size, _ = LEVELS[lvl - 1][0], LEVELS[lvl - 1][1], None
                setattr(self, f"upsample{lvl}", nn.Sequential(nn.Upsample(size=(size, size), mode="bilinear", align_corners=False),
                                                              nn.Conv2d(ch, nxt, kernel_size=3, padding=1)))
        self.final_conv = nn.Sequential(nn.GroupNorm(32, c0), nn.SiLU(), nn.Conv2d(c0, latent_dim, kernel_size=3, padding=1))
        self._initialize_weights()
        self._engine = None

    def _initialize_weights(self) -> None:
        # reference unet.py:405-426: conv kaiming-normal(fan_out, relu); linear xavier-uniform(gain 0.02); norms 1/0;
        # then the output conv is re-drawn with xavier-uniform(gain 0.02).  MHA in_proj_weight is a bare Parameter
        # and keeps nn.MultiheadAttention's own init; out_proj is an nn.Linear subclass and is re-drawn.
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                _small_xavier(m)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        out_conv = self.final_conv[2]
        nn.init.xavier_uniform_(out_conv.weight, gain=0.02)
        nn.init.zeros_(out_conv.bias)

    # -- engine plumbing ------------------------------------------------------------------------------------------
    def engine(self):
        from .engine import UNetEngine
        if self._engine is None:
            self._engine = UNetEngine(self, self.compute_dtype)
        return self._engine

    def forward(self, noisy_latent: torch.Tensor, timesteps: torch.Tensor, text_emb: torch.Tensor) -> torch.Tensor:
        if not noisy_latent.is_cuda:
            raise PsgError("psg_b200 UNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        return self.engine().autograd_forward(noisy_latent, timesteps, text_emb)
