"""Fused SmoothL1 (Huber) loss forward+backward, replacing nn.SmoothL1Loss(beta=0.1) + its autograd backward
(reference: src/training/improved_diffusion_trainer.py:300,388,396).  One pass over pred/target produces the
mean loss and dL/dpred; the reduction is a fixed-order two-level sum (deterministic)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_WS = {}


def _workspace(device) -> torch.Tensor:
    ws = _WS.get(device)
    if ws is None:
        ws = torch.zeros(1024 + 8, dtype=torch.float32, device=device)
        _WS[device] = ws
    return ws


def smooth_l1_fwd_bwd(pred: torch.Tensor, target: torch.Tensor, beta: float = 0.1, grad_scale: float = 1.0,
                      want_grad: bool = True):
    """Returns (loss[0-dim fp32], dL/dpred or None); mean reduction."""
    if not pred.is_cuda:
        raise L.PsgError("smooth_l1_fwd_bwd: CUDA tensors required (no CPU fallback)")
    assert pred.shape == target.shape
    p = pred.detach().contiguous().float()
    t = target.detach().contiguous().float()
    grad = torch.empty_like(p) if want_grad else None
    loss = torch.empty((), dtype=torch.float32, device=p.device)
    L.call("psg_smooth_l1_fwd_bwd", L.ptr(p), L.ptr(t), L.ptr(grad), L.ptr(loss), L.ptr(_workspace(p.device)),
           C.c_longlong(p.numel()), C.c_float(beta), C.c_float(grad_scale), L.stream_ptr())
    return loss, grad


class _SmoothL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, beta):
        loss, grad = smooth_l1_fwd_bwd(pred, target, beta)
        ctx.save_for_backward(grad)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None, None


class SmoothL1Loss(torch.nn.Module):
    """Drop-in for nn.SmoothL1Loss(beta=...) with mean reduction."""

    def __init__(self, beta: float = 1.0):
        super().__init__()
        self.beta = beta

    def forward(self, pred, target):
        return _SmoothL1Fn.apply(pred, target, self.beta)
