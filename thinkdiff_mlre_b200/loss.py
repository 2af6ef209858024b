"""Masked losses on the aligner hot path, forward and gradient produced by one kernel pass.

* ``masked_cross_entropy`` replaces ``CrossEntropyLoss(ignore_index=-100)(lm_logits.view(-1, V), labels.view(-1))``
  (thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:241-246; blip_vision_t5_decoder.py:222-227).
* ``masked_mse`` is the alignment loss BASELINE.json's north_star adds against T5 targets (the reference has no MSE):
  ``F.mse_loss(y[valid].float(), t[valid].float())``.

Both come as autograd functions (drop-in in a ``loss.backward()`` training loop, GradScaler-compatible: the saved
gradient is multiplied by the upstream scalar) and as explicit ``*_fwd_bwd`` calls in ``ops`` that take the gradient scale
up front and return ``(loss, grad)`` with no second pass -- the form the benchmark step uses.
"""
from __future__ import annotations

import torch

from . import ops


class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, target, row_mask):
        y2, t2 = y.reshape(-1, y.shape[-1]), target.reshape(-1, target.shape[-1])
        mask = None if row_mask is None else row_mask.reshape(-1).to(torch.int64).contiguous()
        loss, dy = ops.masked_mse_fwd_bwd(y2.contiguous(), t2.contiguous(), mask, 1.0, want_grad=ctx.needs_input_grad[0])
        ctx.dy, ctx.shape = dy, y.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        dy = ctx.dy
        ctx.dy = None
        return (dy.mul_(g.to(dy.dtype))).view(ctx.shape), None, None


def masked_mse(y: torch.Tensor, target: torch.Tensor, row_mask: torch.Tensor | None = None) -> torch.Tensor:
    """Mean of (y - target)^2 over rows whose mask is non-zero (all rows when ``row_mask`` is None)."""
    return _MaskedMSE.apply(y, target, row_mask)


class _MaskedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        z = logits.reshape(-1, logits.shape[-1]).contiguous()
        loss, dz = ops.masked_ce_fwd_bwd(z, labels.reshape(-1).contiguous(), 1.0, want_grad=ctx.needs_input_grad[0])
        ctx.dz, ctx.shape = dz, logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        dz = ctx.dz
        ctx.dz = None
        return (dz.mul_(g.to(dz.dtype))).view(ctx.shape), None


def masked_cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """``CrossEntropyLoss(ignore_index=-100)`` over ``logits[..., V]`` / ``labels[...]`` (int64)."""
    return _MaskedCE.apply(logits, labels)
