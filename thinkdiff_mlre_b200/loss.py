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


class _LMHeadCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, weight, labels):
        s2 = seq.reshape(-1, seq.shape[-1]).to(torch.bfloat16).contiguous()
        w = weight.detach()
        w = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)  # autocast's cast of the frozen fp32 weight
        loss, dseq, _ = ops.lm_head_ce(s2, w.contiguous(), labels.reshape(-1).contiguous(), 1.0, want_grad=ctx.needs_input_grad[0])
        ctx.dseq, ctx.shape, ctx.dtype = dseq, seq.shape, seq.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        dseq = ctx.dseq
        ctx.dseq = None
        return (dseq.mul_(g.to(dseq.dtype))).view(ctx.shape).to(ctx.dtype), None, None


def lm_head_cross_entropy(sequence_output: torch.Tensor, lm_head_weight: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """``CrossEntropyLoss(ignore_index=-100)(lm_head(sequence_output).view(-1, V), labels.view(-1))`` of the frozen T5 decoder
    (thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:236-246) as one call: tcgen05 GEMM -> single-read CE (the logits' gradient
    overwrites them in place) -> tcgen05 GEMM back to ``sequence_output``. Pass the weight as bfloat16 (``model.language_model.to(
    torch.bfloat16)`` for the frozen parts) to skip the per-call cast autocast would otherwise repeat. The head is frozen (the whole T5 is,
    ...embed_decoder_2.py:715-717, `freeze_language`), so no weight gradient is formed. SURVEY.md section 8 f-1, first slice."""
    if lm_head_weight.requires_grad:
        raise NotImplementedError("lm_head_cross_entropy: the T5 head is frozen in ThinkDiff; a trainable head has no kernel path")
    return _LMHeadCE.apply(sequence_output, lm_head_weight, labels)


class _FrozenLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        w = weight.detach()
        w = (w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)).contiguous()
        ctx.w, ctx.shape, ctx.dtype = w, x.shape, x.dtype
        return ops.linear_bf16(x2, w).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        dx = ops.linear_bf16_dx(dy.reshape(-1, dy.shape[-1]).to(torch.bfloat16).contiguous(), ctx.w)
        return dx.view(ctx.shape).to(ctx.dtype), None


def frozen_linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Bias-free ``nn.Linear`` with a frozen weight under bf16 autocast, forward and input gradient on the tcgen05 GEMM -- the
    shape of every projection of the frozen T5 decoder. First use: the cross-attention K / V projections of the aligner output
    (``T5Attention.k`` / ``.v`` applied to ``encoder_hidden_states``, transformers T5 as pinned by requirements.txt:14, called from
    ...embed_decoder_2.py:211-224), which can take the PACKED rows ``[M, 4096]`` straight from ``forward_packed``: with
    ``weight = cat([Wk, Wv])`` one GEMM yields ``[M, 2 * inner]`` and one GEMM returns the gradient to the aligner output."""
    if weight.requires_grad:
        raise NotImplementedError("frozen_linear: the weight must not require grad (the T5 decoder is frozen in ThinkDiff)")
    return _FrozenLinear.apply(x, weight)

