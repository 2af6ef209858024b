"""fp32 regime of the aligner (fp32 parameters, no autocast -- BASELINE config 1) on the bf16 tensor cores.

tcgen05 has no fp32-input MMA and ``kind::tf32`` keeps 10 mantissa bits (~1e-3), two orders short of the fp32
rtol 1e-5 the north star asks for. So every fp32 GEMM is computed as an error-compensated sum of bf16 GEMMs with fp32
accumulation in TMEM: each operand is split exactly into three bf16 terms ``a = a1 + a2 + a3`` (8 + 8 + 8 mantissa bits)
and the six products of weight >= 2^-24 are kept:

    a.b ~= a1.b3 + a3.b1 + a2.b2 + a1.b2 + a2.b1 + a1.b1        (smallest terms first)

The six products run on the same tcgen05 kernel: the split terms are interleaved along the contraction dimension in
64-element chunks (K' = 6K) and the contraction is cut into up to 16 slices, one launch each, accumulated into the fp32
output by the kernel's TMA reduce-add store -- so each TMEM accumulation chain stays short. This is a parity regime, not a
throughput path (6x the MMA work): the splitting and the bias / GELU / norm elementwise steps are plain fp32 torch ops, the
contractions run on ``td_gemm_bf16_f32out``.
"""
from __future__ import annotations

import math

import torch

from . import ops

# products ordered smallest first, the dominant a1.b1 last: TMEM accumulation rounds relative to the running sum, so the
# 2^-16 and 2^-8 terms are added while the accumulator is still small
_A_PATTERN = (0, 2, 1, 0, 1, 0)
_B_PATTERN = (2, 0, 1, 1, 0, 0)
_CHUNK = 64  # one k-block of the GEMM


def _split3(a: torch.Tensor):
    a1 = a.to(torch.bfloat16)
    r = a - a1.float()          # exact
    a2 = r.to(torch.bfloat16)
    a3 = (r - a2.float()).to(torch.bfloat16)
    return a1, a2, a3


def _expand(a: torch.Tensor, pattern, k_dim: int) -> torch.Tensor:
    """[.., K, ..] fp32 -> bf16 with K' = 6 * ceil64(K): for every 64-wide chunk of K the six split terms follow each
    other, so any split-K slice of the expanded operand holds complete (small ..., dominant) groups."""
    K = a.shape[k_dim]
    pad = (-K) % _CHUNK
    if pad:
        shape = list(a.shape)
        shape[k_dim] = pad
        a = torch.cat([a, a.new_zeros(shape)], dim=k_dim)
    s = _split3(a)
    if k_dim == 1:  # K-major [R, K]
        R = a.shape[0]
        parts = [s[i].view(R, -1, 1, _CHUNK) for i in pattern]
        return torch.cat(parts, dim=2).reshape(R, -1).contiguous()
    R = a.shape[1]  # MN-major [K, R]
    parts = [s[i].view(-1, 1, _CHUNK, R) for i in pattern]
    return torch.cat(parts, dim=1).reshape(-1, R).contiguous()


def gemm_fp32(A, B, a_mn_major: bool, b_mn_major: bool) -> torch.Tensor:
    """fp32-accurate D = A.B^T; K-major operands are [rows, K], MN-major operands [K, rows]. The contraction is cut
    into up to 16 slices (whole groups of the six split terms) whose fp32 partial sums are combined with round-to-nearest
    adds (the kernel's reduce-add store), keeping every tensor-core accumulation chain short."""
    Ae = _expand(A, _A_PATTERN, 0 if a_mn_major else 1)
    Be = _expand(B, _B_PATTERN, 0 if b_mn_major else 1)
    group = 6 * _CHUNK
    groups = (Ae.shape[0] if a_mn_major else Ae.shape[1]) // group
    slices = max(1, min(16, groups))
    per = (groups + slices - 1) // slices * group
    out = None
    for k0 in range(0, groups * group, per):
        a = Ae[k0 : k0 + per] if a_mn_major else Ae[:, k0 : k0 + per]
        b = Be[k0 : k0 + per] if b_mn_major else Be[:, k0 : k0 + per]
        out = ops.gemm_f32out(a, b, a_mn_major, b_mn_major, out=out)
    return out


def _gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


class _AlignerFp32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, g, eps, dp):
        h0 = gemm_fp32(x, W1, False, False).add_(b1)
        h1 = torch.nn.functional.gelu(h0)
        h2 = gemm_fp32(h1, W2, False, False).add_(b2)
        rstd = torch.rsqrt(h2.pow(2).mean(-1, keepdim=True) + eps)
        y = g * (h2 * rstd)
        ctx.save_for_backward(x, h0, h1, h2, rstd, W2, g)
        ctx.dp = dp
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h0, h1, h2, rstd, W2, g = ctx.saved_tensors
        dp = ctx.dp
        dy = dy.float()
        gh = g * dy
        dh2 = rstd * gh - h2 * rstd.pow(3) * (gh * h2).mean(-1, keepdim=True)
        dg = (dy * h2 * rstd).sum(0)
        db2 = dh2.sum(0)
        dW2 = gemm_fp32(dh2, h1, True, True)           # dh2^T . h1, contraction over tokens
        dh1 = gemm_fp32(dh2, W2, False, True)          # dh2 . W2, contraction over W2's row index
        dh0 = dh1 * _gelu_grad(h0)
        dW1 = gemm_fp32(dh0, x, True, True)
        db1 = dh0.sum(0)
        grads = [dW1, db1, dW2, db2, dg]
        if dp is not None and dp.world > 1:
            works = [dp.all_reduce_async(t.div_(dp.world)) for t in grads]
            for w in works:
                w.wait()
        return (None, *grads, None, None)


def aligner_fp32(module, x2d: torch.Tensor) -> torch.Tensor:
    w1, b1, w2, b2, g = module[0].weight, module[0].bias, module[2].weight, module[2].bias, module[3].weight
    with torch.autocast("cuda", enabled=False):
        return _AlignerFp32Fn.apply(x2d, w1, b1, w2, b2, g, module.eps, module._dp)
