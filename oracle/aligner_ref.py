"""CPU restatement of the ThinkDiff aligner ``mm_projector`` (type ``mlp2x_gelu_t5_norm``).  TEST INFRASTRUCTURE ONLY.

Reference:
  * ``build_vision_projector``  thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:41-79
        Sequential(0: Linear(Din, D), 1: GELU(), 2: Linear(D, D), 3: T5LayerNorm(D))   (:58-63)
  * ``T5LayerNorm.forward``     transformers==4.46.1 (requirements.txt:14), modeling_t5.py:
        variance = x.to(float32).pow(2).mean(-1, keepdim=True); x = x * rsqrt(variance + eps)   (eps = 1e-6)
        if weight.dtype in (float16, bfloat16): x = x.to(weight.dtype);  return weight * x
  * call sites                  ...embed_decoder_2.py:585 (train, under autocast bf16 base_task.py:237),
                                :761/:998/:1115 (inference), blip_vision_t5_decoder.py:414/:641

Two forms are given and tested against each other and against the exec'd reference (tests/test_oracle_*.py):
  * ``RefAligner``            an nn.Module restatement (autograd gives the backward) -- also the CPU baseline in bench.py
  * ``aligner_fwd_bwd_manual`` closed-form forward/backward with every rounding point of the bf16-autocast regime
                               made explicit -- what the CUDA kernels implement.
Parity pinning: tests/golden/aligner_*.npz were produced by the reference's own ``build_vision_projector``
(oracle/make_golden.py); this file is checked against them in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
import re

import torch
from torch import nn

EPS = 1e-6


class T5RMSNorm(nn.Module):
    """Restatement of transformers' T5LayerNorm: scale-only RMS norm, fp32 statistics, no mean subtraction."""

    def __init__(self, hidden_size: int, eps: float = EPS):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        hidden_states = hidden_states * torch.rsqrt(variance + self.variance_epsilon)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            hidden_states = hidden_states.to(self.weight.dtype)
        return self.weight * hidden_states


def build_ref_projector(mm_hidden_size: int, hidden_size: int, projector_type: str = "mlp2x_gelu_t5_norm") -> nn.Module:
    """Same construction rules as the reference builder (…embed_decoder_2.py:41-79) for the t5_norm / plain MLP / linear types."""
    if projector_type == "linear":
        return nn.Linear(mm_hidden_size, hidden_size)
    m = re.match(r"^mlp(\d+)x_gelu(_t5_norm)?$", projector_type)
    if not m:
        raise ValueError(f"Unknown projector type: {projector_type}")
    depth = int(m.group(1))
    mods = [nn.Linear(mm_hidden_size, hidden_size)]
    for _ in range(1, depth):
        mods.append(nn.GELU())
        mods.append(nn.Linear(hidden_size, hidden_size))
        mods.append(T5RMSNorm(hidden_size) if m.group(2) else nn.Identity())
    return nn.Sequential(*mods)


class RefAligner(nn.Sequential):
    """``mlp2x_gelu_t5_norm`` with the reference's state-dict keys 0.weight 0.bias 2.weight 2.bias 3.weight."""

    def __init__(self, mm_hidden_size: int, hidden_size: int):
        super().__init__(
            nn.Linear(mm_hidden_size, hidden_size), nn.GELU(), nn.Linear(hidden_size, hidden_size), T5RMSNorm(hidden_size)
        )


def init_params_numpy(mm_hidden_size: int, hidden_size: int, seed: int = 0) -> dict:
    """nn.Linear default init (U(+-1/sqrt(fan_in)) for weight and bias) from a numpy generator, so fixtures do not
    depend on the torch RNG stream; norm weight = ones perturbed so that dg parity is meaningful."""
    import numpy as np

    rng = np.random.RandomState(seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    return {
        "0.weight": u((hidden_size, mm_hidden_size), mm_hidden_size),
        "0.bias": u((hidden_size,), mm_hidden_size),
        "2.weight": u((hidden_size, hidden_size), hidden_size),
        "2.bias": u((hidden_size,), hidden_size),
        "3.weight": torch.from_numpy((1.0 + 0.1 * rng.standard_normal(hidden_size)).astype(np.float32)),
    }


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_erf_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def aligner_fwd_bwd_manual(x, params, dy=None, regime: str = "bf16", out_bf16: bool = False, accum_dtype=torch.float64):
    """Closed-form forward (and backward for an upstream gradient ``dy``) of the aligner on rows ``x[M, Din]``.

    regime "bf16": the autocast regime of training: x, W, b are rounded to bf16, products accumulate
                   exactly (here float64), h0/h1/h2 and the activation gradients dh2/dh1/dh0 are rounded to bf16,
                   the norm runs in fp32 on the rounded h2, y is fp32 (``out_bf16``: pure-bf16 inference rule A.2).
    regime "fp32": no rounding anywhere (config 1).
    Returns dict(y, h0, h1, h2, rstd [, dW1, db1, dW2, db2, dg]) as float32 tensors.
    """
    rnd = _bf16 if regime == "bf16" else (lambda t: t.to(torch.float32))
    f = accum_dtype
    W1, b1, W2, b2, g = (params[k].to(torch.float32) for k in ("0.weight", "0.bias", "2.weight", "2.bias", "3.weight"))
    xq, W1q, b1q, W2q, b2q = rnd(x.to(torch.float32)), rnd(W1), rnd(b1), rnd(W2), rnd(b2)
    h0 = rnd((xq.to(f) @ W1q.to(f).T + b1q.to(f)).to(torch.float32))
    h1 = rnd(gelu_erf(h0.to(f)).to(torch.float32))
    h2 = rnd((h1.to(f) @ W2q.to(f).T + b2q.to(f)).to(torch.float32))
    var = h2.to(f).pow(2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + EPS)
    xhat = h2.to(f) * rstd
    if out_bf16:
        y = _bf16(_bf16(g).to(f) * _bf16(xhat.to(torch.float32)).to(f))
    else:
        y = (g.to(f) * xhat).to(torch.float32)
    out = {"y": y.to(torch.float32), "h0": h0, "h1": h1, "h2": h2, "rstd": rstd.to(torch.float32).squeeze(-1)}
    if dy is None:
        return out
    dyf = dy.to(f)
    gh = g.to(f) * dyf
    dh2 = rnd((rstd * gh - h2.to(f) * rstd.pow(3) * (gh * h2.to(f)).mean(-1, keepdim=True)).to(torch.float32))
    out["dg"] = (dyf * xhat).sum(0).to(torch.float32)
    out["dW2"] = (dh2.to(f).T @ h1.to(f)).to(torch.float32)
    out["db2"] = dh2.to(f).sum(0).to(torch.float32)
    dh1 = rnd((dh2.to(f) @ W2q.to(f)).to(torch.float32))
    dh0 = rnd((dh1.to(f) * gelu_erf_grad(h0.to(f))).to(torch.float32))
    out["dW1"] = (dh0.to(f).T @ xq.to(f)).to(torch.float32)
    out["db1"] = dh0.to(f).sum(0).to(torch.float32)
    out["dh2"], out["dh0"] = dh2, dh0
    return out


def module_fwd_bwd(module: nn.Module, x: torch.Tensor, loss_fn, autocast_bf16: bool = False):
    """Run ``module`` (reference or restatement) forward + ``loss_fn(y).backward()`` on CPU; return y, loss, grads."""
    module.zero_grad(set_to_none=True)
    if autocast_bf16:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = module(x)
            loss = loss_fn(y)
    else:
        y = module(x)
        loss = loss_fn(y)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in module.named_parameters()}
    return y.detach(), loss.detach(), grads
