"""AdamW for the aligner as ONE pass per gradient bucket.

Reference optimiser: ``torch.optim.AdamW`` with decoupled weight decay 0.05 on the two 2-D weights and none on biases /
the norm weight (thinkdiff/runners/runner_base.py:98-127). Same update rule here (parity-tested against torch), but the
kernel (``td_adamw_step``) also writes the bf16 compute copies the next forward's GEMMs read -- autocast's per-call casts
disappear -- and it consumes gradients straight from the flat all-reduce buckets. With ``enable_data_parallel(defer_wait=
True)`` the Linear2 bucket is updated while the Linear1 bucket's all-reduce is still on the wire.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .aligner import ThinkDiffAligner
from .train_step import reference_param_groups


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, aligner: ThinkDiffAligner, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.05):
        if not isinstance(aligner, ThinkDiffAligner):
            raise TypeError("FusedAdamW drives a ThinkDiffAligner (it writes the module's bf16 compute copies)")
        groups = reference_param_groups(aligner, weight_decay)
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.aligner = aligner
        self.grad_scale = 1.0  # multiply gradients by this before the update (1 / static loss scale)
        self._t = 0

    def _group_of(self, p):
        for g in self.param_groups:
            if any(p is q for q in g["params"]):
                return g
        raise KeyError("parameter not in any group")

    @torch.no_grad()
    def step_bucket(self, name: str, t: int | None = None, release_grads: bool = False):
        """Update the parameters of one gradient bucket ('linear2' = 2.weight, 2.bias, 3.weight; 'linear1' = 0.weight,
        0.bias), after ordering the stream behind that bucket's all-reduce. ``t`` = optimizer step number the gradients
        belong to (defaults to the current one); ``release_grads`` drops the ``.grad`` references afterwards."""
        a = self.aligner
        t = self._t if t is None else t
        a.wait_bucket(name)
        named = dict(a.named_parameters())
        bf16 = dict(zip(("0.weight", "0.bias", "2.weight", "2.bias"), a._bf16_buffers()))
        ps = [named[k] for k in ThinkDiffAligner.BUCKETS[name] if named[k].grad is not None]
        keys = [k for k in ThinkDiffAligner.BUCKETS[name] if named[k].grad is not None]
        if not ps:
            return
        g0 = self._group_of(ps[0])
        n = len(ps)
        arr = lambda ptrs: (C.c_void_p * n)(*ptrs)  # noqa: E731
        states = []
        for p in ps:
            st = self.state[p]
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] = t
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise TypeError("FusedAdamW needs contiguous float32 parameters and gradients")
            states.append(st)
        groups = [self._group_of(p) for p in ps]
        if any(g["lr"] != g0["lr"] or g["betas"] != g0["betas"] or g["eps"] != g0["eps"] for g in groups):
            raise ValueError("FusedAdamW: lr / betas / eps must agree across the groups of one bucket")
        L.launch_count += 1
        L.check(
            L.lib().td_adamw_step(
                n, arr([p.data_ptr() for p in ps]), arr([p.grad.data_ptr() for p in ps]),
                arr([s["exp_avg"].data_ptr() for s in states]), arr([s["exp_avg_sq"].data_ptr() for s in states]),
                arr([bf16[k].data_ptr() if k in bf16 else None for k in keys]),
                (C.c_int64 * n)(*[p.numel() for p in ps]), (C.c_float * n)(*[g["weight_decay"] for g in groups]),
                g0["lr"], g0["betas"][0], g0["betas"][1], g0["eps"], t, self.grad_scale, L.stream_ptr()),
            "td_adamw_step",
        )
        if release_grads:
            for p in ps:
                p.grad = None

    def next_step_number(self) -> int:
        self._t += 1
        return self._t

    def mark_bf16_current(self):
        a = self.aligner
        ps = (a[0].weight, a[0].bias, a[2].weight, a[2].bias)
        a._cache_key = tuple((p.data_ptr(), p._version) for p in ps)
        a._bf16_fresh = True

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._t += 1
        order = ("linear1", "linear2") if self.aligner._bwd_order == "linear1_first" else ("linear2", "linear1")
        self.step_bucket(order[0])  # its all-reduce finished first; this update overlaps the other bucket's all-reduce
        self.step_bucket(order[1])
        a = self.aligner
        ps = (a[0].weight, a[0].bias, a[2].weight, a[2].bias)
        a._cache_key = tuple((p.data_ptr(), p._version) for p in ps)
        a._bf16_fresh = True
        return loss
