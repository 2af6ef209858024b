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
        self._named_cache = None
        self._group_cache = {}

    def _grads(self, name: str):
        """{key: gradient} for one bucket. Under data parallel the gradients are read from the flat buckets the backward
        wrote (and the collectives reduced in place); ``.grad`` must alias them -- a mismatch means autograd accumulated or
        cloned (e.g. two backward passes per step), which the in-place bucket reduction cannot support."""
        a = self.aligner
        named = self._named()
        keys = ThinkDiffAligner.BUCKETS[name]
        dp = a._dp
        if a._grad_flats is None or a._grad_flats.get(name) is None:
            return {k: named[k].grad for k in keys if named[k].grad is not None}
        views = dict(zip(keys, a.bucket_grads(name)))
        for k, v in views.items():
            g = named[k].grad
            if g is not None and g.data_ptr() != v.data_ptr():
                # autograd accumulated into (or cloned) .grad: the flat buckets of the LAST backward are not the gradient
                if dp is not None and dp.world > 1:
                    raise RuntimeError(
                        f"{k}.grad does not alias the all-reduced gradient bucket: data-parallel FusedAdamW supports exactly one "
                        "autograd backward per optimizer step -- use AlignerTrainStep(accum_grad_iters=...) for gradient accumulation")
                return {k: named[k].grad for k in keys if named[k].grad is not None}
        return views

    def _named(self):
        """name -> Parameter (cached: the aligner's parameter objects do not change)."""
        if self._named_cache is None:
            self._named_cache = dict(self.aligner.named_parameters())
        return self._named_cache

    def _bf16(self):
        return dict(zip(("0.weight", "0.bias", "2.weight", "2.bias"), self.aligner._bf16_buffers()))

    def _group_of(self, p):
        g = self._group_cache.get(id(p))
        if g is None:
            for g in self.param_groups:
                if any(p is q for q in g["params"]):
                    self._group_cache[id(p)] = g
                    return g
            raise KeyError("parameter not in any group")
        return g

    @torch.no_grad()
    def step_bucket(self, name: str, t: int | None = None, release_grads: bool = False, ctl=None):
        """Update the parameters of one gradient bucket ('linear2' = 2.weight, 2.bias, 3.weight; 'linear1' = 0.weight,
        0.bias), after ordering the stream behind that bucket's all-reduce. ``t`` = optimizer step number the gradients
        belong to (defaults to the current one); ``release_grads`` drops the ``.grad`` references afterwards."""
        a = self.aligner
        if a._dp is not None and a._dp.peer:
            raise RuntimeError("peer data parallel: parameters are updated by AlignerTrainStep(pipelined=True), not by step()")
        t = self._t if t is None else t
        a.wait_bucket(name)
        named = self._named()
        bf16 = self._bf16()
        grads = self._grads(name)
        keys = [k for k in ThinkDiffAligner.BUCKETS[name] if k in grads]
        ps = [named[k] for k in keys]
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
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdamW needs contiguous float32 parameters")
            states.append(st)
        groups = [self._group_of(p) for p in ps]
        if any(g["lr"] != g0["lr"] or g["betas"] != g0["betas"] or g["eps"] != g0["eps"] for g in groups):
            raise ValueError("FusedAdamW: lr / betas / eps must agree across the groups of one bucket")
        L.launch_count += 1
        L.check(
            L.lib().td_adamw_step(
                n, arr([p.data_ptr() for p in ps]), arr([grads[k].data_ptr() for k in keys]),
                arr([s["exp_avg"].data_ptr() for s in states]), arr([s["exp_avg_sq"].data_ptr() for s in states]),
                arr([bf16[k].data_ptr() if k in bf16 else None for k in keys]),
                (C.c_int64 * n)(*[p.numel() for p in ps]), (C.c_float * n)(*[g["weight_decay"] for g in groups]),
                g0["lr"], g0["betas"][0], g0["betas"][1], g0["eps"], t, self.grad_scale, L.ptr(ctl), L.stream_ptr()),
            "td_adamw_step",
        )
        if release_grads:
            for p in ps:
                p.grad = None
            if a._grad_flats is not None:
                a._grad_flats[name] = None

    @torch.no_grad()
    def launch_sharded_update(self, name: str, t: int):
        """Sharded data parallel (current stream = the caller's update stream): wait for this bucket's reduce-scatter,
        run AdamW on this rank's rows of the weight matrix, start the all-gather of the updated bf16 rows, and update the
        bucket's small vectors (all-reduced, replicated). Returns the all-gather's Work handle."""
        a = self.aligner
        dp = a._dp
        named = self._named()
        wkey = "2.weight" if name == "linear2" else "0.weight"
        W = named[wkey]
        Wb = self._bf16()[wkey]
        lo, hi = dp.shard_rows(W.shape[0])
        grads = self._grads(name)
        a.wait_bucket(name + ".big")
        st = self.state[W]
        if "exp_avg" not in st or st["exp_avg"].shape[0] != hi - lo:
            st["exp_avg"] = torch.zeros((hi - lo, W.shape[1]), dtype=torch.float32, device=W.device)
            st["exp_avg_sq"] = torch.zeros((hi - lo, W.shape[1]), dtype=torch.float32, device=W.device)
            st["shard_rows"] = (lo, hi)
        st["step"] = t
        g = self._group_of(W)
        one = lambda x: (C.c_void_p * 1)(x)  # noqa: E731
        L.launch_count += 1
        L.check(
            L.lib().td_adamw_step(1, one(W.data[lo:hi].data_ptr()), one(grads[wkey][lo:hi].data_ptr()), one(st["exp_avg"].data_ptr()),
                                  one(st["exp_avg_sq"].data_ptr()), one(Wb[lo:hi].data_ptr()), (C.c_int64 * 1)((hi - lo) * W.shape[1]),
                                  (C.c_float * 1)(g["weight_decay"]), g["lr"], g["betas"][0], g["betas"][1], g["eps"], t,
                                  self.grad_scale, None, L.stream_ptr()),
            "td_adamw_step",
        )
        ag = dp.all_gather_rows_async(Wb)
        W.grad = None
        a._grad_flats[name] = None
        return ag

    @torch.no_grad()
    def launch_peer_update(self, name: str, t: int, epoch: int):
        """Peer data parallel (current stream = the caller's update stream): wait until every rank has stored its slot of this
        weight's gradient rows into this rank's exchange buffer, run AdamW on the summed slots (bf16 rows go straight into every
        rank's compute copy) and add 1 to the "weights written" counter everywhere. The Linear1 update also carries the three
        replicated vectors: their slots arrived under the same GRAD1 counter, and they are updated BEFORE the W1 signal, so the
        next forward's single wait on W1 covers b1 / b2 / g as well."""
        from .peer import ROW_GRAD1, ROW_GRAD2, ROW_W1, ROW_W2

        a = self.aligner
        px = a._ensure_peer()
        which = 1 if name == "linear1" else 2
        W = self._named()["0.weight" if which == 1 else "2.weight"]
        lo, hi = a._dp.shard_rows(W.shape[0])
        st = self.state[W]
        if "exp_avg" not in st or st["exp_avg"].shape[0] != hi - lo:
            st["exp_avg"] = torch.zeros((hi - lo, W.shape[1]), dtype=torch.float32, device=W.device)
            st["exp_avg_sq"] = torch.zeros((hi - lo, W.shape[1]), dtype=torch.float32, device=W.device)
            st["shard_rows"] = (lo, hi)
        st["step"] = t
        g = self._group_of(W)
        px.wait(ROW_GRAD1 if which == 1 else ROW_GRAD2, epoch)
        px.adamw_rows(which, W.data[lo:hi], st["exp_avg"], st["exp_avg_sq"], g["weight_decay"], g["lr"], g["betas"], g["eps"], t,
                      self.grad_scale)
        if which == 1:
            small, d = a._grad_flats["small"], a.hidden_size
            px.sum_small(small)  # every rank's posted [db2 | dg | db1], summed in rank order in place of the local values
            keys = ["2.bias", "3.weight", "0.bias"]
            self._update_tensors(keys, t, {"2.bias": small[:d], "3.weight": small[d : 2 * d], "0.bias": small[2 * d :]})
            named = self._named()
            for k in keys:
                named[k].grad = None
            a._grad_flats["small"] = None
        px.signal(ROW_W1 if which == 1 else ROW_W2)

    @torch.no_grad()
    def launch_small_update(self, t: int):
        """Sharded data parallel: the three small vectors (b2, g, b1) are all-reduced together and updated on every rank."""
        a = self.aligner
        small, d = a._grad_flats["small"], a.hidden_size
        grads = {"2.bias": small[:d], "3.weight": small[d : 2 * d], "0.bias": small[2 * d :]}  # GradBuckets layout
        a.wait_bucket("small")
        keys = ["2.bias", "3.weight", "0.bias"]
        self._update_tensors(keys, t, grads)
        named = self._named()
        for k in keys:
            named[k].grad = None
        a._grad_flats["small"] = None

    @torch.no_grad()
    def _update_tensors(self, keys, t: int, grads):
        a = self.aligner
        named = self._named()
        bf16 = self._bf16()
        ps = [named[k] for k in keys]
        n = len(ps)
        if n == 0:
            return
        arr = lambda ptrs: (C.c_void_p * n)(*ptrs)  # noqa: E731
        states = []
        for p in ps:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] = t
            states.append(st)
        groups = [self._group_of(p) for p in ps]
        g0 = groups[0]
        L.launch_count += 1
        L.check(
            L.lib().td_adamw_step(
                n, arr([p.data_ptr() for p in ps]), arr([grads[k].data_ptr() for k in keys]),
                arr([s["exp_avg"].data_ptr() for s in states]), arr([s["exp_avg_sq"].data_ptr() for s in states]),
                arr([bf16[k].data_ptr() if k in bf16 else None for k in keys]),
                (C.c_int64 * n)(*[p.numel() for p in ps]), (C.c_float * n)(*[g["weight_decay"] for g in groups]),
                g0["lr"], g0["betas"][0], g0["betas"][1], g0["eps"], t, self.grad_scale, None, L.stream_ptr()),
            "td_adamw_step",
        )

    def next_step_number(self) -> int:
        self._t += 1
        return self._t

    def mark_bf16_current(self):
        a = self.aligner
        ps = (a[0].weight, a[0].bias, a[2].weight, a[2].bias)
        a._cache_key = tuple((p.data_ptr(), p._version) for p in ps)
        a._bf16_fresh = True

    @torch.no_grad()
    def step(self, closure=None, ctl=None):
        """``ctl``: device control block of a ``DeviceGradScaler`` -- skip / unscale / clip / bias corrections are then read on
        the device (the host-side step counter still advances; it is only used when ``ctl`` is None)."""
        loss = closure() if closure is not None else None
        self._t += 1
        order = ("linear1", "linear2") if self.aligner._bwd_order == "linear1_first" else ("linear2", "linear1")
        self.step_bucket(order[0], ctl=ctl)  # its all-reduce finished first; this update overlaps the other bucket's all-reduce
        self.step_bucket(order[1], ctl=ctl)
        a = self.aligner
        ps = (a[0].weight, a[0].bias, a[2].weight, a[2].bias)
        a._cache_key = tuple((p.data_ptr(), p._version) for p in ps)
        a._bf16_fresh = True
        return loss

    # -- checkpointing (the reference runner saves optimizer.state_dict() on rank 0 and restores it on every rank,
    #    thinkdiff/runners/runner_base.py:613, :662)
    def state_dict(self):
        """torch.optim.AdamW-compatible state: full-shape ``exp_avg`` / ``exp_avg_sq`` and a ``step`` tensor per parameter. In
        the sharded / peer modes every rank holds the moments of its own rows only, so this is a COLLECTIVE there (the row
        blocks are all-gathered): call it on all ranks, as ``AlignerTrainStep.flush()``."""
        dp = self.aligner._dp
        shards = {}
        for p, st in self.state.items():
            if "shard_rows" in st:
                shards[p] = st
        if shards:
            import torch.distributed as dist

            for p, st in shards.items():
                full = {}
                for k in ("exp_avg", "exp_avg_sq"):
                    buf = torch.empty(p.shape, dtype=torch.float32, device=p.device)
                    dist.all_gather_into_tensor(buf, st[k].contiguous(), group=dp.group)
                    full[k] = buf
                st["_full"] = full
        sd = super().state_dict()
        # super() packed the per-parameter dicts by index: patch in full-shape moments, tensor steps, drop private keys
        packed = sd["state"]
        index = {}
        i = 0
        for g in self.param_groups:
            for p in g["params"]:
                index[id(p)] = i
                i += 1
        for p, st in self.state.items():
            ent = dict(packed.get(index[id(p)], {}))
            if "_full" in st:
                ent.update(st["_full"])
            ent.pop("_full", None)
            ent.pop("shard_rows", None)
            if "step" in ent and not torch.is_tensor(ent["step"]):
                ent["step"] = torch.tensor(float(ent["step"]), dtype=torch.float32)
            packed[index[id(p)]] = ent
        for st in shards.values():
            st.pop("_full", None)
        sd["fused_adamw_step"] = self._t
        return sd

    def load_state_dict(self, state_dict):
        """Accepts this class's state and ``torch.optim.AdamW``'s (tensor ``step``). The step counter resumes from the
        checkpoint (bias corrections continue where they stopped); full-shape moments are re-sliced to this rank's rows in
        the sharded / peer modes."""
        sd = dict(state_dict)
        t = sd.pop("fused_adamw_step", None)
        super().load_state_dict(sd)
        steps = []
        dp = self.aligner._dp
        sharded = dp is not None and dp.world > 1 and (dp.sharded or dp.peer)
        for p, st in self.state.items():
            if "step" in st:
                steps.append(int(float(st["step"])))
                st["step"] = int(float(st["step"]))
            if sharded and p.dim() == 2 and "exp_avg" in st and st["exp_avg"].shape == p.shape:
                lo, hi = dp.shard_rows(p.shape[0])
                st["exp_avg"] = st["exp_avg"][lo:hi].clone()
                st["exp_avg_sq"] = st["exp_avg_sq"][lo:hi].clone()
                st["shard_rows"] = (lo, hi)
        self._t = int(t) if t is not None else (max(steps) if steps else 0)


class DeviceGradScaler:
    """``torch.amp.GradScaler`` (installed by the reference runner whenever ``amp: True``, thinkdiff/runners/runner_base.py
    :131-139) with its state on the device and no host synchronisation: the loss scale, the growth tracker, the applied-step
    count and the per-step decisions live in one small control block (``td_step_ctl_*``) that the backward's epilogues and
    the AdamW kernels read directly. ``enabled=False`` keeps the scale at 1 and never skips (used for clipping /
    accumulation without a scaler)."""

    def __init__(self, device, init_scale: float = 65536.0, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True, applied_steps: int = 0):
        self.enabled = bool(enabled)
        self.growth_factor, self.backoff_factor, self.growth_interval = float(growth_factor), float(backoff_factor), int(growth_interval)
        n = int(L.lib().td_step_ctl_bytes())
        self._block = torch.zeros((n // 4,), dtype=torch.float32, device=device)
        self.ctl = self._block
        L.check(L.lib().td_step_ctl_init(L.ptr(self._block), init_scale if self.enabled else 1.0, int(applied_steps), L.stream_ptr()),
                "td_step_ctl_init")
        # the scale as a 1-element view: passed to the backward as its upstream scalar
        self.scale_tensor = self._block[L.STEP_CTL_SCALE_OFFSET // 4 : L.STEP_CTL_SCALE_OFFSET // 4 + 1]
        self.stats = torch.zeros((4,), dtype=torch.float32, device=device)

    def begin_step(self):
        self.stats.zero_()

    def accumulate_stats(self, *grad_flats):
        """Non-finite count and sum of squares of the (scaled, reduced) gradients of this step."""
        for f in grad_flats:
            L.launch_count += 1
            L.check(L.lib().td_grad_stats(L.ptr(f), f.numel(), L.ptr(self.stats), L.stream_ptr()), "td_grad_stats")

    def update(self, optimizer, max_grad_norm: float = 0.0):
        """``scaler.unscale_`` + ``clip_grad_norm_`` + the skip decision of ``scaler.step`` + ``scaler.update``, one tiny kernel."""
        g = optimizer.param_groups[0]
        L.launch_count += 1
        L.check(L.lib().td_step_ctl_update(L.ptr(self._block), L.ptr(self.stats), int(self.enabled), self.growth_factor,
                                           self.backoff_factor, self.growth_interval, g["betas"][0], g["betas"][1],
                                           float(max_grad_norm), L.stream_ptr()), "td_step_ctl_update")

    # host-side views (these DO synchronise; for logging / tests / checkpoints)
    def get_scale(self) -> float:
        return float(self.scale_tensor)

    def state(self) -> dict:
        b = self._block.cpu()
        i = b.view(torch.int32)
        return {"skip": int(i[0]), "grad_mult": float(b[1]), "bias_c1": float(b[2]), "sqrt_bias_c2": float(b[3]), "step": float(b[4]),
                "scale": float(b[5]), "growth_tracker": int(i[6]), "grad_norm": float(b[7])}
