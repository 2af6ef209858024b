"""Drop-in replacement for the reference's ``mm_projector`` (the ThinkDiff aligner).

Reference interface mirrored here (paths relative to the reference root):
  * ``build_vision_projector(config)``  thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:41-79
    (duplicated in blip_vision_t5_decoder.py:31-61 and mllama_vllm_generate_1.py:53-83): ``config`` carries
    ``mm_projector_type``, ``mm_hidden_size``, ``hidden_size``; the result is bound to ``self.mm_projector`` (:435).
  * call sites ``self.mm_projector(x)``  ...embed_decoder_2.py:585 (train), :761/:998/:1115 (inference),
    blip_vision_t5_decoder.py:414/:641: ``x[..., Din] -> y[..., D]``.

``ThinkDiffAligner`` IS an ``nn.Sequential`` of the same four children (Linear, GELU, Linear, T5LayerNorm), so
``state_dict`` keys (``0.weight 0.bias 2.weight 2.bias 3.weight``), ``named_parameters`` (the optimiser's weight-decay
split, runners/runner_base.py:104-111), ``.modules()`` scans for ``T5LayerNorm`` (...embed_decoder_2.py:697-700),
``.to()`` and DDP all behave as with the reference object. Only ``forward`` differs: it never calls the children, it
runs the fused sm_100a kernels (two tcgen05 GEMMs with bias/GELU/sum-of-squares epilogues + one normalise pass; backward:
one norm-backward pass + three tcgen05 GEMMs) through the C ABI. There is no fallback path.

Numerical regimes:
  * fp32 parameters under ``torch.autocast('cuda', dtype=torch.bfloat16)`` (training, base_task.py:237): bf16 GEMM
    inputs, fp32 accumulate, bf16 h0/h1/h2, fp32 norm, **fp32 output**, fp32 gradients.
  * bf16 parameters (``model.to(torch.bfloat16)``, scripts/test/test_blip_vision_t5_decoder_flux_text.py:104):
    **bf16 output**, normalised value rounded to bf16 before the weight multiply (T5LayerNorm rule).
  * fp32 parameters without autocast (BASELINE config 1): fp32 semantics, computed on the tensor cores by
    error-compensated bf16x3 splitting (see ``fp32_mode``).
"""
from __future__ import annotations

import os
import re

import torch
from torch import nn

from . import ops

EPS = 1e-6
FUSED_TYPE = "mlp2x_gelu_t5_norm"


def _t5_layer_norm_cls():
    """The reference builds transformers' T5LayerNorm; subclass it when available so isinstance scans keep working."""
    try:
        from transformers.models.t5.modeling_t5 import T5LayerNorm  # noqa: WPS433

        return T5LayerNorm
    except Exception:  # transformers missing: a structurally identical holder (weight, variance_epsilon)
        class T5LayerNorm(nn.Module):
            def __init__(self, hidden_size, eps=EPS):
                super().__init__()
                self.weight = nn.Parameter(torch.ones(hidden_size))
                self.variance_epsilon = eps

        return T5LayerNorm


class DataParallelState:
    """Gradient all-reduce plan for the aligner (replaces DDP's reducer for this module, runner_base.py:88-92).

    Two flat fp32 buckets in gradient-ready order -- {dW2, db2, dg} then {dW1, db1} -- are all-reduced (SUM of
    gradients pre-scaled by 1/world = DDP's mean of per-rank means) on the process group's NCCL stream while the
    remaining backward GEMMs run on the compute stream.
    """

    def __init__(self, group=None, overlap: bool = True, defer_wait: bool = False, sharded: bool = False, peer: bool = False,
                 peer_timeout_s: float = 600.0):
        import torch.distributed as dist

        self.peer_timeout_s = float(peer_timeout_s)  # a rank that stalls longer than this (checkpointing, validation) traps the others
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.overlap = overlap
        # peer (see peer.py; validated on 2 and 8 GPUs): the row-sharded plan below with NO collective on the step -- the weight-gradient
        # GEMM epilogues store each owner's rows into its exchange buffer over NVLink, the owner's AdamW sums the slots and
        # stores the bf16 rows into every rank's compute copy; step-number flags order it. Implies sharded + defer_wait.
        self.peer = peer
        if peer:
            sharded = defer_wait = True
        # sharded (ZeRO-1 style, needs defer_wait + FusedAdamW): the two weight-gradient matrices are reduce-SCATTERED by
        # rows (rank r receives rows [r D/world, (r+1) D/world)), each rank runs AdamW on its rows only and the updated
        # bf16 rows are all-gathered; the three small vectors stay all-reduced and replicated.
        self.sharded = sharded
        if sharded and not defer_wait:
            raise ValueError("sharded=True requires defer_wait=True (the optimizer consumes the shards)")
        # a second communicator for the parameter all-gathers, so they are not queued behind the next reduce-scatter
        self.ag_group = dist.new_group(ranks=dist.get_process_group_ranks(group) if group is not None else None) if sharded and not peer and self.world > 1 else group
        # defer_wait: backward returns without ordering the compute stream after the all-reduces; the consumer of the
        # gradients (FusedAdamW, or aligner.wait_grads()) waits bucket by bucket, so the Linear2 update overlaps the
        # Linear1 all-reduce.
        self.defer_wait = defer_wait

    def all_reduce_async(self, flat):
        return self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def shard_rows(self, rows: int):
        if rows % self.world:
            raise ValueError(f"sharded data parallel needs the {rows} weight rows to divide by world size {self.world}")
        n = rows // self.world
        return self.rank * n, (self.rank + 1) * n

    def reduce_scatter_rows_async(self, mat):
        """In-place reduce-scatter of a contiguous [rows, cols] gradient: this rank's row block ends up holding the sum."""
        lo, hi = self.shard_rows(mat.shape[0])
        return self.dist.reduce_scatter_tensor(mat[lo:hi], mat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_gather_rows_async(self, mat):
        """In-place all-gather of a contiguous [rows, cols] tensor whose row block [lo, hi) is current on this rank."""
        lo, hi = self.shard_rows(mat.shape[0])
        return self.dist.all_gather_into_tensor(mat, mat[lo:hi], group=self.ag_group, async_op=True)


class GradBuckets:
    """The aligner's five gradients laid out as two flat fp32 all-reduce buckets, in gradient-ready order:
    ``linear2 = [dW2 | db2 | dg]`` (ready after the first backward GEMM) and ``linear1 = [dW1 | db1]``.
    The parameter gradients handed to autograd are views into the buckets."""

    def __init__(self, din: int, d: int, device, small_separate: bool = False):
        # small_separate (sharded data parallel): the matrices are reduce-scattered, so the three small vectors live in
        # their own flat buffer ``small = [db2 | dg | db1]`` and need a single all-reduce
        self.small = torch.empty(3 * d, dtype=torch.float32, device=device) if small_separate else None
        self.linear2 = torch.empty(d * d + (0 if small_separate else 2 * d), dtype=torch.float32, device=device)
        self.linear1 = torch.empty(d * din + (0 if small_separate else d), dtype=torch.float32, device=device)
        self.dW2, self.dW1 = self.linear2[: d * d].view(d, d), self.linear1[: d * din].view(d, din)
        if small_separate:
            self.db2, self.dg, self.db1 = self.small[:d], self.small[d : 2 * d], self.small[2 * d :]
        else:
            self.db2, self.dg = self.linear2[d * d : d * d + d], self.linear2[d * d + d :]
            self.db1 = self.linear1[d * din :]

    def flats(self):
        f = {"linear2": self.linear2, "linear1": self.linear1}
        if self.small is not None:
            f["small"] = self.small
        return f

    def in_parameter_order(self):
        return self.dW1, self.db1, self.dW2, self.db2, self.dg


class _AlignerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, W1, b1, W2, b2, g, module, out_bf16):
        # (grad mode is off inside Function.forward; this path is only taken when some parameter needs a gradient)
        W1b, b1b, W2b, b2b = module._bf16_params()
        gf = g.detach() if g.dtype == torch.float32 else g.detach().float()
        y, (h0, h1, h2, rstd) = ops.aligner_fwd(x2d, W1b, b1b, W2b, b2b, gf, module.eps, out_bf16, True)
        ctx.save_for_backward(x2d, h0, h1, h2, rstd, W2b, gf)
        ctx.module = module
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, h0, h1, h2, rstd, W2b, gf = ctx.saved_tensors
        module = ctx.module
        M, Din = x2d.shape
        D = W2b.shape[0]
        dev = x2d.device
        dp = module._dp
        if dp is not None and dp.peer:
            raise NotImplementedError("peer data parallel runs through mse_loss_packed (AlignerTrainStep(fused_loss=True))")
        scale = 1.0 / dp.world if dp is not None else 1.0
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        bwd = ops.AlignerBackward(x2d, (h0, h1, h2, rstd), W2b, gf, dy.contiguous(), grad_scale=scale)
        gb = GradBuckets(Din, D, dev, small_separate=dp is not None and dp.sharded and dp.world > 1)
        module._grad_flats = gb.flats()
        phase2 = lambda: bwd.norm_and_linear2(gb.dW2, gb.db2, gb.dg)  # noqa: E731
        phase1 = lambda: bwd.gelu_and_linear1(gb.dW1, gb.db1)  # noqa: E731
        if dp is not None and dp.sharded and dp.world > 1:
            _sharded_schedule(module, gb, [("linear1", phase1), ("linear2", phase2)] if module._bwd_order == "linear1_first"
                              else [("linear2", phase2), ("linear1", phase1)])
        else:
            _bucket_schedule(module, gb, phase2, phase1)
        return (None, *gb.in_parameter_order(), None, None)


def _bucket_schedule(module, gb, run_linear2, run_linear1):
    """Backward schedule with two flat buckets: first phase -> start its all-reduce -> second phase (overlapping the first
    all-reduce) -> all-reduce the second bucket -> stream-level waits (or, with ``defer_wait``, leave the waits to the consumer
    of each bucket). ``module._bwd_order`` picks which bucket goes first: "linear2_first" is the gradient-ready order of a
    plain backward, "linear1_first" is used by the pipelined train step, where the Linear2 bucket is the one whose update can
    be deferred furthest into the next step (W2 is first read by GEMM2)."""
    dp = module._dp
    order = [("linear2", gb.linear2, run_linear2), ("linear1", gb.linear1, run_linear1)]
    if module._bwd_order == "linear1_first":
        order.reverse()
    works = {}
    order[0][2]()
    if module._record_phase_events:  # lets a side stream start on this bucket while the other phase still runs
        module._phase_done[order[0][0]] = torch.cuda.Event()
        module._phase_done[order[0][0]].record()
    if dp is not None and dp.world > 1 and dp.overlap:
        works[order[0][0]] = dp.all_reduce_async(order[0][1])  # rides NVLink while the remaining GEMMs run
    order[1][2]()
    if module._record_phase_events:
        module._phase_done[order[1][0]] = torch.cuda.Event()
        module._phase_done[order[1][0]].record()
    if dp is not None and dp.world > 1:
        if not dp.overlap:
            works[order[0][0]] = dp.all_reduce_async(order[0][1])
        works[order[1][0]] = dp.all_reduce_async(order[1][1])
        if dp.defer_wait:
            module.wait_grads()  # anything still pending from an earlier backward
            module._pending = works
        else:
            for w in works.values():
                w.wait()  # the compute stream orders after the NCCL stream; the host does not block


def _sharded_schedule(module, gb, phases, small_after_first: bool = False):
    """Sharded (ZeRO-1 style) data parallel: each weight-gradient matrix is reduce-scattered as soon as its GEMM is enqueued;
    the three small vectors share one 48 KB all-reduce -- queued right after the first phase when that phase already produced
    them (``small_after_first``: the next step's first GEMM needs b1 and must never wait for the last reduce-scatter)."""
    dp = module._dp
    # fresh view objects: a collective's Work handle keeps a reference to its tensors, and autograd only adopts a
    # returned gradient as .grad (instead of cloning it) while nobody else references that tensor object
    big = {"linear2": gb.linear2.view(gb.dW2.shape), "linear1": gb.linear1.view(gb.dW1.shape)}
    module.wait_grads()
    works = {}
    for i, (name, launch) in enumerate(phases):
        launch()
        works[name + ".big"] = dp.reduce_scatter_rows_async(big[name])
        if i == 0 and small_after_first:
            works["small"] = dp.all_reduce_async(gb.small[:])
    if "small" not in works:
        works["small"] = dp.all_reduce_async(gb.small[:])  # [db2 | dg | db1], 48 KB
    module._pending = works


def _mse_backward(module, bwd, din: int, d: int, device):
    """Backward of the fused MSE path from an ``ops.AlignerBackwardFromDh2``: runs the phases in the module's schedule, starts the
    gradient exchange, and returns the five gradients in parameter order (None for matrices that only exist in peer memory)."""
    dp = module._dp
    if dp is not None and dp.peer:
        return _peer_backward(module, bwd, d, device)
    sharded = dp is not None and dp.sharded and dp.world > 1
    gb = GradBuckets(din, d, device, small_separate=sharded)
    module._grad_flats = gb.flats()
    if module._bwd_order == "linear1_first":
        # Linear1 first: dh0 GEMM, ONE finisher for all three small vectors (+ the deferred loss), dW1 GEMM; then the dW2 GEMM
        first = lambda: bwd.gelu_linear1_and_small(gb.dW1, gb.db1, gb.db2, gb.dg)  # noqa: E731
        second = lambda: bwd.linear2_only(gb.dW2)  # noqa: E731
        if sharded:
            _sharded_schedule(module, gb, [("linear1", first), ("linear2", second)], small_after_first=True)
        else:
            _bucket_schedule(module, gb, second, first)
    else:
        phase2 = lambda: bwd.norm_and_linear2(gb.dW2, gb.db2, gb.dg)  # noqa: E731
        phase1 = lambda: bwd.gelu_and_linear1(gb.dW1, gb.db1)  # noqa: E731
        if sharded:
            _sharded_schedule(module, gb, [("linear2", phase2), ("linear1", phase1)])
        else:
            _bucket_schedule(module, gb, phase2, phase1)
    return gb.in_parameter_order()


class _AlignerMSEFn(torch.autograd.Function):
    """loss = mean((aligner(x) - target)^2) with y / dy never materialised (td_aligner_mse_fwd / td_aligner_bwd_dh2)."""

    @staticmethod
    def forward(ctx, x2d, target, W1, b1, W2, b2, g, module, target_row_index):
        W1b, b1b, W2b, b2b = module._bf16_params()
        gf = g.detach() if g.dtype == torch.float32 else g.detach().float()
        loss, saved = ops.aligner_mse_fwd(x2d, W1b, b1b, W2b, b2b, gf, module.eps, target, module._between_fwd_stages,
                                          target_row_index)
        ctx.save_for_backward(x2d, W2b, *saved)
        ctx.module = module
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        x2d, W2b, h0, h1, dh2, partials = ctx.saved_tensors
        module = ctx.module
        dp = module._dp
        scale = 1.0 / dp.world if dp is not None else 1.0
        bwd = ops.AlignerBackwardFromDh2(x2d, (h0, h1, dh2, partials), W2b, grad_loss, grad_scale=scale)
        return (None, None, *_mse_backward(module, bwd, x2d.shape[1], W2b.shape[0], x2d.device), None, None)


def _peer_backward(module, bwd, d: int, device):
    """Backward schedule of the peer-memory data-parallel mode: dh0 GEMM + one finisher for the three small vectors, which are
    posted to every rank right away; then each weight-gradient GEMM stores its rows into the owners' exchange buffers and is
    followed by a +1 on the owners' counters (GRAD1 also covers the small vectors). Returns the gradients in parameter order --
    None for the two matrices, which never exist as local tensors."""
    from .peer import ROW_GRAD1, ROW_GRAD2

    px = module._ensure_peer()
    module._peer_epoch += 1
    rot_rank = -1 if os.environ.get("TD_PEER_NO_ROTATE") == "1" else px.rank  # developer A/B: one common tile order on all ranks
    small = torch.empty(3 * d, dtype=torch.float32, device=device)  # [db2 | dg | db1], GradBuckets' small layout
    db2, dg, db1 = small[:d], small[d : 2 * d], small[2 * d :]
    module._grad_flats = {"small": small}
    if os.environ.get("TD_PEER_UNFOLDED") == "1":  # developer A/B: the protocol's tiny launches as separate kernels
        bwd.gelu_and_small(db1, db2, dg)
        px.post_small(small)
        bwd.linear1_only_scatter(px.dw_dst(1), px.world, rot_rank)
        px.signal(ROW_GRAD1)
        bwd.linear2_only_scatter(px.dw_dst(2), px.world, rot_rank)
        px.signal(ROW_GRAD2)
        return None, db1, None, db2, dg
    # 4 launches: dh0 GEMM, finisher (stores the small vectors into this rank's slot at every rank as it writes them), dW1 GEMM
    # (+1 on GRAD1 everywhere once all of its stores have landed), dW2 GEMM (+1 on GRAD2)
    bwd.gelu_and_small_scatter(db1, db2, dg, px.world, px.fold_post_small(small))
    bwd.linear1_only_scatter(px.dw_dst(1), px.world, rot_rank, px.fold_signal(ROW_GRAD1))
    bwd.linear2_only_scatter(px.dw_dst(2), px.world, rot_rank, px.fold_signal(ROW_GRAD2))
    return None, db1, None, db2, dg


class ThinkDiffAligner(nn.Sequential):
    """``mlp2x_gelu_t5_norm`` aligner running on hand-written sm_100a kernels. See the module docstring."""

    def __init__(self, mm_hidden_size: int, hidden_size: int, eps: float = EPS):
        norm = _t5_layer_norm_cls()(hidden_size, eps)
        super().__init__(nn.Linear(mm_hidden_size, hidden_size), nn.GELU(), nn.Linear(hidden_size, hidden_size), norm)
        if mm_hidden_size % 64 or hidden_size % 64 or hidden_size > 4096:
            raise ValueError(
                f"ThinkDiffAligner needs mm_hidden_size ({mm_hidden_size}) and hidden_size ({hidden_size}) to be multiples of 64 "
                "and hidden_size <= 4096 (TMA boxes / one norm row per CTA); there is no fallback path"
            )
        self.mm_hidden_size, self.hidden_size, self.eps = mm_hidden_size, hidden_size, eps
        self._cache_key = None
        self._cache = None      # persistent bf16 compute copies of (W1, b1, W2, b2)
        self._bf16_fresh = False  # set by FusedAdamW: the copies were written by the optimizer step itself
        self._cache_from_training = False  # the copies were cast by a training-mode forward (never reused for inference)
        self._pending = {}      # bucket name -> in-flight all-reduce (defer_wait mode)
        self._grad_flats = None  # the flat gradient buckets of the last backward ({"linear2": ..., "linear1": ...})
        self._bwd_order = "linear2_first"   # or "linear1_first" (pipelined train step)
        self._record_phase_events = False   # pipelined train step: CUDA event after each backward phase
        self._phase_done = {}
        self._between_fwd_stages = None     # callable run between Linear1 and Linear2 of the fused-loss forward
        self._bf16_managed = False          # True: an optimizer keeps the bf16 copies current; training never re-casts
        self._dp: DataParallelState | None = None
        self._peer = None                   # PeerExchange (peer data parallel), created on first use
        self._peer_epoch = 0                # number of peer-mode backward passes so far = the value the flags carry
        self.fp32_mode = "bf16x3"

    # -- reference-compatible config property (IdentityMap has one; harmless here)
    @property
    def config(self):
        return {"mm_projector_type": FUSED_TYPE}

    # -- data parallel (replaces DDP for this module)
    def enable_data_parallel(self, group=None, overlap: bool = True, defer_wait: bool = False, sharded: bool = False,
                             peer: bool = False, peer_timeout_s: float = 600.0):
        """``peer_timeout_s``: how long a peer-mode wait may see no progress before it traps (sticky CUDA error on this rank).
        Every rank must keep stepping (or call ``AlignerTrainStep.flush()`` + a barrier) -- a rank that leaves the loop for
        longer than this, e.g. to write a checkpoint, takes the others down with it."""
        self._dp = DataParallelState(group, overlap, defer_wait, sharded, peer, peer_timeout_s)
        return self

    def _ensure_peer(self):
        """Peer data parallel: allocate / map the exchange buffers (collective, first call only) and move the bf16 compute
        copies of the two weights into this rank's buffer, where the owners of the other rows can store into them."""
        if self._peer is None:
            from .peer import PeerExchange

            w1, w2 = self[0].weight, self[2].weight
            if w1.dtype != torch.float32:
                raise TypeError("peer data parallel is a training mode: parameters must be float32 masters")
            px = PeerExchange(self.mm_hidden_size, self.hidden_size, self._dp.group, w1.device, timeout_s=self._dp.peer_timeout_s)
            self._cache = (px.w1_bf16, torch.empty_like(self[0].bias, dtype=torch.bfloat16), px.w2_bf16,
                           torch.empty_like(self[2].bias, dtype=torch.bfloat16))
            self._cache_key, self._bf16_fresh = None, False
            self._peer = px
        return self._peer

    def sync_parameters(self):
        """Sharded data parallel: all-gather the fp32 master rows every rank updated for the others (collective; call on
        all ranks before reading parameters / saving a checkpoint). No-op otherwise."""
        dp = self._dp
        if dp is None or not dp.sharded or dp.world == 1:
            return
        works = [dp.all_gather_rows_async(self[0].weight.data), dp.all_gather_rows_async(self[2].weight.data)]
        for w in works:
            w.wait()

    BUCKETS = {"linear2": ("2.weight", "2.bias", "3.weight"), "linear1": ("0.weight", "0.bias")}

    def wait_bucket(self, name: str):
        """Order the current stream after the all-reduce of one gradient bucket (no-op unless ``defer_wait``)."""
        w = self._pending.pop(name, None)
        if w is not None:
            w.wait()

    def bucket_grads(self, name: str):
        """Views (in BUCKETS[name] order) into the flat gradient bucket written by the last backward."""
        flats = self._grad_flats
        flat, small = flats[name], flats.get("small")
        d, din = self.hidden_size, self.mm_hidden_size
        if name == "linear2":
            if small is not None:
                return flat.view(d, d), small[:d], small[d : 2 * d]
            return flat[: d * d].view(d, d), flat[d * d : d * d + d], flat[d * d + d :]
        if small is not None:
            return flat.view(d, din), small[2 * d :]
        return flat[: d * din].view(d, din), flat[d * din :]

    def wait_grads(self):
        for name in list(self._pending):
            self.wait_bucket(name)

    def disable_data_parallel(self):
        self._dp = None
        return self

    def _bf16_buffers(self):
        ps = (self[0].weight, self[0].bias, self[2].weight, self[2].bias)
        if self._cache is None or any(b.shape != p.shape or b.device != p.device for b, p in zip(self._cache, ps)):
            self._cache = tuple(torch.empty(p.shape, dtype=torch.bfloat16, device=p.device) for p in ps)
            self._cache_key, self._bf16_fresh = None, False
        return self._cache

    def _bf16_params(self, allow_cache: bool = False):
        """bf16 compute copies of the Linear parameters (persistent buffers).

        Training casts on every forward, exactly as autocast does (torch's fused optimizers update parameters without
        bumping ``Tensor._version``, so a version-keyed cache would go stale) -- unless this repo's ``FusedAdamW`` wrote
        the copies itself during its step (``_bf16_fresh``). ``eval()`` + ``no_grad`` inference reuses the copies,
        keyed on (data_ptr, version) so that ``load_state_dict`` / ``.to()`` still invalidate them."""
        ps = (self[0].weight, self[0].bias, self[2].weight, self[2].bias)
        if ps[0].dtype == torch.bfloat16:
            return tuple(p.detach() for p in ps)
        bufs = self._bf16_buffers()
        if self._bf16_managed and self._cache_key is not None:
            return bufs
        key = tuple((p.data_ptr(), p._version) for p in ps)
        fresh = self._bf16_fresh and key == self._cache_key
        self._bf16_fresh = False
        # a copy cast during training may be stale even if the key still matches: torch's fused optimizers update parameters
        # in place WITHOUT bumping Tensor._version. Only copies made in eval mode (or written by FusedAdamW) are reused.
        reuse = allow_cache and key == self._cache_key and not self._cache_from_training
        if not fresh and not reuse:
            with torch.no_grad():
                for p, b in zip(ps, bufs):
                    ops.cast_to_bf16(p.detach().contiguous(), out=b)
            self._cache_key = key
            self._cache_from_training = self.training
        elif fresh:
            self._cache_from_training = False  # FusedAdamW wrote the copies in the same pass as the update
        return bufs

    def _regime(self, x):
        wdt = self[0].weight.dtype
        if wdt == torch.bfloat16:
            return "bf16_infer"
        if wdt != torch.float32:
            raise TypeError(f"aligner parameters must be float32 or bfloat16, got {wdt}")
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return "bf16_autocast"
        if x.dtype == torch.bfloat16:
            # reference: F.linear(bf16 x, fp32 W) raises a dtype error outside autocast
            raise TypeError("bfloat16 features with float32 parameters need torch.autocast('cuda', dtype=torch.bfloat16)")
        return "fp32"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("ThinkDiffAligner runs on sm_100 CUDA devices only; there is no CPU fallback")
        if x.shape[-1] != self.mm_hidden_size:
            raise ValueError(f"expected last dim {self.mm_hidden_size}, got {tuple(x.shape)}")
        if x.requires_grad:
            raise NotImplementedError("no dx path: the reference never differentiates the aligner w.r.t. its input features")
        regime = self._regime(x)
        lead = x.shape[:-1]
        x2d = x.reshape(-1, self.mm_hidden_size)
        if regime == "fp32":
            from .fp32_path import aligner_fp32

            y = aligner_fp32(self, x2d.float().contiguous())
            return y.reshape(*lead, self.hidden_size)
        x2d = x2d.to(torch.bfloat16).contiguous()
        w1, b1, w2, b2, g = self[0].weight, self[0].bias, self[2].weight, self[2].bias, self[3].weight
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in (w1, b1, w2, b2, g))
        with torch.autocast("cuda", enabled=False):
            if need_grad:
                y = _AlignerFn.apply(x2d, w1, b1, w2, b2, g, self, regime == "bf16_infer")
            else:  # inference: nothing is saved, h0 is never written
                W1b, b1b, W2b, b2b = self._bf16_params(allow_cache=not self.training)
                gf = g.detach() if g.dtype == torch.float32 else g.detach().float()
                y, _ = ops.aligner_fwd(x2d, W1b, b1b, W2b, b2b, gf, self.eps, regime == "bf16_infer", False)
        return y.reshape(*lead, self.hidden_size)

    def mse_loss_packed(self, x_packed: torch.Tensor, target: torch.Tensor, target_row_index: torch.Tensor | None = None) -> torch.Tensor:
        """``target_row_index`` (int64 [M], optional): row of ``target`` for each row of ``x_packed`` -- lets the targets stay
        in their un-packed source layout (``ops.pack_varlen(..., want_index=True)`` returns it).

        ``F.mse_loss(self.forward_packed(x_packed).float(), target.float())`` as ONE fused training path: the norm
        output y and its gradient are formed in registers, never written to HBM. Training regime only (fp32 parameters,
        bf16 compute, as under ``torch.autocast('cuda', dtype=torch.bfloat16)``); differentiable w.r.t. the parameters and
        compatible with a GradScaler (the upstream scalar is applied on the device inside the backward GEMM epilogues)."""
        if self[0].weight.dtype != torch.float32:
            raise TypeError("mse_loss_packed is the training path: parameters must be float32 masters")
        if x_packed.dim() != 2 or x_packed.shape[1] != self.mm_hidden_size or x_packed.shape[0] == 0:
            raise ValueError(f"x_packed must be [M > 0, {self.mm_hidden_size}], got {tuple(x_packed.shape)}")
        if x_packed.requires_grad:
            raise NotImplementedError("no dx path: the reference never differentiates the aligner w.r.t. its input features")
        x2d = x_packed.to(torch.bfloat16).contiguous()
        if target.dtype not in (torch.float32, torch.bfloat16):
            target = target.float()
        w1, b1, w2, b2, g = self[0].weight, self[0].bias, self[2].weight, self[2].bias, self[3].weight
        with torch.autocast("cuda", enabled=False):
            return _AlignerMSEFn.apply(x2d, target.contiguous(), w1, b1, w2, b2, g, self, target_row_index)

    @torch.no_grad()
    def mse_loss_backward_packed(self, x_packed: torch.Tensor, target: torch.Tensor, target_row_index: torch.Tensor | None = None,
                                 upstream: torch.Tensor | None = None, stats: torch.Tensor | None = None,
                                 accumulate_into=None, set_grads: bool = True) -> torch.Tensor:
        """``mse_loss_packed(...)`` followed by ``loss.backward()`` as ONE direct call sequence, without autograd: forward GEMMs,
        fused norm + loss + norm backward, backward GEMMs, gradient exchange (whatever ``enable_data_parallel`` set up). The
        gradients land in the flat buckets (``self._grad_flats``; with ``set_grads`` the parameters' ``.grad`` become views of
        them, as autograd would leave them) and the loss is returned as a device scalar. This is what ``AlignerTrainStep`` runs:
        same kernels and arithmetic as the autograd path, one launch fewer (the loss is finished by the backward's finisher)
        and no autograd bookkeeping on the host. ``upstream``: device scalar multiplied into every gradient (a GradScaler's
        scale); ``stats``: fp32 [2] non-finite counter; ``accumulate_into``: ``GradBuckets`` of an earlier micro-batch to add
        into (gradient accumulation; not with sharded / peer data parallel)."""
        if self[0].weight.dtype != torch.float32:
            raise TypeError("mse_loss_backward_packed is the training path: parameters must be float32 masters")
        if x_packed.dim() != 2 or x_packed.shape[1] != self.mm_hidden_size or x_packed.shape[0] == 0:
            raise ValueError(f"x_packed must be [M > 0, {self.mm_hidden_size}], got {tuple(x_packed.shape)}")
        x2d = x_packed.to(torch.bfloat16).contiguous()
        if target.dtype not in (torch.float32, torch.bfloat16):
            target = target.float()
        W1b, b1b, W2b, b2b = self._bf16_params()
        g = self[3].weight
        gf = g.detach() if g.dtype == torch.float32 else g.detach().float()
        loss, saved = ops.aligner_mse_fwd(x2d, W1b, b1b, W2b, b2b, gf, self.eps, target.contiguous(), self._between_fwd_stages,
                                          target_row_index, defer_loss=True)
        dp = self._dp
        scale = 1.0 / dp.world if dp is not None else 1.0
        bwd = ops.AlignerBackwardFromDh2(x2d, saved, W2b, upstream, grad_scale=scale, loss_out=loss, stats=stats,
                                         accumulate=accumulate_into is not None)
        if accumulate_into is not None:
            if dp is not None and dp.world > 1 and (dp.sharded or dp.peer):
                raise NotImplementedError("gradient accumulation with sharded / peer data parallel")
            gb = accumulate_into
            self._grad_flats = gb.flats()
            bwd.gelu_linear1_and_small(gb.dW1, gb.db1, gb.db2, gb.dg)
            bwd.linear2_only(gb.dW2)
            grads = gb.in_parameter_order()
        else:
            grads = _mse_backward(self, bwd, x2d.shape[1], W2b.shape[0], x2d.device)
        if set_grads:
            for p, gr in zip((self[0].weight, self[0].bias, self[2].weight, self[2].bias, self[3].weight), grads):
                p.grad = gr
        return loss

    def forward_packed(self, x_packed: torch.Tensor, cu_seqlens: torch.Tensor | None = None) -> torch.Tensor:
        """Ragged entry point: ``x_packed[M, Din]`` (rows of all sequences back to back, ``cu_seqlens`` int32 [B+1]).

        The aligner is token-wise, so the packed rows are projected directly -- pad rows are never computed, unlike the
        reference which projects the zero-padded ``[B, L_max, Din]`` batch (...embed_decoder_2.py:585)."""
        if x_packed.dim() != 2:
            raise ValueError("x_packed must be [M, Din]")
        if cu_seqlens is not None and int(cu_seqlens.numel()) < 1:
            raise ValueError("cu_seqlens must have B+1 entries")
        return self.forward(x_packed)


def build_vision_projector(config):
    """Same contract as the reference builder (...embed_decoder_2.py:41-79). ``mlp2x_gelu_t5_norm`` -- the type every
    shipped config uses (configs/*.yaml ``mm_projector_type``) -- returns the fused B200 module; the other valid type
    strings are recognised and rejected explicitly (no silent PyTorch fallback); unknown strings raise the reference's error."""
    projector_type = getattr(config, "mm_projector_type", "linear")
    if projector_type == FUSED_TYPE:
        return ThinkDiffAligner(config.mm_hidden_size, config.hidden_size)
    known = projector_type in ("linear", "identity") or re.match(r"^mlp(\d+)x_gelu(_t5_norm|_rms_norm|_norm)?$", projector_type)
    if known:
        raise NotImplementedError(
            f"mm_projector_type={projector_type!r} has no B200 kernel path; only {FUSED_TYPE!r} (used by every shipped config) is built"
        )
    raise ValueError(f"Unknown projector type: {projector_type}")
