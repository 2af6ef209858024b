"""Learning-rate schedules of the reference runner, as host-side plumbing for ``FusedAdamW``.

Reference: ``thinkdiff/common/optims.py:13-112`` (``linear_warmup_step_lr`` / ``linear_warmup_cosine_lr``, selected by
``run.lr_sched`` in the configs and stepped once per iteration at ``thinkdiff/tasks/base_task.py:230`` *before* the forward
pass). Same class names, constructor arguments and ``step(cur_epoch, cur_step)`` call, so the runner's
``registry.get_lr_scheduler_class(...)`` call site can bind these instead. Each schedule is one pure function ``lr_at`` (what the
parity test checks against the reference's own code) plus the write into ``optimizer.param_groups``.

``FusedAdamW`` reads ``group["lr"]`` when it enqueues an update, i.e. inside the step that produced the gradients -- also in the
pipelined / sharded modes, where the update merely *executes* during the next step -- so the value set before step ``i`` is the
one applied to step ``i``'s gradients, exactly as in the reference loop.
"""
from __future__ import annotations

import math


def _warmup(step: int, warmup_steps: int, start_lr: float, peak_lr: float) -> float:
    return min(peak_lr, start_lr + (peak_lr - start_lr) * step / max(warmup_steps, 1))


class _Schedule:
    def __init__(self, optimizer):
        self.optimizer = optimizer

    def lr_at(self, cur_epoch: int, cur_step: int) -> float:  # pragma: no cover - abstract
        raise NotImplementedError

    def step(self, cur_epoch: int, cur_step: int) -> float:
        lr = self.lr_at(cur_epoch, cur_step)
        for group in self.optimizer.param_groups:
            group["lr"] = lr
        return lr


class LinearWarmupStepLRScheduler(_Schedule):
    """Epoch 0: linear warm-up over ``warmup_steps`` iterations (clamped at ``init_lr``); afterwards
    ``max(min_lr, init_lr * decay_rate ** epoch)`` (optims.py:13-53, :107-112)."""

    def __init__(self, optimizer, max_epoch, min_lr, init_lr, decay_rate=1, warmup_start_lr=-1, warmup_steps=0, **kwargs):
        super().__init__(optimizer)
        self.max_epoch, self.min_lr, self.init_lr, self.decay_rate = max_epoch, min_lr, init_lr, decay_rate
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def lr_at(self, cur_epoch: int, cur_step: int) -> float:
        if cur_epoch == 0:
            return _warmup(cur_step, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        return max(self.min_lr, self.init_lr * (self.decay_rate ** cur_epoch))


class LinearWarmupCosineLRScheduler(_Schedule):
    """Linear warm-up while the GLOBAL iteration is below ``warmup_steps`` -- interpolated on the iteration index *within the
    epoch*, as the reference does (optims.py:81-89) -- then half-cosine from ``init_lr`` to ``min_lr`` over
    ``max_epoch * iters_per_epoch`` iterations (optims.py:90-104)."""

    def __init__(self, optimizer, max_epoch, iters_per_epoch, min_lr, init_lr, warmup_steps=0, warmup_start_lr=-1, **kwargs):
        super().__init__(optimizer)
        self.max_epoch, self.iters_per_epoch, self.min_lr, self.init_lr = max_epoch, iters_per_epoch, min_lr, init_lr
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def lr_at(self, cur_epoch: int, cur_step: int) -> float:
        it = cur_epoch * self.iters_per_epoch + cur_step
        if it < self.warmup_steps:
            return _warmup(cur_step, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        total = self.max_epoch * self.iters_per_epoch
        return (self.init_lr - self.min_lr) * 0.5 * (1.0 + math.cos(math.pi * it / total)) + self.min_lr


SCHEDULERS = {"linear_warmup_step_lr": LinearWarmupStepLRScheduler, "linear_warmup_cosine_lr": LinearWarmupCosineLRScheduler}


def get_lr_scheduler_class(name: str):
    """``registry.get_lr_scheduler_class`` of the reference (runner_base.py:141-177), for the two registered names."""
    try:
        return SCHEDULERS[name]
    except KeyError:
        raise KeyError(f"unknown lr scheduler {name!r}; the reference registers {sorted(SCHEDULERS)}") from None
