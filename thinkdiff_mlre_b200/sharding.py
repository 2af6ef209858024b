"""Data-parallel sharding of a global batch of sequences (one process per GPU, pure data parallel).

The reference gives each rank its own per-GPU batch (``batch_size`` per GPU, configs/train_thinkdiff_lvlm_ccsbu.yaml:36;
``seed + rank``, train.py:52-57). A fixed global batch (BASELINE configs 3-5) is split evenly by sequence; the loss of a
rank is the mean over ITS valid tokens and gradients are averaged with equal weight per rank (DDP semantics,
thinkdiff/runners/runner_base.py:88-92) -- not a global token mean.
"""
from __future__ import annotations


def shard_bounds(num_seqs: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-even split: the first ``num_seqs % world`` ranks get one extra sequence."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(num_seqs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(num_seqs: int, world: int) -> list[int]:
    return [shard_bounds(num_seqs, world, r)[1] - shard_bounds(num_seqs, world, r)[0] for r in range(world)]
