"""Data-parallel sharding of a global batch of sequences (one process per GPU, pure data parallel).

The reference gives each rank its own per-GPU batch (``batch_size`` per GPU, configs/train_thinkdiff_lvlm_ccsbu.yaml:36;
``seed + rank``, train.py:52-57). A fixed global batch (BASELINE configs 3-5) is split evenly by sequence; the loss of a
rank is the mean over ITS valid tokens and gradients are averaged with equal weight per rank (DDP semantics,
thinkdiff/runners/runner_base.py:88-92) -- not a global token mean.

Synchronous data parallel runs in lockstep: every step costs what the rank with the MOST tokens costs. With ragged sequences
(len ~ U{1..256}, 64 per rank) a contiguous split leaves the per-rank token counts about +-7 % apart, which alone caps the
8-GPU efficiency near 0.91. ``balanced_assignment`` keeps the per-rank sequence COUNT equal (the reference's per-GPU batch size)
and evens out the token counts ("optional length-balanced assignment is allowed").
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


def balanced_assignment(lens, world: int) -> list[list[int]]:
    """Sequence indices per rank: equal counts (as ``shard_sizes``), token sums as even as a greedy pass gets them.

    Longest-first greedy onto the rank with the fewest tokens that still has a free slot, then one pass of pairwise swaps
    between the heaviest and the lightest rank. Deterministic (ties by index); each rank's list is returned in ascending
    index order, so a rank's batch keeps the global order."""
    lens = [int(x) for x in lens]
    n = len(lens)
    cap = shard_sizes(n, world)
    order = sorted(range(n), key=lambda i: (-lens[i], i))
    groups, load = [[] for _ in range(world)], [0] * world
    for i in order:
        r = min((r for r in range(world) if len(groups[r]) < cap[r]), key=lambda r: (load[r], r))
        groups[r].append(i)
        load[r] += lens[i]
    for _ in range(4 * n):  # refine: swap one sequence between the heaviest and the lightest rank while that helps
        hi = max(range(world), key=lambda r: (load[r], -r))
        lo = min(range(world), key=lambda r: (load[r], r))
        gap = load[hi] - load[lo]
        best = None
        for a in groups[hi]:
            for b in groups[lo]:
                d = lens[a] - lens[b]
                if 0 < d < gap and (best is None or abs(gap - 2 * d) < best[0]):
                    best = (abs(gap - 2 * d), a, b)
        if best is None or best[0] >= gap:
            break
        _, a, b = best
        groups[hi].remove(a), groups[lo].remove(b)
        groups[hi].append(b), groups[lo].append(a)
        load[hi] -= lens[a] - lens[b]
        load[lo] += lens[a] - lens[b]
    return [sorted(g) for g in groups]


def contiguous_assignment(num_seqs: int, world: int) -> list[list[int]]:
    return [list(range(*shard_bounds(num_seqs, world, r))) for r in range(world)]
