"""Sharding of a global batch over data-parallel ranks (host logic of SURVEY section 8e; CPU only): the length-balanced
assignment bench.py uses by default at N > 1 is a partition with the reference's equal per-GPU sequence counts, never less even
than the contiguous split, deterministic, and the per-rank synthetic batches are exactly the global batch's sequences."""
import random

import pytest
import torch

import thinkdiff_mlre_b200 as td
from thinkdiff_mlre_b200.sharding import balanced_assignment, contiguous_assignment, shard_sizes
from thinkdiff_mlre_b200.synth import global_lengths, lvlm_sequences


def _spread(groups, lens):
    loads = [sum(lens[i] for i in g) for g in groups]
    return max(loads) - min(loads), max(loads)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n,max_len,seed", [(64, 256, 0), (512, 256, 1), (1024, 1024, 2), (13, 7, 3), (8, 1, 4), (5, 100, 5)])
def test_balanced_assignment_is_an_even_partition(world, n, max_len, seed):
    rng = random.Random(seed)
    lens = [rng.randint(1, max_len) for _ in range(n)]
    groups = balanced_assignment(lens, world)
    assert sorted(i for g in groups for i in g) == list(range(n))            # every sequence on exactly one rank
    assert [len(g) for g in groups] == shard_sizes(n, world)                 # the reference's equal per-GPU batch size
    assert all(g == sorted(g) for g in groups)                               # a rank's batch keeps the global order
    assert groups == balanced_assignment(list(lens), world)                  # deterministic
    contiguous = contiguous_assignment(n, world)
    assert _spread(groups, lens)[1] <= _spread(contiguous, lens)[1]          # the slowest rank is never slower than before
    if n % world == 0 and n >= 8 * world:
        # enough sequences per rank: the token counts end up within one longest sequence of each other (in practice within a few
        # tokens; a contiguous split of U{1..256} lengths is several hundred tokens apart)
        assert _spread(groups, lens)[0] <= max(lens)


def test_cfg2_global_batch_of_8_ranks_is_balanced_to_a_few_tokens():
    """The configuration bench.py times at N = 8 (BASELINE config 3: 512 sequences, len ~ U{1..256}): the contiguous split leaves
    the ranks hundreds of tokens apart (synchronous data parallel pays for the heaviest rank), the balanced one a handful."""
    lens = global_lengths(64 * 8, 256, 1234).tolist()
    gap_c, max_c = _spread(contiguous_assignment(512, 8), lens)
    gap_b, max_b = _spread(balanced_assignment(lens, 8), lens)
    mean = sum(lens) / 8
    assert gap_b <= 8 and max_b - mean <= 8
    assert gap_c > 20 * max(gap_b, 1) and max_c > max_b


@pytest.mark.parametrize("balanced", [True, False])
def test_per_rank_synthetic_batches_are_the_global_batch(balanced):
    """Every rank materialises only its own sequences, yet together they are bit for bit the sequences of the global batch
    (each sequence has its own generator seed): N ranks train on the same data one process would see."""
    world, seqs, max_len, din, d, seed = 4, 6, 20, 16, 24, 77
    lens_all = global_lengths(world * seqs, max_len, seed)
    gflat, gstart, gkeep, gtgt = lvlm_sequences(range(world * seqs), lens_all, din, d, seed)
    groups = balanced_assignment(lens_all.tolist(), world) if balanced else contiguous_assignment(world * seqs, world)
    seen = []
    for rank in range(world):
        b = td.synthetic_lvlm_batch(seqs, max_len, din, d, seed, pin=False, world=world, rank=rank, balanced=balanced)
        assert b.lens.tolist() == [int(lens_all[i]) for i in groups[rank]] and b.l_max == int(b.lens.max())
        for k, i in enumerate(groups[rank]):
            s0, g0, n = int(b.src_row_start[k]), int(gstart[i]), int(lens_all[i]) + 1 + i % 32
            assert torch.equal(b.flat[s0 : s0 + n].view(torch.int16), gflat[g0 : g0 + n].view(torch.int16))
            assert torch.equal(b.extras["flat_target"][s0 : s0 + n].view(torch.int16), gtgt[g0 : g0 + n].view(torch.int16))
        seen += groups[rank]
        # the truncated (kept rows only) layout holds the same kept rows
        t = td.synthetic_lvlm_batch(seqs, max_len, din, d, seed, pin=False, truncated=True, world=world, rank=rank, balanced=balanced)
        assert t.flat.shape[0] == int(t.lens.sum()) and t.lens.tolist() == b.lens.tolist()
        for k in range(seqs):
            a0, b0, n = int(t.src_row_start[k]), int(b.src_row_start[k]), int(b.lens[k])
            assert torch.equal(t.flat[a0 : a0 + n].view(torch.int16), b.flat[b0 : b0 + n].view(torch.int16))
    assert sorted(seen) == list(range(world * seqs))
