"""Synthetic batches of the BASELINE.json configs. Pure torch, no CUDA library: this file is also loaded
by path (``importlib``) from ``bench.py --impl reference``, so that the CPU arm never maps ``libthinkdiff_b200.so``.

cfg 2 / 3 / 5 (ThinkDiff-LVLM): kept length ``len_i ~ U{1..max_len}`` (the injected split point of the reference collater,
thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:114), source length ``L_i = len_i + 1 + (i mod 32)``
(keeps the collater's ``1 <= split <= L_i - 1`` precondition), bf16 N(0,1) features ``[sum L_i, din]`` and, for the MSE loss,
bf16 N(0,1) targets ``[sum L_i, d]`` in the same ragged source layout.
"""
from __future__ import annotations

import torch


def lvlm_batch_tensors(num_seqs: int, max_len: int, din: int, d: int, seed: int, with_target: bool = True, truncated: bool = False):
    """Returns ``(flat, src_row_start, lens, flat_target | None)`` as host tensors.

    ``truncated``: the source holds only the kept rows of every sample (what ``FlatCollater(truncate_on_host=True)`` ships:
    the reference collater's ``[:split_point]`` done in the DataLoader worker), so ``src_row_start = cumsum(lens)``."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, max_len + 1, (num_seqs,), generator=g, dtype=torch.int32)
    full = lens.to(torch.int64) + 1 + (torch.arange(num_seqs) % 32)
    start = torch.zeros(num_seqs, dtype=torch.int64)
    start[1:] = torch.cumsum(full[:-1], 0)
    rows = int(full.sum())
    flat = torch.randn((rows, din), generator=g, dtype=torch.float32).to(torch.bfloat16)
    target = torch.randn((rows, d), generator=g, dtype=torch.float32).to(torch.bfloat16) if with_target else None
    if truncated:
        keep = torch.cat([torch.arange(int(s), int(s) + int(n)) for s, n in zip(start.tolist(), lens.tolist())])
        flat = flat[keep].contiguous()
        if target is not None:
            target = target[keep].contiguous()
        start = torch.zeros(num_seqs, dtype=torch.int64)
        start[1:] = torch.cumsum(lens[:-1].to(torch.int64), 0)
    return flat, start, lens, target
