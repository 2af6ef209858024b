"""CPU restatement (numpy, integer/byte arithmetic) of the ragged pad/stack/mask collater.  TEST INFRASTRUCTURE ONLY.

Reference: ``LlavaInstructMllamaEmbedDataset_2.collater``
  thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:34-185
    * random-split branch (:101-131): per sample ``split = random.randint(1, min(L_i - 1, max_split_len))`` (:114);
      keep rows [:split]; zero-pad to the batch max split (:124-127); stack; int64 mask ones-then-zeros;
      target ids = ``output_token_ids[i][split:]`` (:120)
    * fixed-max branch (:132-162): ``max_len = min(cfg_max, max_i L_i)``; truncate or zero-pad each sample to it; ids
      truncated only when the embedding was (:147-149)
    * input-embed branch (:78-99): same as fixed-max, no ids

Embeddings are handled as raw 16-bit words (``uint16`` views of bf16) or any numpy dtype: the operation is a byte
copy, so parity with the CUDA pack kernel is bit-exact.  The *packed* layout (``cu_seqlens``) is this repo's
B200-side representation of the same batch: ``packed[cu[i]:cu[i+1]] == padded[i, :len_i]``.
Parity pinning: tests/golden/collater_*.npz hold the outputs of the reference's own collater (exec'd by
oracle/make_golden.py); tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

import random

import numpy as np


def draw_split_points(full_lens, max_split_len: int, seed: int | None = None, rng: random.Random | None = None):
    """Replay the reference's split-point draw: one ``randint(1, min(L-1, max_split_len))`` per sample, batch order."""
    r = rng if rng is not None else random.Random(seed)
    out = []
    for L in full_lens:
        if L < 2:
            raise ValueError("the reference collater requires L_i >= 2 (randint(1, 0) raises)")
        out.append(r.randint(1, min(L - 1, max_split_len)))
    return out


def valid_lengths(full_lens, mode: str, *, split_points=None, max_len: int | None = None):
    """Rows kept per sample and the padded length L_max for each collater branch."""
    full_lens = [int(x) for x in full_lens]
    if mode == "random_split":
        lens = [int(s) for s in split_points]
        for s, L in zip(lens, full_lens):
            if not (1 <= s <= L):
                raise ValueError(f"split point {s} outside [1, {L}]")
        return lens, (max(lens) if lens else 0)
    if mode == "fixed_max":
        lmax = min(int(max_len), max(full_lens)) if full_lens else 0
        return [min(L, lmax) for L in full_lens], lmax
    raise ValueError(mode)


def collate_padded(embeds, mode: str, *, split_points=None, max_len=None, token_ids=None):
    """Reference-layout output: padded [B, L_max, C], int64 mask [B, L_max], and the target-id lists."""
    full_lens = [e.shape[0] for e in embeds]
    lens, lmax = valid_lengths(full_lens, mode, split_points=split_points, max_len=max_len)
    C = embeds[0].shape[1]
    padded = np.zeros((len(embeds), lmax, C), dtype=embeds[0].dtype)
    mask = np.zeros((len(embeds), lmax), dtype=np.int64)
    for i, (e, n) in enumerate(zip(embeds, lens)):
        padded[i, :n] = e[:n]
        mask[i, :n] = 1
    ids_out = None
    if token_ids is not None:
        if mode == "random_split":
            ids_out = [list(t[s:]) for t, s in zip(token_ids, lens)]
        else:
            ids_out = [list(t[:lmax]) if L > lmax else list(t) for t, L in zip(token_ids, full_lens)]
    return padded, mask, ids_out


def pack_varlen(embeds, lens):
    """Packed layout: rows of all samples back to back + int32 cu_seqlens[B+1]."""
    cu = np.zeros(len(embeds) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(np.asarray(lens, dtype=np.int64))
    C = embeds[0].shape[1] if embeds else 0
    packed = np.zeros((int(cu[-1]), C), dtype=embeds[0].dtype if embeds else np.uint16)
    for i, (e, n) in enumerate(zip(embeds, lens)):
        packed[cu[i] : cu[i + 1]] = e[:n]
    return packed, cu


def pack_from_flat(flat, src_offsets, lens):
    """Same as ``pack_varlen`` but from the flat source layout the CUDA kernel consumes:
    ``flat[src_offsets[i] : src_offsets[i] + lens[i]]`` are the kept rows of sample i."""
    cu = np.zeros(len(lens) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(np.asarray(lens, dtype=np.int64))
    packed = np.zeros((int(cu[-1]), flat.shape[1]), dtype=flat.dtype)
    for i, (o, n) in enumerate(zip(src_offsets, lens)):
        packed[cu[i] : cu[i + 1]] = flat[o : o + n]
    return packed, cu


def unpack_padded(packed, cu, lmax: int | None = None):
    lens = np.diff(cu)
    lmax = int(lens.max()) if lmax is None else lmax
    padded = np.zeros((len(lens), lmax, packed.shape[1]), dtype=packed.dtype)
    mask = np.zeros((len(lens), lmax), dtype=np.int64)
    for i, n in enumerate(lens):
        padded[i, :n] = packed[cu[i] : cu[i + 1]]
        mask[i, :n] = 1
    return padded, mask
