"""Ragged packing of pre-computed VLM embeddings: host-side collater + device-side pack/pad/mask kernels.

Reference being replaced: ``LlavaInstructMllamaEmbedDataset_2.collater``
(thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:34-185) followed by the blocking ``.cuda()`` of the
zero-padded batch (thinkdiff/datasets/data_utils.py:83-96). The reference pads every sample to the batch maximum on the
CPU inside DataLoader workers; pad rows are then copied to the GPU and projected by the aligner.

Here the worker-side collater (``FlatCollater``, pure CPU, safe in DataLoader worker processes -- it never touches CUDA)
only concatenates the samples' *full* embeddings into one flat (pinned) tensor and decides the kept length of each
sample with the reference's own rule (same ``random.randint`` draw, same truncation). Truncation, compaction into the
``cu_seqlens``-indexed packed buffer, and -- when a consumer needs the reference layout -- zero padding and the int64
mask are done by the CUDA kernels (``td_pack_varlen`` / ``td_pack_padded``). All outputs are bit-exact copies.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

import torch

from . import ops


@dataclass
class FlatBatch:
    """CPU-side product of the collater: everything the device needs, no padding."""

    flat: torch.Tensor            # [sum_i L_i, C] the samples' full embeddings back to back (pinned when possible)
    src_row_start: torch.Tensor   # int64 [B] first row of sample i in ``flat``
    lens: torch.Tensor            # int32 [B] rows kept for sample i (split point / truncated length)
    l_max: int                    # padded length the reference collater would have produced
    extras: dict = field(default_factory=dict)  # output_token_ids, generated_texts, ... (reference dict keys)

    @property
    def total_rows(self) -> int:
        return int(self.lens.sum())


@dataclass
class PackedBatch:
    """Device-side ragged batch: ``x[cu[i]:cu[i+1]]`` are the kept rows of sample i."""

    x: torch.Tensor               # [M, C]
    cu_seqlens: torch.Tensor      # int32 [B+1] on the device
    lens_host: torch.Tensor       # int32 [B] on the host
    l_max: int
    extras: dict = field(default_factory=dict)

    def to_padded(self):
        """Reference layout (``[B, L_max, C]`` zero padded, int64 ``[B, L_max]`` mask) from the packed rows."""
        start = self.cu_seqlens[:-1].to(torch.int64)
        return ops.pack_padded(self.x, start, self.cu_seqlens, self.l_max)


def kept_lengths(full_lens, build_info: dict, key: str = "output", rng=random):
    """Rows kept per sample + padded length, by the reference's rules.

    random split (:101-131): ``randint(1, min(L-1, output_embed_max_split_len))`` per sample, in batch order, drawn from
    Python's global ``random`` (seed it to replay); fixed max (:132-162) / input embeds (:78-99): ``min(cfg_max, max_i L_i)``.
    """
    full_lens = [int(x) for x in full_lens]
    if key == "output" and build_info.get("random_split_output_embed"):
        lens = [rng.randint(1, min(L - 1, build_info["output_embed_max_split_len"])) for L in full_lens]
        return lens, max(lens)
    cfg_max = build_info["output_embed_max_len" if key == "output" else "input_embed_max_len"]
    l_max = min(int(cfg_max), max(full_lens))
    return [min(L, l_max) for L in full_lens], l_max


class FlatCollater:
    """Drop-in for the reference collater's role in the DataLoader (``collate_fn``), emitting a ``FlatBatch``.

    ``samples`` are the reference's webdataset dicts: ``sample["json"]`` with ``generated_text`` / ``output_token_ids`` (and
    optionally ``gpt`` / ``revised_generated_text``) and the ``*output_embed*`` / ``*input_embed*`` tensors ``[L_i, C]``. The
    returned ``extras`` carry the same non-tensor keys the reference returns (``generated_texts``, ``output_token_ids`` sliced
    exactly as at :120 / :147-149, ``llava_gpts`` and ``revised_generated_texts`` when the samples have them, :38-46, :172-175).

    ``which``: "output" / "input" collates that embed stream; ``None`` (default) follows ``build_info`` like the reference
    (:50-53): the output stream when ``use_output_embed``, else the input stream, and when BOTH flags are set the output
    stream is the returned batch and the input stream rides along as ``extras["input_batch"]`` (a second ``FlatBatch``).
    ``reference_dict(batch, device)`` turns either into the exact dict of the reference collater (:169-183)."""

    def __init__(self, build_info: dict, pin_memory: bool = True, which: str | None = None, truncate_on_host: bool = False):
        if not (build_info.get("use_output_embed") or build_info.get("use_input_embed")):
            raise ValueError("No input or output embeds are used.")  # reference message (:52-53)
        if which not in (None, "output", "input"):
            raise ValueError("which must be None, 'output' or 'input'")
        self.build_info, self.pin_memory, self.which = build_info, pin_memory, which
        # truncate_on_host: copy only the kept rows of every sample into the flat buffer (what the reference collater's
        # ``[:split_point]`` does) -- the H2D then moves M rows instead of sum(L_i); the device pack degenerates to a copy
        self.truncate_on_host = truncate_on_host

    def _stream(self, samples, which: str) -> FlatBatch:
        key = [k for k in samples[0].keys() if f"{which}_embed" in k][0]
        embeds = [s[key] for s in samples]
        ids = [s["json"]["output_token_ids"] for s in samples]
        full_lens = [int(e.shape[0]) for e in embeds]
        lens, l_max = kept_lengths(full_lens, self.build_info, which)
        src_lens = lens if self.truncate_on_host else full_lens
        flat = torch.cat([e[:n] for e, n in zip(embeds, lens)] if self.truncate_on_host else embeds, dim=0)
        if self.pin_memory and torch.cuda.is_available():
            flat = flat.pin_memory()
        start = torch.zeros(len(embeds), dtype=torch.int64)
        if len(embeds) > 1:
            start[1:] = torch.cumsum(torch.tensor(src_lens[:-1], dtype=torch.int64), 0)
        if which == "output" and self.build_info.get("random_split_output_embed"):
            out_ids = [t[n:] for t, n in zip(ids, lens)]
        elif which == "output":
            out_ids = [t[:l_max] if L > l_max else t for t, L in zip(ids, full_lens)]
        else:
            out_ids = ids
        js0 = samples[0]["json"]
        extras = {"generated_texts": [s["json"]["generated_text"] for s in samples], "output_token_ids": out_ids,
                  "embed_key": key.replace(".pth", ""), "mask_key": f"{which}_embed_mask"}
        if "gpt" in js0:  # pass-through keys, present only when the first sample has them (reference :38-46)
            extras["llava_gpts"] = [s["json"]["gpt"] for s in samples]
        if "revised_generated_text" in js0:
            extras["revised_generated_texts"] = [s["json"]["revised_generated_text"] for s in samples]
        return FlatBatch(flat, start, torch.tensor(lens, dtype=torch.int32), l_max, extras)

    def __call__(self, samples) -> FlatBatch:
        bi = self.build_info
        if self.which is not None:
            return self._stream(samples, self.which)
        if bi.get("use_output_embed"):
            batch = self._stream(samples, "output")
            if bi.get("use_input_embed"):
                batch.extras["input_batch"] = self._stream(samples, "input")
            return batch
        return self._stream(samples, "input")

    @staticmethod
    def reference_dict(batch: FlatBatch, device="cuda") -> dict:
        """The reference collater's return value (:169-183) from a ``FlatBatch``: H2D of the flat source, device-side zero
        padding + int64 mask (``td_pack_padded``), and the pass-through lists. With a dual-stream batch both tensor pairs are
        present, as in the reference when ``use_input_embed`` and ``use_output_embed`` are both set."""
        ex = batch.extras
        out = {"generated_texts": ex["generated_texts"], "output_token_ids": ex["output_token_ids"]}
        for k in ("llava_gpts", "revised_generated_texts"):
            if k in ex:
                out[k] = ex[k]
        streams = [batch] + ([ex["input_batch"]] if "input_batch" in ex else [])
        for fb in streams:
            padded, mask = pack_batch(fb, device).to_padded()
            out[fb.extras["embed_key"]] = padded
            out[fb.extras["mask_key"]] = mask
        return out


def pack_batch(batch: FlatBatch, device="cuda", non_blocking: bool = True) -> PackedBatch:
    """H2D of the flat source (one async copy from pinned memory) + device-side compaction into the packed layout."""
    flat = batch.flat.to(device, non_blocking=non_blocking)
    start = batch.src_row_start.to(device, non_blocking=non_blocking)
    lens = batch.lens.to(device, non_blocking=non_blocking)
    cb = batch.extras.get("_h2d_enqueued")
    if cb is not None and flat.is_cuda:
        ev = torch.cuda.Event()
        ev.record()
        cb([ev])  # the producer of the pinned source (EmbedShardReader) may recycle it once this has fired
    return pack_device(flat, start, lens, batch.total_rows, batch.l_max, batch.lens, batch.extras)


def pack_device(flat, src_row_start, lens_dev, total_rows: int, l_max: int, lens_host=None, extras=None) -> PackedBatch:
    """Pack from tensors already resident in HBM (the timed part of the benchmark's device-only number)."""
    cu = ops.cu_seqlens(lens_dev)
    x = ops.pack_varlen(flat, src_row_start, cu, total_rows)
    return PackedBatch(x, cu, lens_host, l_max, extras or {})


def compose_image_text(image_tokens: torch.Tensor, text_flat: torch.Tensor, text_lens) -> PackedBatch:
    """Ragged composition of aligner outputs with T5 text embeddings (BASELINE config 4; layout of the reference's
    two-image demo, scripts/test/test_blip_vision_t5_decoder_flux_text.py:184-208: ``torch.cat([img1, img2, text], dim=1)``
    per sample). ``image_tokens [B, N_img, D]`` (e.g. 2 x 32 aligner tokens), ``text_flat [sum T_i, D]`` with ``text_lens[B]``
    (host ints). One gather kernel writes ``[img_i | text_i]`` for every sample back to back; ``to_padded()`` then gives
    the ``[B, L_max, D]`` prompt_embeds + int64 mask a Flux / T5 consumer takes."""
    B, n_img, D = image_tokens.shape
    text_lens = [int(t) for t in text_lens]
    if len(text_lens) != B or text_flat.shape[1] != D or text_flat.dtype != image_tokens.dtype:
        raise ValueError("text_flat / text_lens do not match image_tokens")
    dev = image_tokens.device
    text_start = [0] * B
    for i in range(1, B):
        text_start[i] = text_start[i - 1] + text_lens[i - 1]
    starts, lens = [], []
    for i in range(B):  # two segments per sample: its image rows (first source), then its text rows (second source: -(row + 1))
        starts += [i * n_img, -(text_start[i] + 1)]
        lens += [n_img, text_lens[i]]
    start_t = torch.tensor(starts, dtype=torch.int64).to(dev, non_blocking=True)
    lens_t = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
    cu = ops.cu_seqlens(lens_t)
    # one gather launch reading both tensors where they lie (no torch.cat of the aligner output and the text embeddings)
    x = ops.pack_varlen2(image_tokens.reshape(B * n_img, D).contiguous(), text_flat.contiguous(), start_t, cu, sum(lens))
    sample_lens = torch.tensor([n_img + t for t in text_lens], dtype=torch.int32)
    return PackedBatch(x, cu[::2].contiguous(), sample_lens, int(sample_lens.max()), {"n_img": n_img})
