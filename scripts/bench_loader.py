#!/usr/bin/env python
"""Loader measurement for the flat shard format, CPU only: time to produce one training batch
(64 samples, Qwen2-VL width 3584, config-2 length distribution) ready for H2D, from the page cache:
  * flat shard: EmbedShardReader.batch()  (one slab copy, no unpickling, no padding)
  * reference-style: per-sample torch.load of pickled tensors (what wds.decode does for the .pth entries written at
    thinkdiff/tasks/image_text_process_data.py:111-116) + the collater's pad/stack/mask (oracle restatement of
    llava_instruct_dataset_mllama_embed_2.py:101-131)
    python scripts/bench_loader.py [--out profiles/r01_loader_bench.json]"""
import argparse
import io
import json
import os
import random
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import thinkdiff_mlre_b200 as td  # noqa: E402
from oracle import pack_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--batches", type=int, default=8)
    args = ap.parse_args()
    B, C = 64, 3584
    n = B * args.batches
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(2, 258, (n,), generator=g).tolist()
    bi = dict(use_input_embed=0, use_output_embed=1, random_split_output_embed=1, output_embed_max_split_len=128,
              output_embed_max_len=256, input_embed_max_len=256)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "bench.tdemb")
    pickles = []
    with td.EmbedShardWriter(path, C) as w:
        for L in lens:
            e = torch.randn((L, C), generator=g).to(torch.bfloat16)
            ids = list(range(L))
            w.add(e, ids, "text")
            buf = io.BytesIO()
            torch.save(e.clone(), buf)
            pickles.append((buf.getvalue(), ids))
    by_threads, slab_prefaulted = {}, {}
    for threads, bound in ((1, True), (2, True), (4, True), (8, True), (8, False)):
        r = td.EmbedShardReader(path, copy_threads=threads, pin_copy_threads=bound)
        r.batch(0, B, bi, pin_memory=False)  # touch the pages
        if bound:
            random.seed(0)
            t0 = time.perf_counter()
            rows = 0
            for fb in r.batches(B, bi, pin_memory=False):
                rows += fb.flat.shape[0]
            by_threads[threads] = (time.perf_counter() - t0) / args.batches
        # the slab copy alone into a PRE-FAULTED destination -- what the pinned ring of the GPU path is (batches() above
        # allocates a fresh pageable buffer per batch, so its time includes ~14 k first-touch faults per batch)
        r1 = int(r.row_start[B])
        dst = np.empty((r1, C), dtype=np.uint16)
        dst[:] = 0
        ts = []
        for _ in range(12):
            t0 = time.perf_counter()
            r._slab_copy(dst, 0, r1)
            ts.append(time.perf_counter() - t0)
        slab_prefaulted[f"{threads}{'' if bound else '_unbound'}"] = {"first_ms": ts[0] * 1e3, "median_ms": sorted(ts)[len(ts) // 2] * 1e3}
        r.close()
    # shuffled epochs through EmbedShardSet: batches of arbitrary samples (one row-range copy per sample), with and without
    # host-side truncation to the kept rows, into a pre-faulted buffer is not possible through the public API on a CPU-only host,
    # so these include the fresh pageable destination like by_threads above
    ss = td.EmbedShardSet([path])
    shuffled = {}
    for trunc in (False, True):
        random.seed(0)
        t0 = time.perf_counter()
        nb = sum(1 for _ in ss.batches(B, bi, seed=1, epoch=0, pin_memory=False, truncate_on_host=trunc))
        shuffled["kept_rows_only" if trunc else "full_samples"] = (time.perf_counter() - t0) / nb * 1e3
    ss.close()
    t_flat = by_threads[1]
    random.seed(0)
    t0 = time.perf_counter()
    for b in range(args.batches):
        embeds, ids = [], []
        for blob, i in pickles[b * B : (b + 1) * B]:
            embeds.append(torch.load(io.BytesIO(blob)).view(torch.int16).numpy().view(np.uint16))
            ids.append(i)
        split = pack_ref.draw_split_points([e.shape[0] for e in embeds], 128, rng=random)
        pack_ref.collate_padded(embeds, "random_split", split_points=split, token_ids=ids)
    t_ref = (time.perf_counter() - t0) / args.batches
    mb = rows / args.batches * C * 2 / 1e6
    res = {"batch": B, "width": C, "mean_source_MB_per_batch": mb, "flat_shard_ms_per_batch": t_flat * 1e3,
           "flat_shard_GBps": mb / 1e3 / t_flat, "reference_style_ms_per_batch": t_ref * 1e3, "speedup": t_ref / t_flat,
           "cores_used": 1, "flat_shard_ms_per_batch_by_copy_threads": {str(k): v * 1e3 for k, v in by_threads.items()},
           "slab_copy_into_prefaulted_buffer_by_copy_threads": slab_prefaulted,
           "shuffled_epoch_ms_per_batch": shuffled,
           "host_cpus": len(os.sched_getaffinity(0))}
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
