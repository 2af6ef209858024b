"""Flat embedding shards: the on-disk format + loader that feed ``td_pack_varlen``.

Reference being replaced: the pre-compute task writes every sample's hidden states as a pickled ``torch.save`` tensor
inside WebDataset tar shards (thinkdiff/tasks/image_text_process_data.py:94-118, 500 MB shards) and the training set
reads them back through ``wds.tarfile_to_samples`` + ``wds.decode`` in Python workers
(thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:15-22), then pads/stacks on the CPU.

Here a shard is ONE flat, mmap-able file: a fixed header, the per-sample lengths, the token ids, a small JSON blob for the
texts, and all embedding rows back to back (bf16, 4096-byte aligned). A batch of consecutive samples is a single contiguous
slab of rows: ``EmbedShardReader.batch()`` copies it once into pinned memory and returns the ``FlatBatch`` that
``pack_batch`` / ``AlignerTrainStep.prefetch`` send to the GPU -- no unpickling, no per-sample tensors, no padding.

    header (64 bytes, little endian): magic "TDEMB1\\0\\0" | u32 version | u32 dtype (1 = bf16) | u32 width | u32 n_samples |
                                      u64 total_rows | u64 off_lens | u64 off_ids_index | u64 off_ids | u64 off_meta | u64 off_rows
    lens       int32 [n_samples]      full length L_i of every sample (rows)
    ids_index  int64 [n_samples + 1]  prefix offsets into ids
    ids        int32 [...]            output_token_ids of every sample, back to back
    meta       u64 length + UTF-8 JSON {"generated_text": [...], "keys": [...]} (+ "gpt" / "revised_generated_text": [...] when the
               samples carry those pass-through fields of the reference json, llava_instruct_dataset_mllama_embed_2.py:38-46)
    rows       bf16  [total_rows, width]
"""
from __future__ import annotations

import json
import mmap
import struct
import threading

import numpy as np
import torch

from .pack import FlatBatch, kept_lengths

MAGIC = b"TDEMB1\0\0"
# json fields the reference collater passes through when the batch's first sample has them -> key of the collated dict
PASSTHROUGH_JSON_KEYS = {"gpt": "llava_gpts", "revised_generated_text": "revised_generated_texts"}
_HEADER = struct.Struct("<8sIIIIQQQQQQ")
_ALIGN = 4096


def _align(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class EmbedShardWriter:
    """``add()`` samples, ``close()`` finishes the shard. Embeddings are kept as raw 16-bit words (bit-exact). The rows are
    streamed to ``<path>.rows.tmp`` as they arrive (a 500 MB shard never sits in memory); ``close()`` writes the header, the
    index and the metadata -- whose sizes are only known then -- and appends the rows behind them."""

    def __init__(self, path: str, width: int):
        self.path, self.width = path, int(width)
        self._lens, self._ids, self._texts, self._keys = [], [], [], []
        self._passthrough = {k: [] for k in PASSTHROUGH_JSON_KEYS}
        self._rows_path = path + ".rows.tmp"
        self._rows_file = open(self._rows_path, "wb")

    def add(self, embed: torch.Tensor, token_ids, generated_text: str = "", key: str = "", passthrough: dict | None = None):
        """``passthrough``: the sample's ``gpt`` / ``revised_generated_text`` json fields, when it has them (kept as they are)."""
        if embed.dtype != torch.bfloat16 or embed.dim() != 2 or embed.shape[1] != self.width:
            raise ValueError(f"expected a bfloat16 [L, {self.width}] embedding, got {embed.dtype} {tuple(embed.shape)}")
        if self._rows_file is None:
            raise ValueError("shard already closed")
        self._rows_file.write(embed.contiguous().view(torch.int16).numpy().tobytes())
        self._lens.append(int(embed.shape[0]))
        self._ids.append(np.asarray(token_ids, dtype=np.int32))
        self._texts.append(generated_text)
        self._keys.append(key)
        for k, col in self._passthrough.items():
            col.append((passthrough or {}).get(k))

    def add_reference_sample(self, sample: dict, which: str = "output"):
        """A sample dict as the reference's webdataset pipeline yields it (keys ``json``, ``*.{which}_embed.pth``)."""
        k = [k for k in sample if f"{which}_embed" in k][0]
        js = sample["json"]
        self.add(sample[k], js["output_token_ids"], js.get("generated_text", ""), sample.get("__key__", ""),
                 {p: js[p] for p in PASSTHROUGH_JSON_KEYS if p in js})

    def close(self):
        n = len(self._lens)
        lens = np.asarray(self._lens, dtype=np.int32)
        ids_index = np.zeros(n + 1, dtype=np.int64)
        ids_index[1:] = np.cumsum([len(i) for i in self._ids])
        ids = np.concatenate(self._ids) if n else np.zeros(0, np.int32)
        meta = {"generated_text": self._texts, "keys": self._keys}
        meta.update({k: col for k, col in self._passthrough.items() if any(v is not None for v in col)})
        meta = json.dumps(meta).encode("utf-8")
        off_lens = _HEADER.size
        off_idx = off_lens + lens.nbytes
        off_ids = off_idx + ids_index.nbytes
        off_meta = off_ids + ids.nbytes
        off_rows = _align(off_meta + 8 + len(meta))
        total_rows = int(lens.sum())
        if self._rows_file is None:
            return self.path
        self._rows_file.close()
        self._rows_file = None
        import os
        import shutil

        try:
            with open(self.path, "wb") as f:
                f.write(_HEADER.pack(MAGIC, 1, 1, self.width, n, total_rows, off_lens, off_idx, off_ids, off_meta, off_rows))
                f.write(lens.tobytes()), f.write(ids_index.tobytes()), f.write(ids.tobytes())
                f.write(struct.pack("<Q", len(meta))), f.write(meta)
                f.write(b"\0" * (off_rows - f.tell()))
                with open(self._rows_path, "rb") as rows:
                    shutil.copyfileobj(rows, f, 16 << 20)
        finally:
            os.remove(self._rows_path)
        return self.path

    def abort(self):
        """Drop the partly written shard (the temporary rows file); nothing is left at ``path``."""
        import os

        if self._rows_file is not None:
            self._rows_file.close()
            self._rows_file = None
            os.remove(self._rows_path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        else:
            self.abort()


class _PinnedRing:
    """Ring of pinned staging buffers ``[rows, width]`` bf16 (3 slots, or prefetch depth + 2). A slot is refilled only after the H2D
    copies that read it have finished: whoever enqueues those copies hands their CUDA events back through
    ``extras["_h2d_enqueued"]`` (``AlignerTrainStep.prefetch`` / ``step_host`` / ``pack_batch`` do), and the next use of the slot
    waits for them on the host. Without events (a consumer that never reports) the slot's previous buffer is dropped instead of
    being overwritten. One ring can serve several readers of the same width (``EmbedShardSet``): switching shards re-pins nothing."""

    def __init__(self, width: int, slots: int = 3):
        self.width, self.nslots, self.max_rows = int(width), int(slots), 0
        self.bufs, self.events, self.reported, self.pos = [], [], [], 0

    def take(self, rows: int):
        """(bf16 tensor [rows, width] in the next slot, callback that receives the H2D events of its consumer)."""
        while len(self.bufs) < self.nslots:  # (grown by batches_prefetched: depth + 2 batches are alive at once)
            self.bufs.append(None), self.events.append(None), self.reported.append(True)
        self.pos = (self.pos + 1) % len(self.bufs)
        slot = self.pos
        if self.events[slot] is not None:
            for ev in self.events[slot]:
                ev.synchronize()
            self.events[slot] = None
        elif not self.reported[slot]:
            self.bufs[slot] = None  # copies of unknown state may still read the old buffer: leave it to its tensor
        buf = self.bufs[slot]
        self.max_rows = max(self.max_rows, rows, 1)
        if buf is None or buf.shape[0] < rows:
            # pinning is expensive (cudaHostAlloc: tens of ms for a 70 MB slab): size a new buffer for the largest batch any
            # slot has seen plus 1/8, so that ragged batches stop re-pinning after the first few
            cap = self.max_rows + (self.max_rows >> 3)
            buf = self.bufs[slot] = torch.empty((cap, self.width), dtype=torch.bfloat16).pin_memory()
        self.reported[slot] = False

        def slot_cb(events, _slot=slot):
            self.events[_slot], self.reported[_slot] = list(events), True

        return buf[:rows], slot_cb


class EmbedShardReader:
    """mmap view of a shard. ``batch(lo, hi, build_info)`` -> FlatBatch for samples [lo, hi) with the reference's kept-length
    rule (random split / fixed max; seed Python's ``random`` to replay the reference's split points)."""

    def __init__(self, path: str, copy_threads: int | None = None, pin_copy_threads: bool = True, ring: "_PinnedRing | None" = None):
        """``copy_threads``: threads that share one batch's slab copy (default: up to 8 of the CPUs this process may run on). One
        core moves 5-8 GB/s from the page cache into pinned memory -- 20 ms for a 140 MB batch, many times the GPU step it
        feeds; the copy is a plain memcpy of disjoint row ranges, which numpy runs without the GIL.
        ``pin_copy_threads``: bind every copy thread to its own CPU. Left to the scheduler, threads woken by the same waker for
        a few milliseconds of work stay stacked on the waker's CPU and the threaded copy runs at one core's speed (measured: 140 MB
        into a pre-faulted buffer, 8 threads: 22-26 ms unbound for the first ~50 calls, 3-4 ms bound from the first call; on a
        B200 box the unbound pool was no faster than one thread -- profiles/r02_bench_n1_from_shards_threaded_copy.json)."""
        import os

        try:
            cpus = sorted(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            cpus = list(range(os.cpu_count() or 1))
        if copy_threads is None:
            copy_threads = min(8, len(cpus))
        self.copy_threads = max(1, int(copy_threads))
        # evenly spaced over the allowed CPUs (neighbouring numbers are often SMT siblings of one core)
        self._copy_cpus = [cpus[(i * len(cpus)) // self.copy_threads] for i in range(self.copy_threads)] if pin_copy_threads and len(cpus) >= self.copy_threads else None
        self._pool, self._bound, self._bind_lock = None, 0, threading.Lock()
        self._f = open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        for advice in (getattr(mmap, "MADV_POPULATE_READ", 22), getattr(mmap, "MADV_WILLNEED", 3)):
            try:  # pre-fault the page tables: a batch copy is then a plain memcpy instead of ~14 k minor faults
                self._mm.madvise(advice)
                break
            except (OSError, ValueError):
                continue
        magic, ver, dtype, width, n, total, off_lens, off_idx, off_ids, off_meta, off_rows = _HEADER.unpack_from(self._mm, 0)
        if magic != MAGIC or ver != 1 or dtype != 1:
            raise ValueError(f"{path}: not a TDEMB1 bf16 shard")
        self.width, self.n_samples, self.total_rows = width, n, total
        self.lens = np.frombuffer(self._mm, dtype=np.int32, count=n, offset=off_lens)
        self.ids_index = np.frombuffer(self._mm, dtype=np.int64, count=n + 1, offset=off_idx)
        self.ids = np.frombuffer(self._mm, dtype=np.int32, count=int(self.ids_index[-1]) if n else 0, offset=off_ids)
        (mlen,) = struct.unpack_from("<Q", self._mm, off_meta)
        self._meta = json.loads(bytes(self._mm[off_meta + 8 : off_meta + 8 + mlen]).decode("utf-8"))
        self.rows = np.frombuffer(self._mm, dtype=np.uint16, count=total * width, offset=off_rows).reshape(total, width)
        self.row_start = np.zeros(n + 1, dtype=np.int64)
        self.row_start[1:] = np.cumsum(self.lens)
        self._ring = ring  # pinned staging ring (created on first use; an EmbedShardSet shares one among its readers)

    def __len__(self):
        return self.n_samples

    def token_ids(self, i: int):
        return self.ids[self.ids_index[i] : self.ids_index[i + 1]]

    def embedding(self, i: int) -> torch.Tensor:
        a = np.array(self.rows[self.row_start[i] : self.row_start[i + 1]])
        return torch.from_numpy(a.view(np.int16)).view(torch.bfloat16)

    def _bind_copy_thread(self):
        """Initializer of a copy-pool thread: take the next CPU of ``_copy_cpus`` for this thread alone."""
        import os

        if self._copy_cpus is None:
            return
        with self._bind_lock:
            i, self._bound = self._bound, self._bound + 1
        try:
            os.sched_setaffinity(threading.get_native_id(), {self._copy_cpus[i % len(self._copy_cpus)]})
        except (AttributeError, OSError):  # not Linux / CPU taken away meanwhile: the thread stays where the scheduler puts it
            pass

    def _slab_copy(self, dst: np.ndarray, r0: int, r1: int):
        """dst[:] = rows[r0:r1], split by rows over the reader's threads when the slab is large enough to pay for it."""
        n = r1 - r0
        t = self.copy_threads if n * self.width * 2 >= (8 << 20) else 1
        if t <= 1:
            dst[:] = self.rows[r0:r1]
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=self.copy_threads, thread_name_prefix="td-shard-copy", initializer=self._bind_copy_thread)
        step = -(-n // t)
        futs = [self._pool.submit(np.copyto, dst[a : min(n, a + step)], self.rows[r0 + a : r0 + min(n, a + step)])
                for a in range(0, n, step)]
        for f in futs:
            f.result()

    def _gather_copy(self, dst: np.ndarray, segs):
        """dst[o : o + n] = rows[r : r + n] for every (o, r, n) of ``segs`` (disjoint), shared by the copy threads in runs of
        consecutive segments with about equal row counts when there is enough to move."""
        total = sum(n for _, _, n in segs)
        t = self.copy_threads if total * self.width * 2 >= (8 << 20) else 1
        if t <= 1 or len(segs) < 2:
            if len(segs) == 1:
                o, r, n = segs[0]
                return self._slab_copy(dst[o : o + n], r, r + n)
            for o, r, n in segs:
                dst[o : o + n] = self.rows[r : r + n]
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=self.copy_threads, thread_name_prefix="td-shard-copy", initializer=self._bind_copy_thread)

        def run(group):
            for o, r, n in group:
                np.copyto(dst[o : o + n], self.rows[r : r + n])

        groups, cur, acc, per = [], [], 0, -(-total // t)
        for sg in segs:
            cur.append(sg)
            acc += sg[2]
            if acc >= per:
                groups.append(cur)
                cur, acc = [], 0
        if cur:
            groups.append(cur)
        for f in [self._pool.submit(run, g) for g in groups]:
            f.result()

    def _staging(self, rows: int, pin_memory: bool):
        """Destination of one batch: (bf16 tensor [rows, width], its uint16 numpy view, callback for the H2D events or None)."""
        if pin_memory and torch.cuda.is_available():
            if self._ring is None:
                self._ring = _PinnedRing(self.width)
            flat, slot_cb = self._ring.take(rows)
            return flat, flat.view(torch.int16).numpy().view(np.uint16), slot_cb
        host = np.empty((rows, self.width), dtype=np.uint16)
        return torch.from_numpy(host.view(np.int16)).view(torch.bfloat16), host, None

    def batch(self, lo: int, hi: int, build_info: dict, pin_memory: bool = True, truncate_on_host: bool = False) -> FlatBatch:
        """Samples [lo, hi): one contiguous slab of the shard."""
        return self.batch_indices(range(lo, hi), build_info, pin_memory, truncate_on_host)

    def batch_indices(self, ids, build_info: dict, pin_memory: bool = True, truncate_on_host: bool = False) -> FlatBatch:
        """The batch made of samples ``ids`` in that order (any order, e.g. a shuffled sampler's): consecutive ids are one slab
        copy, anything else one row-range copy per sample -- still no unpickling and no padding. ``truncate_on_host``: copy only
        the rows the reference's rule keeps (``FlatCollater(truncate_on_host=True)``): fewer bytes into pinned memory and over
        PCIe, the device pack then degenerates to a copy."""
        ids = [int(i) for i in ids]
        if not ids:
            raise ValueError("empty batch")
        if min(ids) < 0 or max(ids) >= self.n_samples:
            raise IndexError(f"sample index outside [0, {self.n_samples})")
        full_lens = [int(self.lens[i]) for i in ids]
        lens, l_max = kept_lengths(full_lens, build_info, "output")  # draws the split points in batch order, like the reference
        src_lens = lens if truncate_on_host else full_lens
        contiguous = not truncate_on_host and all(b == a + 1 for a, b in zip(ids, ids[1:]))
        start = np.zeros(len(ids), dtype=np.int64)
        start[1:] = np.cumsum(src_lens[:-1])
        rows = int(start[-1]) + src_lens[-1]
        flat, dst, slot_cb = self._staging(rows, pin_memory)
        if contiguous:
            self._slab_copy(dst, int(self.row_start[ids[0]]), int(self.row_start[ids[-1] + 1]))  # one slab, page cache -> staging
        else:
            self._gather_copy(dst, [(int(o), int(self.row_start[i]), int(n)) for o, i, n in zip(start, ids, src_lens) if n])
        if build_info.get("random_split_output_embed"):
            out_ids = [self.token_ids(i)[n:].tolist() for i, n in zip(ids, lens)]
        else:
            out_ids = [self.token_ids(i)[:l_max].tolist() if L > l_max else self.token_ids(i).tolist() for i, L in zip(ids, full_lens)]
        texts = self._meta["generated_text"]
        extras = {"generated_texts": [texts[i] for i in ids], "output_token_ids": out_ids,
                  "embed_key": "model.norm.output_embed", "mask_key": "output_embed_mask", "sample_ids": ids}
        for k, out_key in PASSTHROUGH_JSON_KEYS.items():  # present iff the batch's FIRST sample has the field (reference :38-46)
            col = self._meta.get(k)
            if col is not None and col[ids[0]] is not None:
                vals = [col[i] for i in ids]
                if any(v is None for v in vals):
                    raise KeyError(k)  # the reference indexes json[k] of every sample once the first one has it (:64-68)
                extras[out_key] = vals
        if slot_cb is not None:
            extras["_h2d_enqueued"] = slot_cb
        return FlatBatch(flat, torch.from_numpy(start), torch.tensor(lens, dtype=torch.int32), l_max, extras)

    def batches(self, batch_size: int, build_info: dict, drop_last: bool = True, pin_memory: bool = True):
        for lo in range(0, self.n_samples, batch_size):
            hi = min(lo + batch_size, self.n_samples)
            if hi - lo < batch_size and drop_last:
                return
            yield self.batch(lo, hi, build_info, pin_memory)

    def batches_prefetched(self, batch_size: int, build_info: dict, depth: int = 2, drop_last: bool = True, pin_memory: bool = True):
        """``batches()`` with the slab copies done by a background thread, ``depth`` (1 or 2) batches ahead of the consumer: the
        page-cache -> pinned-memory memcpy of batch i+1 (numpy releases the GIL for it) overlaps the H2D and the training step of
        batch i -- what the reference gets from DataLoader workers + PrefetchLoader (thinkdiff/datasets/datasets/dataloader_utils.py
        :45-118), without worker processes or pickling. Same batches, same order, same split points as ``batches()`` (the thread is
        the only caller of ``random`` while it runs). The pinned ring is grown to ``depth + 2`` slots -- the batch the consumer
        holds, ``depth`` queued ones and the one the thread is filling -- so a slot is never refilled while its batch is alive."""
        depth = max(1, min(int(depth), 2))
        if self._ring is None:
            self._ring = _PinnedRing(self.width)
        self._ring.nslots = max(self._ring.nslots, depth + 2)
        return _prefetched(lambda: self.batches(batch_size, build_info, drop_last, pin_memory), depth)

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        self.rows = self.lens = self.ids_index = self.ids = None
        try:
            self._mm.close()
        except BufferError:  # numpy views still alive somewhere; the OS reclaims the mapping at exit
            pass
        self._f.close()


def _prefetched(make_iter, depth: int):
    """Run ``make_iter()`` on a background thread, at most ``depth`` items ahead of the consumer. The thread stops when the
    consumer abandons the generator; an exception inside it is re-raised in the consumer."""
    import queue

    q: queue.Queue = queue.Queue(maxsize=depth)
    stop = threading.Event()
    done = object()

    def produce():
        try:
            for b in make_iter():
                while not stop.is_set():
                    try:
                        q.put(b, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(done)
        except BaseException as e:  # noqa: BLE001  (re-raised in the consumer)
            q.put(e)

    th = threading.Thread(target=produce, name="td-shard-prefetch", daemon=True)
    th.start()
    try:
        while True:
            item = q.get()
            if item is done:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        th.join(timeout=5)


class EmbedShardSet:
    """Several flat shards as one training set: shuffled, rank-partitioned batches for data-parallel training.

    Reference being replaced: ``wds.ResampledShards`` + ``wds.shuffle(1000)`` feeding a DataLoader with the dataset's collater
    (thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py:15-22, thinkdiff/runners/runner_clip_t5.py:71-79) --
    shards drawn at random, samples mixed inside a 1000-sample buffer, every rank drawing on its own. Here an epoch is a
    deterministic function of ``(seed, epoch)``: the shards are visited in a shuffled order, the samples of a shard are permuted,
    and rank ``r`` of ``world`` takes every ``world``-th sample of that permutation, cut to the same count on every rank. Every
    batch comes from one shard (the mixing radius of a 1000-sample buffer over 500 MB shards), every sample is used at most once
    per epoch, and **all ranks yield the same number of batches** -- synchronous data parallel (and the peer exchange, which
    needs every rank to keep stepping) never sees a rank run dry. ``shuffle=False`` walks the shards in order (consecutive
    samples: one slab copy per batch)."""

    def __init__(self, paths, copy_threads: int | None = None, max_open: int = 2):
        """Only the 64-byte headers are read here. A shard is opened (mmap + page-table pre-fault) when its first batch is built
        and at most ``max_open`` shards stay open; all of them stage through ONE pinned ring, so walking through hundreds of
        shards neither maps them all nor re-pins a buffer per shard."""
        self.paths = [paths] if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__") else list(paths)
        if not self.paths:
            raise ValueError("no shards")
        self.copy_threads, self.max_open = copy_threads, max(1, int(max_open))
        self.counts, widths = [], []
        for path in self.paths:
            with open(path, "rb") as f:
                head = f.read(_HEADER.size)
            if len(head) < _HEADER.size:
                raise ValueError(f"{path}: not a TDEMB1 bf16 shard")
            magic, ver, dtype, width, n = _HEADER.unpack(head)[:5]
            if magic != MAGIC or ver != 1 or dtype != 1:
                raise ValueError(f"{path}: not a TDEMB1 bf16 shard")
            self.counts.append(int(n)), widths.append(int(width))
        if len(set(widths)) != 1:
            raise ValueError("shards of different embedding widths: " + ", ".join(f"{p}: {w}" for p, w in zip(self.paths, widths)))
        self.width = widths[0]
        self._ring = _PinnedRing(self.width)
        self._open = {}  # shard index -> reader, in least-recently-used order
        self._lock, self._lookahead = threading.Lock(), None

    def __len__(self):
        return sum(self.counts)

    def reader(self, si: int) -> EmbedShardReader:
        """The (lazily opened) reader of shard ``si``; the least recently used one is closed when more than ``max_open`` are open."""
        with self._lock:
            r = self._open.pop(si, None)
            if r is None:
                while len(self._open) >= self.max_open:
                    self._open.pop(next(iter(self._open))).close()
                r = EmbedShardReader(self.paths[si], self.copy_threads, ring=self._ring)
            self._open[si] = r
            return r

    def plan(self, batch_size: int, seed: int = 0, epoch: int = 0, rank: int = 0, world: int = 1, shuffle: bool = True):
        """[(shard index, [sample ids])] of one epoch for ``rank`` -- pure index arithmetic, identical code on every rank."""
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world {world}")
        rng = np.random.RandomState((int(seed) * 1000003 + int(epoch)) % (2**32))
        order = rng.permutation(len(self.paths)) if shuffle else np.arange(len(self.paths))
        out = []
        for si in order.tolist():
            n = self.counts[si]
            perm = rng.permutation(n) if shuffle else np.arange(n)  # drawn on every rank: the streams stay in step
            per_rank = n // world // batch_size * batch_size        # same count everywhere, whole batches only
            mine = perm[rank::world][:per_rank] if shuffle else perm[rank * per_rank : (rank + 1) * per_rank]
            out += [(si, mine[b : b + batch_size].tolist()) for b in range(0, per_rank, batch_size)]
        return out

    def batches(self, batch_size: int, build_info: dict, seed: int = 0, epoch: int = 0, rank: int = 0, world: int = 1,
                shuffle: bool = True, pin_memory: bool = True, truncate_on_host: bool = False):
        plan = self.plan(batch_size, seed, epoch, rank, world, shuffle)
        shards = [si for k, (si, _) in enumerate(plan) if k == 0 or plan[k - 1][0] != si]  # in visiting order
        cur = None
        for si, ids in plan:
            if si != cur:
                cur = si
                if self._lookahead is not None:
                    self._lookahead.join()
                r = self.reader(si)
                nxt = shards[shards.index(si) + 1 : shards.index(si) + 2]
                if nxt and self.max_open >= 2:
                    # open (mmap + pre-fault = read from disk) the next shard while this one is being consumed
                    self._lookahead = threading.Thread(target=self.reader, args=(nxt[0],), name="td-shard-open", daemon=True)
                    self._lookahead.start()
            fb = r.batch_indices(ids, build_info, pin_memory, truncate_on_host)
            fb.extras["shard"] = self.paths[si]
            yield fb

    def batches_prefetched(self, batch_size: int, build_info: dict, depth: int = 2, **kw):
        """``batches(...)`` produced by a background thread, ``depth`` (1 or 2) batches ahead (see EmbedShardReader.batches_prefetched)."""
        depth = max(1, min(int(depth), 2))
        self._ring.nslots = max(self._ring.nslots, depth + 2)
        return _prefetched(lambda: self.batches(batch_size, build_info, **kw), depth)

    def close(self):
        if self._lookahead is not None:
            self._lookahead.join()
            self._lookahead = None
        with self._lock:
            for r in self._open.values():
                r.close()
            self._open = {}


# ------------------------------------------------------------------------------------------ migration from the reference format
def _split_wds_name(name: str):
    """WebDataset's member-name rule: the sample key is the path up to the FIRST dot of the last path component, the field name
    is everything after it (``dir/000123.model.norm.output_embed.pth`` -> ``("dir/000123", "model.norm.output_embed.pth")``)."""
    head, _, tail = name.rpartition("/")
    base, dot, ext = tail.partition(".")
    if not dot or not base:
        return None, None
    return (head + "/" if head else "") + base, ext


def iter_webdataset_samples(tar_paths, decode_images: bool = False):
    """The samples of the reference's pre-computed shards, in file order, as the dicts its dataset yields: tar members grouped by
    key (consecutive members of one sample, as ``wds.ShardWriter`` writes them, thinkdiff/tasks/image_text_process_data.py
    :104-118), ``json`` decoded, every ``*.pth`` field ``torch.load``-ed. The ``jpg`` field is skipped unless ``decode_images``
    (then left as raw bytes): the aligner path never looks at the image. Standard library ``tarfile`` only -- neither
    ``webdataset`` nor PIL is needed to migrate."""
    import io
    import tarfile

    def decode(key, fields):
        out = {"__key__": key}
        for ext, blob in fields.items():
            if ext == "json":
                out["json"] = json.loads(blob.decode("utf-8"))
            elif ext.endswith(".pth") or ext == "pth":
                out[ext] = torch.load(io.BytesIO(blob), map_location="cpu", weights_only=True)
            else:
                out[ext] = blob
        return out

    for path in ([tar_paths] if isinstance(tar_paths, (str, bytes)) or hasattr(tar_paths, "__fspath__") else tar_paths):
        with tarfile.open(path, "r|*") as tf:  # streamed: shards are 500 MB and may be compressed
            cur, fields = None, {}
            for m in tf:
                if not m.isfile():
                    continue
                key, ext = _split_wds_name(m.name)
                if key is None:
                    continue
                if key != cur:
                    if cur is not None:
                        yield decode(cur, fields)
                    cur, fields = key, {}
                if ext.lower() in ("jpg", "jpeg", "png") and not decode_images:
                    continue
                fields[ext] = tf.extractfile(m).read()
            if cur is not None:
                yield decode(cur, fields)


def convert_webdataset_shards(tar_paths, out_prefix: str, streams=("output", "input"), max_samples: int | None = None) -> dict:
    """Rewrite the reference's WebDataset tar shards (per-sample pickled ``torch.save`` tensors + json,
    thinkdiff/tasks/image_text_process_data.py:94-118) as flat shards: one ``<out_prefix>.<stream>.tdemb`` per embed stream
    (``output`` = ``*output_embed*``, ``input`` = ``*input_embed*``; a stream no sample has is skipped). Embeddings are copied bit
    for bit (they must be bfloat16 ``[L, C]``, as vLLM's hidden states are saved); token ids, generated text, the sample key and
    the ``gpt`` / ``revised_generated_text`` fields travel in the shard's metadata. Returns ``{stream: path}``."""
    writers, paths, n = {}, {}, 0
    try:
        for sample in iter_webdataset_samples(tar_paths):
            if max_samples is not None and n >= max_samples:
                break
            if "json" not in sample:
                raise ValueError(f"sample {sample['__key__']!r} has no json member")
            for which in streams:
                ks = [k for k in sample if f"{which}_embed" in k]
                if not ks:
                    if which in writers:
                        raise ValueError(f"sample {sample['__key__']!r} lacks the {which}_embed field the earlier samples have")
                    continue
                if which not in writers:
                    if n:
                        raise ValueError(f"sample {sample['__key__']!r} is the first one with a {which}_embed field")
                    paths[which] = f"{out_prefix}.{which}.tdemb"
                    writers[which] = EmbedShardWriter(paths[which], int(sample[ks[0]].shape[-1]))
                writers[which].add_reference_sample(sample, which)
            n += 1
        if not writers:
            raise ValueError("no sample with an input_embed / output_embed field found")
    except BaseException:
        for w in writers.values():
            w.abort()  # no half-written shard and no temporary rows file is left behind
        raise
    for w in writers.values():
        w.close()
    return paths


def _main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(prog="scripts/convert_shards.py",
                                 description="convert the reference's WebDataset embedding shards (.tar) to flat .tdemb shards")
    ap.add_argument("tars", nargs="+")
    ap.add_argument("--out", required=True, help="output prefix: writes <out>.output.tdemb and / or <out>.input.tdemb")
    ap.add_argument("--streams", default="output,input")
    ap.add_argument("--max-samples", type=int, default=None)
    a = ap.parse_args(argv)
    for which, path in convert_webdataset_shards(a.tars, a.out, tuple(x for x in a.streams.split(",") if x), a.max_samples).items():
        r = EmbedShardReader(path)
        print(f"{which}: {path}: {len(r)} samples, {r.total_rows} rows x {r.width}")
        r.close()

