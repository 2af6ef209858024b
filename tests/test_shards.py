"""Flat embedding shards: write -> mmap read is bit-exact, and a batch drawn from a shard + the
reference's kept-length rule reproduces the reference collater's golden output. CPU only."""
import random

import numpy as np
import pytest
import torch

import thinkdiff_mlre_b200 as td
from oracle import pack_ref
from oracle.golden import load_golden


def _golden_samples(g):
    full = [int(v) for v in g["full_lens"]]
    off = np.concatenate([[0], np.cumsum(full)])
    out = []
    for i in range(len(full)):
        e = torch.from_numpy(g["src_bits"][off[i] : off[i + 1]].view(np.int16).copy()).view(torch.bfloat16)
        ids = [int(v) for v in g["src_ids_flat"][off[i] : off[i + 1]]]
        out.append({"__key__": f"k{i}", "json": {"generated_text": f"sample {i}", "output_token_ids": ids},
                    "model.norm.input_embed.pth": e, "model.norm.output_embed.pth": e})
    return out


@pytest.mark.parametrize("name", ["collater_random_split.npz", "collater_fixed_max.npz"])
def test_shard_roundtrip_and_batch_equals_reference_collater(tmp_path, name):
    g = load_golden(name)
    samples = _golden_samples(g)
    path = str(tmp_path / "s.tdemb")
    with td.EmbedShardWriter(path, width=int(g["C"])) as w:
        for s in samples:
            w.add_reference_sample(s)
    r = td.EmbedShardReader(path)
    assert len(r) == len(samples) and r.width == int(g["C"]) and r.total_rows == int(sum(g["full_lens"]))
    for i, s in enumerate(samples):
        assert torch.equal(r.embedding(i).view(torch.int16), s["model.norm.output_embed.pth"].view(torch.int16))
        assert r.token_ids(i).tolist() == s["json"]["output_token_ids"]
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    random.seed(int(g["seed"]))
    fb = r.batch(0, len(samples), bi, pin_memory=False)
    bits = fb.flat.view(torch.int16).numpy().view(np.uint16)
    packed, cu = pack_ref.pack_from_flat(bits, fb.src_row_start.tolist(), fb.lens.tolist())
    padded, mask = pack_ref.unpack_padded(packed, cu, fb.l_max)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    ids = [list(g["ids_flat"][g["ids_off"][i] : g["ids_off"][i + 1]]) for i in range(len(samples))]
    assert fb.extras["output_token_ids"] == [list(map(int, t)) for t in ids]
    assert fb.extras["generated_texts"] == [f"sample {i}" for i in range(len(samples))]
    # sub-range batches index rows relative to their own slab
    random.seed(0)
    fb2 = r.batch(2, 5, dict(bi, random_split_output_embed=0, output_embed_max_len=1000), pin_memory=False)
    assert fb2.src_row_start.tolist() == [0, int(g["full_lens"][2]), int(g["full_lens"][2] + g["full_lens"][3])]
    assert fb2.flat.shape[0] == int(sum(g["full_lens"][2:5]))
    r.close()


def test_shard_rejects_garbage(tmp_path):
    p = tmp_path / "bad.tdemb"
    p.write_bytes(b"\0" * 128)
    with pytest.raises(ValueError, match="TDEMB1"):
        td.EmbedShardReader(str(p))
    w = td.EmbedShardWriter(str(tmp_path / "x.tdemb"), width=8)
    with pytest.raises(ValueError):
        w.add(torch.zeros(3, 8), [1, 2, 3])  # float32, not bf16


def test_prefetched_batches_equal_plain_batches_in_order(tmp_path):
    """batches_prefetched(): the slab copies move to a background thread; batches, order and replayed split points are those of
    batches(). Also: abandoning the iterator early stops the thread, and a failure inside it surfaces in the consumer."""
    rng = np.random.RandomState(3)
    path = str(tmp_path / "p.tdemb")
    n, width = 23, 64
    with td.EmbedShardWriter(path, width=width) as w:
        for i in range(n):
            L = int(rng.randint(2, 40))
            e = torch.from_numpy(rng.standard_normal((L, width)).astype(np.float32)).to(torch.bfloat16)
            w.add(e, list(range(L)), f"t{i}", f"k{i}")
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=1, output_embed_max_split_len=16, output_embed_max_len=64,
              input_embed_max_len=64)
    r = td.EmbedShardReader(path)
    random.seed(11)
    plain = list(r.batches(4, bi, pin_memory=False))
    for depth in (1, 2):
        random.seed(11)
        pre = list(r.batches_prefetched(4, bi, depth=depth, pin_memory=False))
        assert len(pre) == len(plain) == n // 4
        for a, b in zip(plain, pre):
            assert torch.equal(a.flat.view(torch.int16), b.flat.view(torch.int16))
            assert a.lens.tolist() == b.lens.tolist() and a.src_row_start.tolist() == b.src_row_start.tolist() and a.l_max == b.l_max
            assert a.extras["output_token_ids"] == b.extras["output_token_ids"]
    it = r.batches_prefetched(4, bi, depth=2, pin_memory=False)
    next(it)
    it.close()  # consumer walks away: the producer thread must stop instead of blocking on a full queue
    import threading
    import time

    t0 = time.monotonic()
    while any(t.name == "td-shard-prefetch" for t in threading.enumerate()) and time.monotonic() - t0 < 5:
        time.sleep(0.05)
    assert not any(t.name == "td-shard-prefetch" for t in threading.enumerate())
    with pytest.raises(KeyError):  # a broken build_info fails in the thread and is re-raised here
        list(r.batches_prefetched(4, {"random_split_output_embed": 1}, pin_memory=False))
    r.close()


def test_threaded_slab_copy_is_bit_exact(tmp_path):
    """A slab of >= 8 MB is copied by several threads over disjoint row ranges: same bytes as the single-threaded copy, for the
    whole shard and for an unaligned sub-range."""
    rng = np.random.RandomState(5)
    path = str(tmp_path / "big.tdemb")
    width, n = 1024, 40
    with td.EmbedShardWriter(path, width=width) as w:
        for i in range(n):
            L = int(rng.randint(100, 200))
            bits = rng.randint(0, 65536, size=(L, width)).astype(np.uint16)
            w.add(torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16), [i] * L)
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=1 << 30, input_embed_max_len=1 << 30)
    one, many = td.EmbedShardReader(path, copy_threads=1), td.EmbedShardReader(path, copy_threads=5)
    assert one.total_rows * width * 2 >= (8 << 20)
    for lo, hi in ((0, n), (3, n - 2)):
        a, b = one.batch(lo, hi, bi, pin_memory=False), many.batch(lo, hi, bi, pin_memory=False)
        assert torch.equal(a.flat.view(torch.int16), b.flat.view(torch.int16)) and a.lens.tolist() == b.lens.tolist()
    assert many._pool is not None and one._pool is None
    ids = rng.permutation(n).tolist()  # a shuffled batch: one row-range copy per sample, shared by the threads in runs of samples
    for trunc in (False, True):
        a, b = one.batch_indices(ids, dict(bi, output_embed_max_len=150), pin_memory=False, truncate_on_host=trunc), \
            many.batch_indices(ids, dict(bi, output_embed_max_len=150), pin_memory=False, truncate_on_host=trunc)
        assert a.flat.shape[0] * width * 2 >= (8 << 20)
        assert torch.equal(a.flat.view(torch.int16), b.flat.view(torch.int16)) and a.src_row_start.tolist() == b.src_row_start.tolist()
    one.close(), many.close()


def _write_wds_tar(path, samples, with_image=True):
    """A shard as the reference's pre-compute task writes it through wds.ShardWriter (image_text_process_data.py:104-118): per
    sample the members <key>.jpg, <key>.json and one <key>.<flattened name>.pth per tensor, each tensor pickled by torch.save."""
    import io
    import json
    import tarfile

    with tarfile.open(path, "w") as tf:
        for s in samples:
            fields = {}
            if with_image:
                fields["jpg"] = b"\xff\xd8not a real jpeg\xff\xd9"  # never decoded by the converter
            fields["json"] = json.dumps(s["json"]).encode("utf-8")
            for k, v in s.items():
                if k.endswith(".pth"):
                    buf = io.BytesIO()
                    torch.save(v.clone(), buf)
                    fields[k] = buf.getvalue()
            for ext, blob in fields.items():
                ti = tarfile.TarInfo(f"{s['__key__']}.{ext}")
                ti.size = len(blob)
                tf.addfile(ti, io.BytesIO(blob))


@pytest.mark.parametrize("name", ["collater_random_split.npz", "collater_fixed_max.npz"])
def test_converted_webdataset_shards_reproduce_the_reference_collater(tmp_path, name):
    """Reference tar shards -> convert_webdataset_shards -> EmbedShardReader.batch == the reference collater's golden output on
    the same samples, for both embed streams; the samples are spread over two tars and carry the pass-through json fields."""
    g = load_golden(name)
    samples = _golden_samples(g)
    for i, s in enumerate(samples):
        s["json"]["gpt"] = f"gpt {i}"
        s["json"]["revised_generated_text"] = f"revised {i}"
        s["model.norm.input_embed.pth"] = s["model.norm.input_embed.pth"][: max(1, s["model.norm.input_embed.pth"].shape[0] // 2)].clone()
        s["__key__"] = f"train/{i:06d}"
    half = len(samples) // 2
    tars = [str(tmp_path / "00000.tar"), str(tmp_path / "00001.tar")]
    _write_wds_tar(tars[0], samples[:half])
    _write_wds_tar(tars[1], samples[half:], with_image=False)
    decoded = list(td.iter_webdataset_samples(tars))
    assert [d["__key__"] for d in decoded] == [s["__key__"] for s in samples]
    assert all("jpg" not in d and d["json"] == s["json"] for d, s in zip(decoded, samples))
    paths = td.convert_webdataset_shards(tars, str(tmp_path / "flat"))
    assert sorted(paths) == ["input", "output"]
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    r = td.EmbedShardReader(paths["output"])
    random.seed(int(g["seed"]))
    fb = r.batch(0, len(samples), bi, pin_memory=False)
    bits = fb.flat.view(torch.int16).numpy().view(np.uint16)
    packed, cu = pack_ref.pack_from_flat(bits, fb.src_row_start.tolist(), fb.lens.tolist())
    padded, mask = pack_ref.unpack_padded(packed, cu, fb.l_max)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    assert fb.extras["llava_gpts"] == [f"gpt {i}" for i in range(len(samples))]
    assert fb.extras["revised_generated_texts"] == [f"revised {i}" for i in range(len(samples))]
    assert r._meta["keys"] == [s["__key__"] for s in samples]
    r.close()
    ri = td.EmbedShardReader(paths["input"])
    for i, s in enumerate(samples):
        assert torch.equal(ri.embedding(i).view(torch.int16), s["model.norm.input_embed.pth"].view(torch.int16))
    ri.close()
    # the same batch through the worker-side collater on the decoded dicts (the other way into the device pack)
    random.seed(int(g["seed"]))
    fc = td.FlatCollater(dict(bi, use_output_embed=1, use_input_embed=0), pin_memory=False, which="output")(decoded)
    assert fc.lens.tolist() == fb.lens.tolist() and torch.equal(fc.flat.view(torch.int16), fb.flat.view(torch.int16))


def test_converter_errors_and_passthrough_rule(tmp_path):
    e = torch.zeros((3, 8), dtype=torch.bfloat16)
    mk = lambda i, **js: {"__key__": f"k{i}", "json": dict({"generated_text": "t", "output_token_ids": [1, 2, 3]}, **js),  # noqa: E731
                          "model.norm.output_embed.pth": e}
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=8, input_embed_max_len=8)
    # the pass-through keys follow the batch's FIRST sample, as in the reference: absent there -> absent; present there but
    # missing later -> KeyError (the reference indexes json["gpt"] of every sample)
    tar = str(tmp_path / "a.tar")
    _write_wds_tar(tar, [mk(0), mk(1, gpt="g1"), mk(2, gpt="g2")])
    r = td.EmbedShardReader(td.convert_webdataset_shards(tar, str(tmp_path / "a"))["output"])
    assert "llava_gpts" not in r.batch(0, 2, bi, pin_memory=False).extras
    assert r.batch(1, 3, bi, pin_memory=False).extras["llava_gpts"] == ["g1", "g2"]
    r.close()
    _write_wds_tar(tar, [mk(0, gpt="g0"), mk(1)])
    r = td.EmbedShardReader(td.convert_webdataset_shards(tar, str(tmp_path / "b"))["output"])
    with pytest.raises(KeyError):
        r.batch(0, 2, bi, pin_memory=False)
    r.close()
    # a stream that appears or disappears mid-way, a non-bf16 tensor, a tar without embeddings
    s1 = mk(1)
    s1["model.norm.input_embed.pth"] = e
    _write_wds_tar(tar, [mk(0), s1])
    with pytest.raises(ValueError, match="first one with a input_embed"):
        td.convert_webdataset_shards(tar, str(tmp_path / "c"))
    s0 = mk(0)
    s0["model.norm.output_embed.pth"] = e.float()
    _write_wds_tar(tar, [s0])
    with pytest.raises(ValueError, match="bfloat16"):
        td.convert_webdataset_shards(tar, str(tmp_path / "d"))
    _write_wds_tar(tar, [{"__key__": "x", "json": {"generated_text": "", "output_token_ids": []}}])
    with pytest.raises(ValueError, match="no sample"):
        td.convert_webdataset_shards(tar, str(tmp_path / "e"))
    # a failed conversion leaves neither a half-written shard nor the writer's temporary rows file
    import os

    left = sorted(f for f in os.listdir(tmp_path) if f.startswith(("c.", "d.", "e.")) or f.endswith(".rows.tmp"))
    assert left == [], left
    # the writer as a context manager: an exception inside drops the shard, a closed writer refuses more samples
    with pytest.raises(RuntimeError):
        with td.EmbedShardWriter(str(tmp_path / "f.tdemb"), 8) as w:
            w.add(e, [1, 2, 3])
            raise RuntimeError("producer failed")
    assert not os.path.exists(tmp_path / "f.tdemb") and not os.path.exists(str(tmp_path / "f.tdemb") + ".rows.tmp")
    w = td.EmbedShardWriter(str(tmp_path / "g.tdemb"), 8)
    w.add(e, [1, 2, 3])
    w.close()
    assert not os.path.exists(str(tmp_path / "g.tdemb") + ".rows.tmp") and w.close() == str(tmp_path / "g.tdemb")
    with pytest.raises(ValueError, match="closed"):
        w.add(e, [1, 2, 3])


def _small_shard(tmp_path, n=12, width=32, seed=9):
    rng = np.random.RandomState(seed)
    path = str(tmp_path / f"s{seed}.tdemb")
    embeds = []
    with td.EmbedShardWriter(path, width=width) as w:
        for i in range(n):
            L = int(rng.randint(3, 30))
            bits = rng.randint(0, 65536, size=(L, width)).astype(np.uint16)
            e = torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16)
            embeds.append(e)
            w.add(e, list(range(100 * i, 100 * i + L)), f"t{i}", f"k{i}")
    return path, embeds


def test_batch_indices_any_order_and_truncate_on_host(tmp_path):
    path, embeds = _small_shard(tmp_path)
    r = td.EmbedShardReader(path)
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=1, output_embed_max_split_len=10, output_embed_max_len=64,
              input_embed_max_len=64)
    ids = [7, 2, 2, 11, 0]
    random.seed(5)
    fb = r.batch_indices(ids, bi, pin_memory=False)
    random.seed(5)
    want_lens = [random.randint(1, min(embeds[i].shape[0] - 1, 10)) for i in ids]  # the reference's draw, in batch order
    assert fb.lens.tolist() == want_lens and fb.l_max == max(want_lens) and fb.extras["sample_ids"] == ids
    assert fb.extras["generated_texts"] == [f"t{i}" for i in ids]
    assert fb.extras["output_token_ids"] == [list(range(100 * i + n, 100 * i + embeds[i].shape[0])) for i, n in zip(ids, want_lens)]
    for j, i in enumerate(ids):
        s0 = int(fb.src_row_start[j])
        assert torch.equal(fb.flat[s0 : s0 + embeds[i].shape[0]].view(torch.int16), embeds[i].view(torch.int16))
    # truncate_on_host: only the kept rows travel; packing either batch gives the same rows
    random.seed(5)
    ft = r.batch_indices(ids, bi, pin_memory=False, truncate_on_host=True)
    assert ft.lens.tolist() == want_lens and ft.flat.shape[0] == sum(want_lens)
    pa, _ = pack_ref.pack_from_flat(fb.flat.view(torch.int16).numpy(), fb.src_row_start.tolist(), fb.lens.tolist())
    pb, _ = pack_ref.pack_from_flat(ft.flat.view(torch.int16).numpy(), ft.src_row_start.tolist(), ft.lens.tolist())
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(pb, ft.flat.view(torch.int16).numpy())  # already packed
    # a consecutive run equals batch(lo, hi)
    fixed = dict(bi, random_split_output_embed=0, output_embed_max_len=12)
    a, b = r.batch(3, 8, fixed, pin_memory=False), r.batch_indices([3, 4, 5, 6, 7], fixed, pin_memory=False)
    assert torch.equal(a.flat.view(torch.int16), b.flat.view(torch.int16)) and a.lens.tolist() == b.lens.tolist() == [min(embeds[i].shape[0], 12) for i in range(3, 8)]
    with pytest.raises(IndexError):
        r.batch_indices([0, 12], bi, pin_memory=False)
    with pytest.raises(ValueError):
        r.batch_indices([], bi, pin_memory=False)
    r.close()


def test_pinned_ring_recycles_slots_only_after_their_h2d_events(tmp_path, monkeypatch):
    """The pinned staging ring of the GPU path, driven on the CPU with stand-ins for pinning and CUDA events: a slot is refilled
    only after the events its consumer reported have been waited for; a slot whose consumer never reported is dropped, not
    overwritten; a reported slot keeps its buffer."""
    path, embeds = _small_shard(tmp_path, n=16)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    r = td.EmbedShardReader(path)
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=64, input_embed_max_len=64)
    waited = []

    class Ev:
        def __init__(self, tag):
            self.tag = tag

        def synchronize(self):
            waited.append(self.tag)

    seen = []
    for j in range(7):
        fb = r.batch(2 * j, 2 * j + 2, bi)
        want = torch.cat([embeds[2 * j], embeds[2 * j + 1]])
        assert torch.equal(fb.flat.view(torch.int16), want.view(torch.int16))
        seen.append((fb.flat.data_ptr(), fb))  # keep every batch alive: a dropped buffer must not be recycled by the allocator
        if j != 1:
            fb.extras["_h2d_enqueued"]([Ev(j)])  # batch 1's consumer never reports
        # slot of batch j is reused by batch j + 3: the events of batch j are waited for exactly then
        assert waited == [k for k in range(j - 2) if k != 1], (j, waited)
    ptr = [p for p, _ in seen]
    assert ptr[3] == ptr[0] and ptr[6] == ptr[0]  # reported slot: same pinned buffer again
    assert ptr[4] != ptr[1]                        # unreported slot: a fresh buffer, the old one left to its tensor
    assert torch.equal(seen[1][1].flat.view(torch.int16), torch.cat([embeds[2], embeds[3]]).view(torch.int16))  # ... and intact
    r.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_set_epoch_plan_is_a_rank_partition_with_equal_batch_counts(tmp_path, world):
    paths = [_small_shard(tmp_path, n=n, seed=s)[0] for n, s in ((37, 1), (16, 2), (64, 3))]
    ss = td.EmbedShardSet(paths)
    assert len(ss) == 117
    bs = 2
    plans = [ss.plan(bs, seed=7, epoch=3, rank=r, world=world) for r in range(world)]
    assert len({len(p) for p in plans}) == 1                       # no rank runs dry before the others
    assert [[si for si, _ in p] for p in plans] == [[si for si, _ in plans[0]]] * world  # same shard at the same step on all ranks
    used = [(si, i) for p in plans for si, ids in p for i in ids]
    assert len(used) == len(set(used))                             # a sample is used at most once per epoch, by one rank
    assert all(len(ids) == bs for p in plans for _, ids in p)
    per_shard = {si: ss.counts[si] // world // bs * bs * world for si in range(3)}
    assert len(used) == sum(per_shard.values())                    # only the remainders (< world * bs per shard) are dropped
    assert ss.plan(bs, 7, 3, 0, world) == plans[0] and ss.plan(bs, 7, 4, 0, world) != plans[0]  # f(seed, epoch)
    seq = [ss.plan(bs, rank=r, world=world, shuffle=False) for r in range(world)]
    assert all(ids == list(range(ids[0], ids[0] + bs)) for p in seq for _, ids in p)  # unshuffled: consecutive ids (slab copies)
    with pytest.raises(ValueError):
        ss.plan(bs, rank=world, world=world)
    ss.close()


def test_shard_set_batches_are_bit_exact(tmp_path):
    shards = [_small_shard(tmp_path, n=n, seed=s) for n, s in ((10, 4), (9, 5))]
    ss = td.EmbedShardSet([p for p, _ in shards])
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=20, input_embed_max_len=20)
    n = 0
    for fb in ss.batches(3, bi, seed=1, epoch=0, rank=1, world=2, pin_memory=False, truncate_on_host=True):
        embeds = shards[[p for p, _ in shards].index(fb.extras["shard"])][1]
        for j, i in enumerate(fb.extras["sample_ids"]):
            s0, k = int(fb.src_row_start[j]), int(fb.lens[j])
            assert k == min(embeds[i].shape[0], fb.l_max) and torch.equal(fb.flat[s0 : s0 + k].view(torch.int16), embeds[i][:k].view(torch.int16))
        n += 1
    assert n == len(ss.plan(3, 1, 0, 1, 2)) == 2
    ss.close()
    # shards are opened on demand and at most max_open stay mapped; all of them stage through one pinned ring
    lazy = td.EmbedShardSet([p for p, _ in shards], max_open=1)
    assert lazy._open == {} and len(lazy) == 19
    r0 = lazy.reader(0)
    assert list(lazy._open) == [0] and lazy.reader(0) is r0
    r1 = lazy.reader(1)
    assert list(lazy._open) == [1] and r0.rows is None and r1._ring is lazy._ring  # shard 0 closed, same ring handed on
    lazy.close()
    assert lazy._open == {}
    w = td.EmbedShardWriter(str(tmp_path / "w8.tdemb"), width=8)
    w.add(torch.zeros((2, 8), dtype=torch.bfloat16), [1, 2])
    w.close()
    with pytest.raises(ValueError, match="widths"):
        td.EmbedShardSet([shards[0][0], str(tmp_path / "w8.tdemb")])


def test_prefetch_ring_has_a_slot_for_every_live_batch(tmp_path, monkeypatch):
    """batches_prefetched(depth=2) on the (mocked) pinned path: the consumer's batch, two queued ones and the one being filled
    are four different ring slots -- the batch in the consumer's hands is never overwritten, and with every consumer reporting
    its H2D events no slot is ever dropped (exactly depth + 2 buffers over the whole epoch)."""
    import time

    width, L, n = 16, 5, 40
    path = str(tmp_path / "eq.tdemb")
    rng = np.random.RandomState(0)
    embeds = []
    with td.EmbedShardWriter(path, width=width) as w:
        for i in range(n):
            e = torch.from_numpy(rng.randint(0, 65536, size=(L, width)).astype(np.uint16).view(np.int16)).view(torch.bfloat16)
            embeds.append(e)
            w.add(e, [i] * L)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)

    class Ev:
        def synchronize(self):
            pass

    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=64, input_embed_max_len=64)
    ss = td.EmbedShardSet([path])
    ptrs = set()
    for j, fb in enumerate(ss.batches_prefetched(2, bi, depth=2, shuffle=False)):
        fb.extras["_h2d_enqueued"]([Ev()])
        ptrs.add(fb.flat.data_ptr())
        time.sleep(0.002)  # let the producer run as far ahead as it is allowed to
        assert torch.equal(fb.flat.view(torch.int16), torch.cat(embeds[2 * j : 2 * j + 2]).view(torch.int16)), j
    assert j == n // 2 - 1 and len(ptrs) == 4
    ss.close()


def test_pinned_ring_stops_repinning_on_ragged_batches(tmp_path, monkeypatch):
    """Pinned buffers are sized for the largest batch seen so far plus 1/8: once every slot has been refilled after the largest
    batch showed up, nothing is re-pinned any more (a re-pin is a cudaHostAlloc of the whole slab)."""
    path, embeds = _small_shard(tmp_path, n=24, seed=21)
    pins = []
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: (pins.append(self.shape[0]), self)[1])
    r = td.EmbedShardReader(path)
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=64, input_embed_max_len=64)

    class Ev:
        def synchronize(self):
            pass

    for epoch in range(6):
        if epoch == 4:
            early = len(pins)
        for j in range(8):
            fb = r.batch(3 * j, 3 * j + 3, bi)
            assert torch.equal(fb.flat.view(torch.int16), torch.cat(embeds[3 * j : 3 * j + 3]).view(torch.int16))
            fb.extras["_h2d_enqueued"]([Ev()])
    biggest = max(sum(e.shape[0] for e in embeds[3 * j : 3 * j + 3]) for j in range(8))
    assert len(pins) == early <= 6                # a slot is re-pinned only when it meets a batch larger than its buffer: at most twice here
    assert pins == sorted(pins)
    assert max(pins) <= biggest + (biggest >> 3)  # and never more than 1/8 above the largest batch
    r.close()


def test_shard_set_opens_the_next_shard_ahead_and_keeps_two_open(tmp_path):
    shards = [_small_shard(tmp_path, n=8, seed=30 + k) for k in range(4)]
    ss = td.EmbedShardSet([p for p, _ in shards])
    bi = dict(use_output_embed=1, use_input_embed=0, random_split_output_embed=0, output_embed_max_len=64, input_embed_max_len=64)
    plan = ss.plan(4, seed=2, epoch=0)
    order = [si for k, (si, _) in enumerate(plan) if k == 0 or plan[k - 1][0] != si]
    assert sorted(order) == [0, 1, 2, 3]
    seen = []
    for k, fb in enumerate(ss.batches(4, bi, seed=2, epoch=0, pin_memory=False)):
        si = plan[k][0]
        embeds = shards[si][1]
        for j, i in enumerate(fb.extras["sample_ids"]):
            s0 = int(fb.src_row_start[j])
            assert torch.equal(fb.flat[s0 : s0 + embeds[i].shape[0]].view(torch.int16), embeds[i].view(torch.int16))
        ss._lookahead and ss._lookahead.join()
        pos = order.index(si)
        # this shard and the next one are open (after the last shard: its predecessor is simply not evicted), never more than two
        assert set(order[pos : pos + 2]) <= set(ss._open) <= set(order[max(pos - 1, 0) : pos + 2]) and len(ss._open) <= 2, (k, list(ss._open))
        seen.append(si)
    assert len(seen) == len(plan) == 8
    ss.close()
