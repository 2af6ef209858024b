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
    one.close(), many.close()
