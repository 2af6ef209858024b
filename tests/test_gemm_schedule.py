"""The GEMM's work schedule (whole-tile waves + stream-K tail, thinkdiff_mlre_b200/csrc/gemm_sm100.cuh / gemm_host.cuh), checked
on the CPU through `td_gemm_schedule`: the library runs the same `plan_schedule()` / `for_each_segment()` on the host that the launch
and the kernel's workers run, so these are properties of the shipped scheduler, not of a model of it. The GPU tests check numbers;
this checks that every k-block of every tile is computed exactly once, by whom, and in an order that cannot deadlock."""
import ctypes as C

import pytest

from thinkdiff_mlre_b200 import _lib as L

KMIN = 8  # kMinTailKBlocks
SHAPES = [
    (8460, 4096, 3584), (8460, 4096, 4096), (4096, 4096, 8460), (4096, 3584, 8460),  # the five GEMMs of a config-2 step
    (65600, 4096, 3584), (4096, 4096, 65600),                                       # config 5
    (8192, 32128, 4096), (8192, 4096, 32128),                                       # frozen T5 head and its input gradient
    (1, 4096, 768), (300, 512, 192), (2048, 4096, 768), (512, 192, 300), (257, 288, 64),
]


def schedule(M, N, K, workers, stream_k=1):
    f = L.lib().td_gemm_schedule
    n = f(M, N, K, workers, stream_k, None, 0)
    assert n > 0
    buf = (C.c_int32 * (8 * n))()
    assert f(M, N, K, workers, stream_k, buf, n) == n
    keys = ("unit", "tile", "kb0", "kb1", "kind", "contrib0", "contrib_n", "range")
    return [dict(zip(keys, buf[8 * i : 8 * i + 8])) for i in range(n)]


@pytest.mark.parametrize("workers", [74, 37, 3, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_every_k_block_of_every_tile_is_computed_exactly_once(shape, workers):
    M, N, K = shape
    tiles = -(-M // 256) * -(-N // 256)
    KB = -(-K // 64)
    segs = schedule(M, N, K, workers)
    seen = {}
    for s in segs:
        assert 0 <= s["tile"] < tiles and 0 <= s["kb0"] < s["kb1"] <= KB
        for kb in range(s["kb0"], s["kb1"]):
            assert (s["tile"], kb) not in seen
            seen[(s["tile"], kb)] = s["unit"]
    assert len(seen) == tiles * KB
    units = sorted({s["unit"] for s in segs})
    assert units == list(range(len(units)))  # unit numbers are dense: the claim counter hands out 0, 1, 2, ...
    whole = [s for s in segs if s["kind"] == 0 and s["kb0"] == 0 and s["kb1"] == KB]
    tail = [s for s in segs if s not in whole]
    # the whole-tile part is full waves of `w` workers (or everything, when there are fewer tiles than workers)
    assert len(whole) % min(workers, max(len(whole), 1)) == 0 or not tail
    assert len({s["range"] for s in tail}) <= workers


@pytest.mark.parametrize("workers", [74, 37, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_shared_tiles_have_one_owner_whose_contributors_are_claimed_first(shape, workers):
    M, N, K = shape
    KB = -(-K // 64)
    segs = schedule(M, N, K, workers)
    by_tile = {}
    for s in segs:
        by_tile.setdefault(s["tile"], []).append(s)
    unit_of_range = {}
    for s in segs:
        if s["kind"] in (1, 2) or (s["kind"] == 0 and not (s["kb0"] == 0 and s["kb1"] == KB)):
            unit_of_range.setdefault(s["range"], s["unit"])
    for tile, parts in by_tile.items():
        if len(parts) == 1:
            p = parts[0]
            assert p["kind"] == 0 and p["kb0"] == 0 and p["kb1"] == KB  # computed by one worker, plain epilogue
            continue
        owners = [p for p in parts if p["kind"] == 1]
        contribs = [p for p in parts if p["kind"] == 2]
        assert len(owners) == 1 and len(contribs) == len(parts) - 1
        o = owners[0]
        assert o["kb0"] == 0 and o["kb1"] < KB  # the owner holds the tile's first k-block
        # its contributors are exactly the tail ranges that hold the rest of the tile, in worker order
        assert sorted(c["range"] for c in contribs) == list(range(o["contrib0"], o["contrib0"] + o["contrib_n"]))
        assert all(c["kb0"] > 0 for c in contribs)
        # tail ranges are handed out in DESCENDING order: every contributor is claimed -- by a worker that is running -- before
        # the owner that will wait for it; a late-resident worker can therefore never be waited on for work nobody started
        assert all(c["unit"] < o["unit"] for c in contribs)
    # a range spans at most two tiles, and (when the tail is long enough to choose) no range is shorter than KMIN k-blocks
    per_unit = {}
    for s in segs:
        per_unit.setdefault(s["unit"], []).append(s)
    assert all(len(v) <= 2 for v in per_unit.values())
    tail_units = [v for v in per_unit.values() if not (len(v) == 1 and v[0]["kind"] == 0 and v[0]["kb1"] - v[0]["kb0"] == KB)]
    if tail_units:
        lens = [sum(s["kb1"] - s["kb0"] for s in v) for v in tail_units]
        assert max(lens) - min(lens) <= 1  # equal ranges
        r_tiles = len({s["tile"] for v in tail_units for s in v})
        if r_tiles * KB >= KMIN * r_tiles and len(tail_units) > r_tiles:
            assert min(lens) >= KMIN


@pytest.mark.parametrize("shape", SHAPES[:6])
def test_without_a_workspace_every_tile_is_a_whole_tile(shape):
    M, N, K = shape
    segs = schedule(M, N, K, 74, stream_k=0)
    KB = -(-K // 64)
    assert len(segs) == -(-M // 256) * -(-N // 256)
    assert all(s["kind"] == 0 and s["kb0"] == 0 and s["kb1"] == KB for s in segs)
    assert [s["tile"] for s in segs] == list(range(len(segs)))


def test_schedule_rejects_bad_arguments():
    f = L.lib().td_gemm_schedule
    assert f(0, 4096, 64, 74, 1, None, 0) == -1 and f(16, 0, 64, 74, 1, None, 0) == -1 and f(16, 32, 64, 0, 1, None, 0) == -1
