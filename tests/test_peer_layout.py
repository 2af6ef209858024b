"""Host logic of the peer-memory data-parallel mode (thinkdiff_mlre_b200/peer.py): the exchange-buffer layout is pure
arithmetic shared by all ranks -- regions must not overlap, must be aligned for float4 / TMA-free vector access, and the
per-rank destination addresses must tile each region exactly. No GPU needed."""
import pytest

from thinkdiff_mlre_b200.peer import FLAG_BYTES, FLAG_ROW_STRIDE, ROW_GRAD1, ROW_W2, ExchangeLayout


@pytest.mark.parametrize("world,din,d", [(1, 192, 512), (2, 192, 512), (4, 4096, 4096), (8, 4096, 4096), (8, 1024, 4096)])
def test_regions_are_disjoint_aligned_and_tiled(world, din, d):
    lay = ExchangeLayout(world, din, d)
    regions = [
        ("flags", 0, FLAG_BYTES),
        ("g1", lay.off_g1, 4 * world * lay.slot1_numel),
        ("g2", lay.off_g2, 4 * world * lay.slot2_numel),
        ("small", lay.off_small, 4 * world * lay.small_numel),
        ("w1", lay.off_w1, 2 * d * din),
        ("w2", lay.off_w2, 2 * d * d),
    ]
    end = 0
    for name, off, size in regions:
        assert off >= end, name
        assert off % 256 == 0, name
        end = off + size
    assert lay.total_bytes >= end
    # the N slots of one weight tile its region; the N owners' row blocks tile the bf16 weight
    for which, numel, cols, off_g, off_w in ((1, lay.slot1_numel, din, lay.off_g1, lay.off_w1), (2, lay.slot2_numel, d, lay.off_g2, lay.off_w2)):
        assert numel == (d // world) * cols and numel % 4 == 0
        assert [lay.grad_slot_offset(which, s) for s in range(world)] == [off_g + 4 * s * numel for s in range(world)]
        assert [lay.weight_rows_offset(which, o) for o in range(world)] == [off_w + 2 * o * numel for o in range(world)]
        assert lay.weight_rows_offset(which, world - 1) + 2 * numel == off_w + 2 * d * cols
    assert lay.small_numel == 3 * d and lay.small_slot_offset(world - 1) + 4 * lay.small_numel <= lay.off_w1
    # flags: one monotone int32 counter per row, one 128-byte line per row
    assert lay.flag_offset(ROW_GRAD1) == 0 and lay.flag_offset(ROW_W2) == 4 * ROW_W2 * FLAG_ROW_STRIDE
    assert 4 * FLAG_ROW_STRIDE >= 128 and lay.flag_offset(ROW_W2) + 4 <= FLAG_BYTES


def test_rejects_unsupported_worlds():
    with pytest.raises(ValueError):
        ExchangeLayout(9, 512, 512)
    with pytest.raises(ValueError):
        ExchangeLayout(3, 512, 512)  # 512 rows do not divide by 3


import ctypes as C

import pytest


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("N", [3584, 4096])
def test_scattered_gemm_tile_order_is_a_permutation_across_ranks(world, N):
    """Traffic shaping of the data-parallel weight-gradient GEMMs (host arithmetic of the library, no GPU): every rank walks every
    output tile exactly once; at every position of that walk the `world` ranks store to `world` DIFFERENT owners (NVLink sees a
    permutation, never N senders into one port); each rank starts with the rows of owner rank + 1 and ends with its own rows."""
    from thinkdiff_mlre_b200 import _lib as L

    M = 4096
    f = L.lib().td_scatter_tile_owner
    tiles = (M // 256) * ((N + 255) // 256)
    per_rank = []
    for rank in range(world):
        owners, coords = [], set()
        for t in range(tiles):
            m, n = C.c_int32(-1), C.c_int32(-1)
            o = f(t, M, N, world, rank, C.byref(m), C.byref(n))
            assert 0 <= o < world and o == (m.value * 256) // (M // world)
            owners.append(o)
            coords.add((m.value, n.value))
        assert len(coords) == tiles  # a bijection onto the tile grid
        assert owners[0] == (rank + 1) % world and owners[-1] == rank
        assert all(owners[i] == owners[i - 1] or owners[i] == (owners[i - 1] + 1) % world for i in range(1, tiles))
        per_rank.append(owners)
    for t in range(tiles):
        assert len({per_rank[r][t] for r in range(world)}) == world
    # without a rank (rank < 0: the local-output tile order) all callers agree, and bad arguments are rejected
    assert f(0, M, N, world, -1, None, None) == 0
    assert f(tiles, M, N, world, 0, None, None) == -1 and f(0, M, N, 3, 0, None, None) == -1


def test_scattered_gemm_tile_order_small_owner_blocks_keep_the_common_order():
    """Owner blocks smaller than a 256-row tile (the dims of bench.py's dp_parity check) cannot be rotated tile-wise: every rank
    keeps the common order, which is also what keeps that check bit-identical to the local-output GEMM."""
    from thinkdiff_mlre_b200 import _lib as L

    f = L.lib().td_scatter_tile_owner
    for rank in range(8):
        m = C.c_int32(-1)
        assert f(0, 512, 192, 8, rank, C.byref(m), None) == 0 and m.value == 0
