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
