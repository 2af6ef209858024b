"""GPU parity: ragged pack / pad / mask kernels through the C ABI vs the numpy oracle (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _u16(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


def _ragged(B, max_len, C, seed, extra=True):
    rng = np.random.RandomState(seed)
    lens = rng.randint(1, max_len + 1, size=B)
    full = lens + (1 + np.arange(B) % 32 if extra else 0)
    start = np.concatenate([[0], np.cumsum(full)[:-1]]).astype(np.int64)
    bits = rng.randint(0, 65536, size=(int(full.sum()), C)).astype(np.uint16)  # arbitrary bit patterns, NaNs included
    return lens.astype(np.int32), start, bits


@pytest.mark.parametrize("B,max_len,C", [(64, 256, 3584), (7, 33, 4096), (5, 9, 8), (1, 1, 64), (128, 64, 768)])
def test_pack_and_padded_bit_exact(B, max_len, C):
    import thinkdiff_mlre_b200 as td
    from oracle import pack_ref

    lens, start, bits = _ragged(B, max_len, C, seed=B * 7 + C)
    dev = torch.device("cuda")
    flat = torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16).to(dev)
    pb = td.pack_device(flat, torch.from_numpy(start).to(dev), torch.from_numpy(lens).to(dev), int(lens.sum()), int(lens.max()))
    ref_packed, ref_cu = pack_ref.pack_from_flat(bits, start.tolist(), lens.tolist())
    assert pb.cu_seqlens.dtype == torch.int32
    np.testing.assert_array_equal(pb.cu_seqlens.cpu().numpy(), ref_cu)
    np.testing.assert_array_equal(_u16(pb.x), ref_packed)
    # reference layout straight from the flat source, and from the packed rows (inverse of pack)
    ref_padded, ref_mask = pack_ref.unpack_padded(ref_packed, ref_cu)
    padded, mask = td.ops.pack_padded(flat, torch.from_numpy(start).to(dev), pb.cu_seqlens, int(lens.max()))
    assert mask.dtype == torch.int64
    np.testing.assert_array_equal(_u16(padded), ref_padded)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)
    padded2, mask2 = pb.to_padded()
    assert torch.equal(padded2.view(torch.int16), padded.view(torch.int16)) and torch.equal(mask2, mask)


def test_pack_matches_reference_collater_golden():
    """End to end against the reference collater's own output (golden): FlatCollater -> H2D -> device pack -> padded."""
    import random

    import thinkdiff_mlre_b200 as td
    from oracle.golden import load_golden

    g = load_golden("collater_random_split.npz")
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    full = [int(v) for v in g["full_lens"]]
    off = np.concatenate([[0], np.cumsum(full)])
    samples = []
    for i in range(len(full)):
        e = torch.from_numpy(g["src_bits"][off[i] : off[i + 1]].view(np.int16).copy()).view(torch.bfloat16)
        ids = [int(v) for v in g["src_ids_flat"][off[i] : off[i + 1]]]
        samples.append({"json": {"generated_text": "", "output_token_ids": ids}, "model.norm.input_embed.pth": e,
                        "model.norm.output_embed.pth": e})
    random.seed(int(g["seed"]))
    fb = td.FlatCollater(bi)(samples)
    pb = td.pack_batch(fb, "cuda")
    padded, mask = pb.to_padded()
    np.testing.assert_array_equal(_u16(padded), g["out_embed_bits"])
    np.testing.assert_array_equal(mask.cpu().numpy(), g["out_mask"])


def test_reference_dict_reproduces_the_reference_collater_output_for_both_streams():
    """FlatCollater.reference_dict == the dict the reference collater returns (golden), both embed streams of one call padded on
    the device, int64 masks, pass-through lists."""
    import random

    import thinkdiff_mlre_b200 as td
    from oracle.golden import load_golden

    g = load_golden("collater_input_embed.npz")
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    full = [int(v) for v in g["full_lens"]]
    off = np.concatenate([[0], np.cumsum(full)])
    samples = []
    for i in range(len(full)):
        e = torch.from_numpy(g["src_bits"][off[i] : off[i + 1]].view(np.int16).copy()).view(torch.bfloat16)
        ids = [int(v) for v in g["src_ids_flat"][off[i] : off[i + 1]]]
        samples.append({"json": {"generated_text": f"t{i}", "output_token_ids": ids, "gpt": f"g{i}"},
                        "model.norm.input_embed.pth": e, "model.norm.output_embed.pth": e})
    random.seed(int(g["seed"]))
    out = td.FlatCollater.reference_dict(td.FlatCollater(bi)(samples), "cuda")
    assert set(out) == {"generated_texts", "output_token_ids", "llava_gpts", "model.norm.input_embed", "input_embed_mask",
                        "model.norm.output_embed", "output_embed_mask"}
    np.testing.assert_array_equal(_u16(out["model.norm.output_embed"]), g["out_embed_bits"])
    np.testing.assert_array_equal(out["output_embed_mask"].cpu().numpy(), g["out_mask"])
    np.testing.assert_array_equal(_u16(out["model.norm.input_embed"]), g["in_embed_bits"])
    np.testing.assert_array_equal(out["input_embed_mask"].cpu().numpy(), g["in_mask"])
    assert out["output_embed_mask"].dtype == torch.int64 and out["llava_gpts"] == [f"g{i}" for i in range(len(full))]
    ids = [list(map(int, g["ids_flat"][g["ids_off"][i] : g["ids_off"][i + 1]])) for i in range(len(full))]
    assert [list(map(int, t)) for t in out["output_token_ids"]] == ids


def test_pack_fp32_rows_and_int_payloads():
    import thinkdiff_mlre_b200 as td

    dev = torch.device("cuda")
    flat = torch.arange(40 * 12, dtype=torch.float32, device=dev).reshape(40, 12)  # 48-byte rows
    lens = torch.tensor([3, 1, 10], dtype=torch.int32, device=dev)
    start = torch.tensor([0, 8, 20], dtype=torch.int64, device=dev)
    pb = td.pack_device(flat, start, lens, 14, 10)
    want = torch.cat([flat[0:3], flat[8:9], flat[20:30]])
    assert torch.equal(pb.x, want)


def test_pack_edge_cases_and_errors():
    import thinkdiff_mlre_b200 as td

    dev = torch.device("cuda")
    empty = td.pack_device(torch.empty((0, 64), dtype=torch.bfloat16, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                           torch.empty(0, dtype=torch.int32, device=dev), 0, 0)
    assert empty.x.shape == (0, 64) and empty.cu_seqlens.tolist() == [0]
    flat = torch.zeros((4, 4), dtype=torch.bfloat16, device=dev)  # 8-byte rows: not a multiple of 16
    with pytest.raises(RuntimeError, match="multiple of 16"):
        td.pack_device(flat, torch.zeros(1, dtype=torch.int64, device=dev), torch.ones(1, dtype=torch.int32, device=dev), 1, 1)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        td.ops.cu_seqlens(torch.ones(3, dtype=torch.int32))
    big = torch.randint(1, 5, (5000,), dtype=torch.int32, device=dev)  # multi-chunk scan
    cu = td.ops.cu_seqlens(big)
    assert torch.equal(cu[1:].long(), torch.cumsum(big.long(), 0)) and int(cu[0]) == 0


def test_pack_full_size_long_context_roundtrip():
    """BASELINE config 5 per-GPU shard (128 sequences, len <= 1024, d = 3584): size-independent properties."""
    import thinkdiff_mlre_b200 as td

    dev = torch.device("cuda")
    b = td.synthetic_lvlm_batch(128, 1024, 3584, 4096, seed=1234, pin=False, with_target=False)
    flat, start, lens = b.flat.to(dev), b.src_row_start.to(dev), b.lens.to(dev)
    pb = td.pack_device(flat, start, lens, b.total_rows, b.l_max)
    assert pb.x.shape == (b.total_rows, 3584) and int(pb.cu_seqlens[-1]) == b.total_rows
    # every kept row equals its source row (gather by index computed with torch), and padding is exact zeros
    seq = torch.repeat_interleave(torch.arange(128, device=dev), lens.long())
    within = torch.arange(b.total_rows, device=dev) - pb.cu_seqlens[:-1].long()[seq]
    assert torch.equal(pb.x.view(torch.int16), flat[start[seq] + within].view(torch.int16))
    padded, mask = pb.to_padded()
    assert torch.equal(mask.sum(1).int(), lens) and torch.equal(mask, (torch.arange(b.l_max, device=dev)[None] < lens[:, None]).long())
    assert not padded[mask == 0].view(torch.int16).any()
    # packing the padded batch again (source = padded rows) is idempotent
    again = td.pack_device(padded.reshape(-1, 3584), torch.arange(128, device=dev) * b.l_max, lens, b.total_rows, b.l_max)
    assert torch.equal(again.x.view(torch.int16), pb.x.view(torch.int16))


def test_clip_two_image_plus_text_composition_config4():
    """BASELINE config 4 (per-GPU shard: 32 samples): 2 x 32 CLIP tokens [.., 768] -> aligner (pure-bf16 inference
    regime) -> [64, 4096] per sample, composed with ragged T5 text embeddings [T_i, 4096], T_i ~ U{1..128}. The packed
    rows and the padded prompt_embeds must equal the reference's per-sample torch.cat (bit-exact: byte moves)."""
    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    dev = torch.device("cuda")
    B, n_img, din, d = 32, 64, 768, 4096
    g = torch.Generator().manual_seed(4)
    m = td.ThinkDiffAligner(din, d).to(dev)
    m.load_state_dict(aligner_ref.init_params_numpy(din, d, seed=4))
    m = m.to(torch.bfloat16).eval()
    img = torch.randn((B, n_img, din), generator=g).to(torch.bfloat16).to(dev)
    with torch.no_grad():
        y = m(img)
    assert y.shape == (B, n_img, d) and y.dtype == torch.bfloat16
    t_lens = torch.randint(1, 129, (B,), generator=g).tolist()
    text = torch.randn((sum(t_lens), d), generator=g).to(torch.bfloat16).to(dev)
    pb = td.compose_image_text(y, text, t_lens)
    off, want = 0, []
    for i in range(B):
        want.append(torch.cat([y[i], text[off : off + t_lens[i]]]))
        off += t_lens[i]
    assert torch.equal(pb.x.view(torch.int16), torch.cat(want).view(torch.int16))
    assert pb.cu_seqlens.tolist() == [0] + torch.cumsum(torch.tensor([n_img + t for t in t_lens]), 0).tolist()
    padded, mask = pb.to_padded()
    assert padded.shape == (B, n_img + max(t_lens), d)
    for i in range(B):
        n = n_img + t_lens[i]
        assert torch.equal(padded[i, :n].view(torch.int16), want[i].view(torch.int16))
        assert not padded[i, n:].view(torch.int16).any() and mask[i].sum().item() == n
    # the aligner rows themselves: pure-bf16 regime vs the oracle on a few samples
    p16 = {k: v.to(torch.bfloat16).float() for k, v in aligner_ref.init_params_numpy(din, d, seed=4).items()}
    ref = aligner_ref.aligner_fwd_bwd_manual(img[:2].reshape(-1, din).float().cpu(), p16, regime="bf16", out_bf16=True, accum_dtype=torch.float32)
    err = (y[:2].reshape(-1, d).float().cpu() - ref["y"]).norm() / ref["y"].norm()
    assert err < 2e-2


@pytest.mark.parametrize("k", [1, 2, 4])
def test_prefetch_multi_stream_h2d_is_exact(k):
    """AlignerTrainStep.prefetch cuts the flat features / targets into row chunks over k copy streams: the device tensors
    must equal the pinned host tensors bit for bit once the returned events have been waited on."""
    import thinkdiff_mlre_b200 as td

    m = td.ThinkDiffAligner(192, 512).cuda()
    step = td.AlignerTrainStep(m, None)
    step.copy_streams = k
    b = td.synthetic_lvlm_batch(9, 70, 192, 512, seed=77)
    (flat, start, lens, total_rows, l_max, tgt), events = step.prefetch(b, "cuda")
    for ev in events:
        torch.cuda.current_stream().wait_event(ev)
    torch.cuda.synchronize()
    assert torch.equal(flat.cpu().view(torch.int16), b.flat.view(torch.int16))
    assert torch.equal(tgt.cpu().view(torch.int16), b.extras["flat_target"].view(torch.int16))
    assert torch.equal(start.cpu(), b.src_row_start) and torch.equal(lens.cpu(), b.lens)
    assert total_rows == b.total_rows and l_max == b.l_max
