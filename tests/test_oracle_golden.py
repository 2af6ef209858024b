"""The oracle (oracle/) against the committed golden vectors, which were produced by the REFERENCE'S OWN code
(oracle/make_golden.py exec's build_vision_projector / the collater from /root/reference). CPU only."""
import numpy as np
import pytest
import torch

from oracle import aligner_ref, loss_ref, pack_ref
from oracle.golden import load_golden


def _params(g):
    return {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p_")}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)


@pytest.mark.parametrize("name,autocast", [("aligner_small_fp32.npz", False), ("aligner_small_bf16.npz", True)])
def test_module_restatement_matches_reference_outputs(name, autocast):
    g = load_golden(name)
    m = aligner_ref.RefAligner(int(g["din"]), int(g["d"]))
    m.load_state_dict(_params(g))
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    y, loss, grads = aligner_ref.module_fwd_bwd(m, x, lambda y: torch.nn.functional.mse_loss(y, t), autocast_bf16=autocast)
    assert str(y.dtype) == str(g["y_dtype"]) == "torch.float32"  # training regime returns fp32 (fp32 norm weight)
    # same torch build, same ops: the restatement must reproduce the reference module bit for bit
    np.testing.assert_array_equal(y.numpy(), g["y"])
    np.testing.assert_array_equal(loss.numpy(), g["loss"])
    for k, v in grads.items():
        np.testing.assert_array_equal(v.numpy(), g["g_" + k])
        assert v.dtype == torch.float32


@pytest.mark.parametrize("name,regime,tol", [("aligner_small_fp32.npz", "fp32", 1e-5), ("aligner_small_bf16.npz", "bf16", 2e-2)])
def test_closed_form_matches_reference_outputs(name, regime, tol):
    """The closed-form forward/backward (what the CUDA kernels implement) against the reference's autograd."""
    g = load_golden(name)
    p = _params(g)
    x, t = torch.from_numpy(g["x"]).reshape(-1, int(g["din"])), torch.from_numpy(g["t"]).reshape(-1, int(g["d"]))
    fwd = aligner_ref.aligner_fwd_bwd_manual(x, p, regime=regime)
    dy = 2.0 * (fwd["y"] - t) / t.numel()
    out = aligner_ref.aligner_fwd_bwd_manual(x, p, dy=dy, regime=regime)
    assert _rel(out["y"].numpy(), g["y"].reshape(-1, int(g["d"]))) < tol
    for key, name_ in (("0.weight", "dW1"), ("0.bias", "db1"), ("2.weight", "dW2"), ("2.bias", "db2"), ("3.weight", "dg")):
        assert _rel(out[name_].numpy(), g["g_" + key]) < tol, name_
    if regime == "fp32":
        np.testing.assert_allclose(out["y"].numpy(), g["y"].reshape(-1, int(g["d"])), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["aligner_mid_fp32.npz", "aligner_mid_bf16.npz", "aligner_mid_bf16_heavy.npz"])
def test_mid_fixture_rows_and_sampled_grads(name):
    g = load_golden(name)
    din, d, seed = int(g["din"]), int(g["d"]), int(g["seed"])
    autocast = bool(g["autocast_bf16"])
    p = aligner_ref.init_params_numpy(din, d, seed)
    rng = np.random.RandomState(seed + 1)
    x = rng.standard_normal(tuple(g["x_shape"])).astype(np.float32)
    if int(g["heavy_tail"]):
        ch = rng.choice(din, size=8, replace=False)
        x[..., ch] *= 50.0
    t = rng.standard_normal(tuple(g["x_shape"][:-1]) + (d,)).astype(np.float32)
    m = aligner_ref.RefAligner(din, d)
    m.load_state_dict(p)
    y, loss, grads = aligner_ref.module_fwd_bwd(m, torch.from_numpy(x), lambda y: torch.nn.functional.mse_loss(y, torch.from_numpy(t)), autocast)
    np.testing.assert_array_equal(y.numpy().reshape(-1, d)[g["y_rows"]], g["y_sel"])
    np.testing.assert_array_equal(loss.numpy(), g["loss"])
    for k, v in grads.items():
        if v.ndim == 1:
            np.testing.assert_array_equal(v.numpy(), g["g_" + k])
        else:
            np.testing.assert_array_equal(v.numpy().reshape(-1)[g["gi_" + k]], g["gs_" + k])


def test_t5_rmsnorm_restatement_equals_transformers():
    T5LayerNorm = pytest.importorskip("transformers.models.t5.modeling_t5").T5LayerNorm
    torch.manual_seed(0)
    ref, mine = T5LayerNorm(96), aligner_ref.T5RMSNorm(96)
    w = torch.randn(96)
    ref.weight.data.copy_(w), mine.weight.data.copy_(w)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(5, 7, 96).to(dt)
        assert torch.equal(ref(x), mine(x))
        assert torch.equal(ref.to(torch.bfloat16)(x), mine.to(torch.bfloat16)(x))
        ref.float(), mine.float()


# ------------------------------------------------------------------------------------------------ collater
def _samples(g):
    full = [int(v) for v in g["full_lens"]]
    off = np.concatenate([[0], np.cumsum(full)])
    embeds = [g["src_bits"][off[i] : off[i + 1]] for i in range(len(full))]
    ids = [list(g["src_ids_flat"][off[i] : off[i + 1]]) for i in range(len(full))]
    return full, embeds, ids


def _ragged(flat, off):
    return [list(flat[off[i] : off[i + 1]]) for i in range(len(off) - 1)]


def test_collater_random_split_golden():
    g = load_golden("collater_random_split.npz")
    full, embeds, ids = _samples(g)
    split = pack_ref.draw_split_points(full, int(g["bi_output_embed_max_split_len"]), seed=int(g["seed"]))
    padded, mask, ids_out = pack_ref.collate_padded(embeds, "random_split", split_points=split, token_ids=ids)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    assert mask.dtype == np.int64
    assert ids_out == _ragged(g["ids_flat"], g["ids_off"])
    # packed layout contract
    packed, cu = pack_ref.pack_varlen(embeds, split)
    for i, n in enumerate(split):
        np.testing.assert_array_equal(packed[cu[i] : cu[i + 1]], g["out_embed_bits"][i, :n])
    back, mask2 = pack_ref.unpack_padded(packed, cu)
    np.testing.assert_array_equal(back, g["out_embed_bits"])
    np.testing.assert_array_equal(mask2, g["out_mask"])
    assert cu.dtype == np.int32 and np.array_equal(cu[1:], np.cumsum(g["out_mask"].sum(1)))


@pytest.mark.parametrize("name", ["collater_fixed_max.npz", "collater_fixed_max_uncapped.npz"])
def test_collater_fixed_max_golden(name):
    g = load_golden(name)
    full, embeds, ids = _samples(g)
    padded, mask, ids_out = pack_ref.collate_padded(embeds, "fixed_max", max_len=int(g["bi_output_embed_max_len"]), token_ids=ids)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    assert ids_out == _ragged(g["ids_flat"], g["ids_off"])


def test_collater_input_embed_golden():
    g = load_golden("collater_input_embed.npz")
    full, embeds, ids = _samples(g)
    padded, mask, _ = pack_ref.collate_padded(embeds, "fixed_max", max_len=int(g["bi_input_embed_max_len"]))
    np.testing.assert_array_equal(padded, g["in_embed_bits"])
    np.testing.assert_array_equal(mask, g["in_mask"])
    # the same call also ran the random-split branch on the output embeds (one randint per sample, batch order)
    split = pack_ref.draw_split_points(full, int(g["bi_output_embed_max_split_len"]), seed=int(g["seed"]))
    padded_o, mask_o, ids_o = pack_ref.collate_padded(embeds, "random_split", split_points=split, token_ids=ids)
    np.testing.assert_array_equal(padded_o, g["out_embed_bits"])
    np.testing.assert_array_equal(mask_o, g["out_mask"])
    assert ids_o == _ragged(g["ids_flat"], g["ids_off"])


def test_collater_edge_cases():
    with pytest.raises(ValueError):
        pack_ref.draw_split_points([1], 8, seed=0)  # the reference's randint(1, 0) raises
    packed, cu = pack_ref.pack_varlen([], [])
    assert packed.shape[0] == 0 and cu.tolist() == [0]
    e = [np.arange(12, dtype=np.uint16).reshape(3, 4)]
    padded, mask, _ = pack_ref.collate_padded(e, "fixed_max", max_len=8)
    assert padded.shape == (1, 3, 4) and mask.sum() == 3


# ------------------------------------------------------------------------------------------------ losses
def test_cross_entropy_golden():
    g = load_golden("ce_small.npz")
    loss, dz, n = loss_ref.cross_entropy_fwd_bwd(g["logits"], g["labels"])
    assert n == int((g["labels"] != -100).sum())
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-5)
    np.testing.assert_allclose(dz, g["dlogits"], rtol=1e-4, atol=1e-7)
    assert np.all(dz[g["labels"] == -100] == 0)


def test_cross_entropy_all_ignored_is_nan():
    loss, dz, n = loss_ref.cross_entropy_fwd_bwd(np.zeros((3, 8), np.float32), np.full(3, -100))
    assert np.isnan(loss) and n == 0 and not dz.any()


def test_masked_mse_matches_torch():
    rng = np.random.RandomState(0)
    y, t = rng.standard_normal((9, 16)).astype(np.float32), rng.standard_normal((9, 16)).astype(np.float32)
    valid = np.array([1, 1, 0, 1, 0, 1, 1, 1, 0], bool)
    yt = torch.from_numpy(y).requires_grad_(True)
    ref = torch.nn.functional.mse_loss(yt[torch.from_numpy(valid)].float(), torch.from_numpy(t)[torch.from_numpy(valid)].float())
    ref.backward()
    loss, dy, n = loss_ref.masked_mse_fwd_bwd(y, t, valid)
    np.testing.assert_allclose(loss, ref.item(), rtol=1e-6)
    np.testing.assert_allclose(dy, yt.grad.numpy(), rtol=1e-6, atol=1e-9)
    assert n == 6


def test_lm_head_ce_restatement_matches_reference_expression_golden():
    """oracle/t5_head_ref.py (frozen lm_head + CE + backward to the decoder output, section 8 f-1) vs the reference's expression
    executed by torch under CPU bf16 autocast (tests/golden/lm_head_ce_small.npz, written by oracle/make_golden.py)."""
    from oracle import t5_head_ref
    from oracle.golden import load_golden

    g = load_golden("lm_head_ce_small.npz")
    loss, dseq, logits = t5_head_ref.lm_head_ce_fwd_bwd(g["seq"], g["weight"], g["labels"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-5)
    # bf16 outputs: identical except where a different fp32 summation order flips the last bf16 bit
    assert np.abs(logits - g["logits"]).max() <= 2.0 ** -7 * np.abs(g["logits"]).max()
    assert (logits != g["logits"]).mean() < 5e-3
    assert np.linalg.norm(dseq - g["dseq"]) / np.linalg.norm(g["dseq"]) < 1e-3
    ignored = g["labels"] == -100
    assert ignored.any() and not dseq[ignored].any() and not g["dseq"][ignored].any()
    # the K / V projection shape: a frozen bias-free Linear and its input gradient
    y = t5_head_ref.frozen_linear_fwd(g["seq"], g["weight"])
    assert np.array_equal(y, logits)
