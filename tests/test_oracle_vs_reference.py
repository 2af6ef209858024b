"""Live differential test: the oracle against the reference's OWN functions, ast-extracted from /root/reference and
exec'd (oracle/ref_loader.py). Runs in the dev container only -- /root/reference does not exist on the GPU box, where
the committed golden vectors (tests/test_oracle_golden.py) pin the same thing."""
import random

import numpy as np
import pytest
import torch

from oracle import aligner_ref, pack_ref, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("din,d", [(64, 128), (768, 256)])
@pytest.mark.parametrize("autocast", [False, True])
def test_restatement_equals_reference_module(din, d, autocast):
    ref = ref_loader.build_reference_projector(din, d, "mlp2x_gelu_t5_norm")
    assert [type(m).__name__ for m in ref] == ["Linear", "GELU", "Linear", "T5LayerNorm"]
    mine = aligner_ref.RefAligner(din, d)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias", "3.weight"]
    mine.load_state_dict(ref.state_dict())
    torch.manual_seed(1)
    x, t = torch.randn(2, 9, din), torch.randn(2, 9, d)
    loss_fn = lambda y: torch.nn.functional.mse_loss(y, t)  # noqa: E731
    y0, l0, g0 = aligner_ref.module_fwd_bwd(ref, x, loss_fn, autocast)
    y1, l1, g1 = aligner_ref.module_fwd_bwd(mine, x, loss_fn, autocast)
    assert y0.dtype == y1.dtype == torch.float32
    assert torch.equal(y0, y1) and torch.equal(l0, l1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]) and g0[k].dtype == torch.float32


def test_reference_builder_type_strings():
    build = ref_loader.build_reference_projector
    assert type(build(8, 16, "linear")).__name__ == "Linear"
    assert len(build(8, 16, "mlp2x_gelu")) == 4  # Linear GELU Linear Identity
    assert len(build(8, 16, "mlp3x_gelu_t5_norm")) == 7  # a norm after EVERY extra Linear
    with pytest.raises(ValueError):
        build(8, 16, "conv")
    for t in ("mlp2x_gelu", "mlp3x_gelu_t5_norm", "linear"):
        a, b = build(8, 16, t), aligner_ref.build_ref_projector(8, 16, t)
        assert list(a.state_dict().keys()) == list(b.state_dict().keys())


def test_bf16_inference_regime_matches_reference():
    ref = ref_loader.build_reference_projector(64, 128).to(torch.bfloat16)
    x = torch.randn(11, 64).to(torch.bfloat16)
    y = ref(x)
    assert y.dtype == torch.bfloat16
    params = {k: v.float() for k, v in ref.state_dict().items()}
    out = aligner_ref.aligner_fwd_bwd_manual(x.float(), params, regime="bf16", out_bf16=True)
    err = (out["y"] - y.float()).norm() / y.float().norm()
    assert err < 2e-2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_collater_random_split_equals_reference(seed):
    collater = ref_loader.load_collater()
    rng = np.random.RandomState(seed)
    lens = [int(v) for v in rng.randint(2, 60, size=7)]
    samples, embeds, ids = [], [], []
    for L in lens:
        e = torch.randn(L, 24).to(torch.bfloat16)
        i = [int(v) for v in rng.randint(0, 1000, size=L)]
        embeds.append(e.view(torch.int16).numpy().view(np.uint16)), ids.append(i)
        samples.append({"json": {"generated_text": "t", "output_token_ids": i}, "a.input_embed.pth": e, "a.output_embed.pth": e})
    bi = dict(use_input_embed=False, use_output_embed=True, random_split_output_embed=True, output_embed_max_split_len=20,
              output_embed_max_len=32, input_embed_max_len=32)
    random.seed(seed)
    out = collater(bi, samples)
    split = pack_ref.draw_split_points(lens, 20, seed=seed)
    padded, mask, ids_out = pack_ref.collate_padded(embeds, "random_split", split_points=split, token_ids=ids)
    assert np.array_equal(out["a.output_embed"].view(torch.int16).numpy().view(np.uint16), padded)
    assert np.array_equal(out["output_embed_mask"].numpy(), mask) and out["output_embed_mask"].dtype == torch.int64
    assert out["output_token_ids"] == ids_out
    assert split == out["output_embed_mask"].sum(1).tolist()


def test_collater_fixed_max_equals_reference():
    collater = ref_loader.load_collater()
    lens = [3, 40, 17, 9]
    samples, embeds, ids = [], [], []
    for L in lens:
        e = torch.randn(L, 8).to(torch.bfloat16)
        i = list(range(L))
        embeds.append(e.view(torch.int16).numpy().view(np.uint16)), ids.append(i)
        samples.append({"json": {"generated_text": "t", "output_token_ids": i}, "a.input_embed.pth": e, "a.output_embed.pth": e})
    for cap in (12, 100):
        bi = dict(use_input_embed=True, use_output_embed=True, random_split_output_embed=False, output_embed_max_split_len=20,
                  output_embed_max_len=cap, input_embed_max_len=cap - 2)
        out = collater(bi, samples)
        padded, mask, ids_out = pack_ref.collate_padded(embeds, "fixed_max", max_len=cap, token_ids=ids)
        assert np.array_equal(out["a.output_embed"].view(torch.int16).numpy().view(np.uint16), padded)
        assert np.array_equal(out["output_embed_mask"].numpy(), mask)
        assert out["output_token_ids"] == ids_out
        padded_i, mask_i, _ = pack_ref.collate_padded(embeds, "fixed_max", max_len=cap - 2)
        assert np.array_equal(out["a.input_embed"].view(torch.int16).numpy().view(np.uint16), padded_i)
        assert np.array_equal(out["input_embed_mask"].numpy(), mask_i)
