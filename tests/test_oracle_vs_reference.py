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


@pytest.mark.parametrize("case", range(24))
def test_product_collater_and_shard_reader_equal_the_live_reference_collater(case, tmp_path):
    """Fuzz of the PRODUCT's host side of a-1 / f-2 against the reference collater executed live: random batch sizes, lengths,
    widths and build_info (random split / fixed max, capped or not, one or both embed streams, pass-through json fields), through
    FlatCollater (with and without host-side truncation) and through a flat shard + EmbedShardReader. The device pack is stood
    in for by the numpy pack oracle (the GPU tests check the kernel against that same oracle bit for bit)."""
    import thinkdiff_mlre_b200 as td

    collater = ref_loader.load_collater()
    rng = np.random.RandomState(1000 + case)
    B, C = int(rng.randint(1, 9)), int(rng.choice([8, 24, 64]))
    split = bool(case % 2)
    lens = [int(v) for v in rng.randint(2, 70, size=B)]
    both = case % 3 == 0
    with_gpt, with_rev = bool(rng.randint(2)), bool(rng.randint(2))
    samples = []
    for i, L in enumerate(lens):
        bits = rng.randint(0, 65536, size=(2, L, C)).astype(np.uint16)
        bits[(bits & 0x7F80) == 0x7F80] = 0x3F80  # no NaN / inf patterns: the reference pads with F.pad on real bf16 values
        eo, ei = (torch.from_numpy(b.view(np.int16)).view(torch.bfloat16) for b in bits)
        js = {"generated_text": f"t{i}", "output_token_ids": [int(v) for v in rng.randint(0, 32000, size=L)]}
        if with_gpt:
            js["gpt"] = f"g{i}"
        if with_rev:
            js["revised_generated_text"] = f"r{i}"
        samples.append({"__key__": f"k{i}", "json": js, "m.input_embed.pth": ei, "m.output_embed.pth": eo})
    bi = dict(use_input_embed=both, use_output_embed=True, random_split_output_embed=split,
              output_embed_max_split_len=int(rng.randint(1, 40)), output_embed_max_len=int(rng.randint(1, 90)),
              input_embed_max_len=int(rng.randint(1, 90)))
    random.seed(case)
    ref = collater(bi, samples)

    def padded_of(fb):
        packed, cu = pack_ref.pack_from_flat(fb.flat.view(torch.int16).numpy().view(np.uint16), fb.src_row_start.tolist(), fb.lens.tolist())
        return pack_ref.unpack_padded(packed, cu, fb.l_max)

    def check(fb, key, mask_key):
        padded, mask = padded_of(fb)
        assert np.array_equal(ref[key].view(torch.int16).numpy().view(np.uint16), padded)
        assert np.array_equal(ref[mask_key].numpy(), mask)

    for trunc in (False, True):
        random.seed(case)
        fb = td.FlatCollater(bi, pin_memory=False, truncate_on_host=trunc)(samples)
        check(fb, "m.output_embed", "output_embed_mask")
        assert fb.extras["output_token_ids"] == ref["output_token_ids"] and fb.extras["generated_texts"] == ref["generated_texts"]
        assert fb.extras.get("llava_gpts") == ref.get("llava_gpts") and fb.extras.get("revised_generated_texts") == ref.get("revised_generated_texts")
        if both:
            check(fb.extras["input_batch"], "m.input_embed", "input_embed_mask")
        else:
            assert "input_batch" not in fb.extras and "m.input_embed" not in ref
    path = str(tmp_path / "s.tdemb")
    with td.EmbedShardWriter(path, C) as w:
        for s in samples:
            w.add_reference_sample(s, "output")
    r = td.EmbedShardReader(path)
    for trunc in (False, True):
        random.seed(case)
        fb = r.batch(0, B, bi, pin_memory=False, truncate_on_host=trunc)
        check(fb, "m.output_embed", "output_embed_mask")
        assert fb.extras["output_token_ids"] == ref["output_token_ids"] and fb.extras.get("llava_gpts") == ref.get("llava_gpts")
        assert fb.extras.get("revised_generated_texts") == ref.get("revised_generated_texts")
    r.close()


@pytest.mark.parametrize("beta2", [None, 0.98])
def test_optimizer_groups_and_hyperparameters_equal_the_live_reference_runner(beta2):
    """FusedAdamW / make_reference_optimizer build the parameter groups of the reference runner's own ``optimizer`` property
    (runner_base.py:98-127, executed live on a model that holds the aligner as ``mm_projector`` next to a frozen weight):
    same parameters per group, same weight decay, lr, betas and eps."""
    import thinkdiff_mlre_b200 as td
    from thinkdiff_mlre_b200.train_step import reference_param_groups

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mm_projector = td.ThinkDiffAligner(64, 128)
            self.frozen = torch.nn.Linear(4, 4)
            for p in self.frozen.parameters():
                p.requires_grad_(False)

    model = Model()
    ref = ref_loader.build_reference_optimizer(model, init_lr=3e-4, weight_decay=0.05, beta2=beta2)
    assert type(ref) is torch.optim.AdamW
    kw = {} if beta2 is None else {"betas": (0.9, beta2)}
    mine = td.FusedAdamW(model.mm_projector, lr=3e-4, weight_decay=0.05, **kw)
    plain = reference_param_groups(model, 0.05)
    assert len(ref.param_groups) == len(mine.param_groups) == len(plain) == 2
    for g_ref, g_mine, g_plain in zip(ref.param_groups, mine.param_groups, plain):
        ids = [id(p) for p in g_ref["params"]]
        assert ids == [id(p) for p in g_mine["params"]] == [id(p) for p in g_plain["params"]]
        assert float(g_ref["weight_decay"]) == float(g_mine["weight_decay"]) == float(g_plain["weight_decay"])
        for k in ("lr", "betas", "eps"):
            assert tuple(g_ref[k]) == tuple(g_mine[k]) if k == "betas" else g_ref[k] == g_mine[k]
    assert not any(id(p) in {id(q) for g in ref.param_groups for q in g["params"]} for p in model.frozen.parameters())
