"""The oracle of the frozen T5 decoder stack on PACKED aligner rows (oracle/t5_decoder_ref.py) against the
third-party implementation the reference runs (HF transformers T5Stack, flan-t5 v1.1 flavour: gated GELU, no biases, relative
position bias, no attention scaling), live, on CPU fp32: outputs, and the gradient with respect to the encoder states -- what flows
back into the aligner -- on ragged batches. Also the relative-position buckets and the claim the f-1 design rests on: projecting
K / V on the packed rows and attending per sample equals HF's zero-padded batch + mask, and pad rows receive exactly no gradient."""
import numpy as np
import pytest
import torch

from oracle import t5_decoder_ref as ref

transformers = pytest.importorskip("transformers")
D, H, DKV, DFF, LAYERS = 64, 4, 16, 160, 3


def _hf_decoder(seed=0):
    from transformers import T5Config
    from transformers.models.t5.modeling_t5 import T5Stack

    cfg = T5Config(vocab_size=16, d_model=D, d_kv=DKV, d_ff=DFF, num_layers=LAYERS, num_decoder_layers=LAYERS, num_heads=H,
                   feed_forward_proj="gated-gelu", relative_attention_num_buckets=32, relative_attention_max_distance=128,
                   dropout_rate=0.0, is_decoder=True, use_cache=False, tie_word_embeddings=False)
    cfg.is_decoder = True
    torch.manual_seed(seed)
    dec = T5Stack(cfg).eval()
    with torch.no_grad():  # HF's default init leaves the norms at 1 and the bias table tiny: make every parameter matter
        for k, p in dec.named_parameters():
            p.copy_(torch.randn(p.shape) * (0.5 if "relative_attention_bias" in k else 0.08) + (1.0 if "layer_norm" in k else 0.0))
    return dec


@pytest.mark.parametrize("lens,T", [([7, 4], 5), ([1, 9, 3, 9], 12), ([6], 1), ([2, 2, 2], 40)])
def test_packed_decoder_stack_equals_hf_t5_on_the_padded_batch(lens, T):
    dec = _hf_decoder(seed=len(lens))
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    B, L = len(lens), max(lens)
    g = torch.Generator().manual_seed(3)
    dec_in = torch.randn((B, T, D), generator=g)
    enc_packed = torch.randn((sum(lens), D), generator=g, requires_grad=True)
    cu = np.concatenate([[0], np.cumsum(lens)])
    # HF: zero-padded encoder states + the collater's mask (reference ...embed_decoder_2.py:590-598)
    enc_padded = torch.zeros((B, L, D))
    mask = torch.zeros((B, L), dtype=torch.long)
    for b, n in enumerate(lens):
        enc_padded[b, :n], mask[b, :n] = enc_packed.detach()[cu[b] : cu[b + 1]], 1
    enc_padded.requires_grad_(True)
    want = dec(inputs_embeds=dec_in, encoder_hidden_states=enc_padded, encoder_attention_mask=mask, use_cache=False).last_hidden_state
    got = ref.decoder_stack_packed(sd, dec_in, enc_packed, cu, n_heads=H)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-5)
    # the gradient that reaches the aligner: same upstream on both, packed rows == valid rows of HF's padded gradient, pads get none
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    got.backward(up)
    for b, n in enumerate(lens):
        torch.testing.assert_close(enc_packed.grad[cu[b] : cu[b + 1]], enc_padded.grad[b, :n], rtol=1e-4, atol=2e-6)
        assert float(enc_padded.grad[b, n:].abs().max()) == 0.0 if n < L else True


def test_relative_position_buckets_and_bias_match_hf():
    from transformers.models.t5.modeling_t5 import T5Attention

    rel = torch.arange(-300, 50)[None, :].repeat(3, 1)
    want = T5Attention._relative_position_bucket(rel, bidirectional=False, num_buckets=32, max_distance=128)
    assert torch.equal(ref.relative_position_bucket(rel, 32, 128), want)
    assert int(want.min()) == 0 and int(want.max()) == 31
    dec = _hf_decoder()
    att = dec.block[0].layer[0].SelfAttention
    T = 33
    hf_bias = att.compute_bias(T, T)[0]  # [heads, T, T], no mask
    mine = ref.self_attention_bias(att.relative_attention_bias.weight.detach(), T)
    keep = torch.ones(T, T).tril().bool()
    torch.testing.assert_close(mine[:, keep], hf_bias[:, keep].detach())
    assert bool((mine[:, ~keep] < -1e30).all())  # causal part


def test_layer_norm_and_gelu_restatements_match_hf():
    from transformers.activations import ACT2FN
    from transformers.models.t5.modeling_t5 import T5LayerNorm

    x = torch.randn(5, 7, D) * 3
    n = T5LayerNorm(D)
    with torch.no_grad():
        n.weight.copy_(torch.randn(D))
    torch.testing.assert_close(ref.rms_norm(x, n.weight.detach()), n(x).detach(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ref.gelu_new(x), ACT2FN["gelu_new"](x), rtol=1e-6, atol=1e-6)
