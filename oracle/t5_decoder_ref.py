"""CPU restatement of the frozen T5 v1.1 decoder stack that consumes the aligner's output (the part
behind the built lm_head slice).  TEST INFRASTRUCTURE ONLY -- the product path never imports this, and no kernel for it exists
yet: this file pins the arithmetic the next kernels (varlen cross-attention on PACKED aligner rows, gated-GELU FFN, relative
position bias) will be tested against.

Reference: ``T5ForDecoder.forward`` hands the aligner output to the HF decoder as ``encoder_hidden_states`` with the collater's mask
as ``encoder_attention_mask`` (thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:211-224, :590-598).  The decoder itself is
third-party and not vendored: ``transformers==4.46.1`` (requirements.txt:14) ``models/t5/modeling_t5.py`` -- ``T5Stack`` /
``T5Block`` / ``T5Attention`` / ``T5DenseGatedActDense`` / ``T5LayerNorm``; google/flan-t5-xxl: 24 blocks, d_model 4096, 64 heads
of d_kv 64, d_ff 10240, gated GELU (``gelu_new``, the tanh form), 32 relative-position buckets up to distance 128, no biases.
Its published algorithm, per block on decoder states ``h [B, T, d]`` (pre-norm residual everywhere, dropout off: frozen, eval):

    h += O_s . softmax(Q_s(n0(h)) K_s(n0(h))^T + rel_bias[T, T] + causal) V_s(n0(h))          # NO 1/sqrt(d_kv) scaling in T5
    h += O_c . softmax(Q_c(n1(h)) K_c(enc)^T + encoder_mask) V_c(enc)                           # no position bias in cross-attn
    h += Wo( gelu_new(Wi0(n2(h))) * Wi1(n2(h)) )
    out = final_norm(h)                                  n* = T5LayerNorm (RMS norm, no mean subtraction, weight only)

``rel_bias`` comes from block 0's ``relative_attention_bias`` table and is shared by all blocks; softmax is taken in fp32.

What this file adds to the published algorithm is the layout: the encoder states arrive PACKED (``enc [M, d]`` +
``cu_seqlens [B + 1]``, what ``ThinkDiffAligner.forward_packed`` produces).  K_c / V_c are projected once per block on the packed
rows (the built ``td_linear_bf16``) and sample ``b`` attends to rows ``cu[b]:cu[b+1]`` only -- no un-pack, no pad rows, no mask.
Parity is pinned LIVE against the installed ``transformers`` (5.5.0 here; the T5 modules' forward is unchanged from 4.46.1 apart
from the cache plumbing, which this path does not use) in tests/test_t5_decoder_oracle.py: outputs and the gradient with respect
to the encoder states (what flows back into the aligner) equal HF's on the zero-padded batch + mask.
"""
from __future__ import annotations

import math

import torch


def rms_norm(h: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """T5LayerNorm: fp32 mean of squares, no mean subtraction, no bias."""
    var = h.float().pow(2).mean(-1, keepdim=True)
    return w * (h * torch.rsqrt(var + eps)).to(w.dtype if w.dtype in (torch.float16, torch.bfloat16) else h.dtype)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def relative_position_bucket(rel: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """Decoder (unidirectional) buckets of ``rel = key_position - query_position``: keys ahead of the query share bucket 0 (they are
    masked anyway), the first half of the buckets counts exact distances, the second half grows logarithmically to max_distance."""
    n = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return torch.where(n < max_exact, n, large)


def self_attention_bias(table: torch.Tensor, T: int, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """[heads, T, T]: relative position bias of block 0 (shared by every block) plus the causal mask."""
    pos = torch.arange(T)
    bucket = relative_position_bucket(pos[None, :] - pos[:, None], num_buckets, max_distance)
    bias = table[bucket].permute(2, 0, 1)  # table [buckets, heads]
    causal = torch.full((T, T), torch.finfo(table.dtype).min).triu(1)
    return bias + causal


def _heads(x: torch.Tensor, n_heads: int) -> torch.Tensor:  # [L, H*dk] -> [H, L, dk]
    return x.view(x.shape[0], n_heads, -1).transpose(0, 1)


def _attend(q, k, v, bias=None):
    s = q @ k.transpose(-1, -2)  # T5: no 1/sqrt(d_kv)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s.float(), dim=-1).to(s.dtype)
    return (p @ v).transpose(0, 1).reshape(q.shape[1], -1)


def decoder_stack_packed(sd: dict, dec_embeds: torch.Tensor, enc_packed: torch.Tensor, cu_seqlens, n_heads: int, num_buckets: int = 32,
                         max_distance: int = 128, eps: float = 1e-6) -> torch.Tensor:
    """T5 decoder stack (state dict ``sd`` in HF ``T5Stack`` naming) on ``dec_embeds [B, T, d]`` attending to the packed encoder rows
    ``enc_packed [M, d]`` of the samples delimited by ``cu_seqlens``. Returns the final-normed decoder states ``[B, T, d]``."""
    B, T, _ = dec_embeds.shape
    cu = [int(c) for c in cu_seqlens]
    n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("block."))
    bias = self_attention_bias(sd["block.0.layer.0.SelfAttention.relative_attention_bias.weight"], T, num_buckets, max_distance)
    h = dec_embeds
    for i in range(n_blocks):
        p = f"block.{i}.layer."
        # self-attention
        x = rms_norm(h, sd[p + "0.layer_norm.weight"], eps)
        a = p + "0.SelfAttention."
        q, k, v = x @ sd[a + "q.weight"].t(), x @ sd[a + "k.weight"].t(), x @ sd[a + "v.weight"].t()
        att = torch.stack([_attend(_heads(q[b], n_heads), _heads(k[b], n_heads), _heads(v[b], n_heads), bias) for b in range(B)])
        h = h + att @ sd[a + "o.weight"].t()
        # cross-attention over the PACKED encoder rows: K / V projected once for all M rows, sample b sees rows cu[b]:cu[b+1] only
        x = rms_norm(h, sd[p + "1.layer_norm.weight"], eps)
        c = p + "1.EncDecAttention."
        q = x @ sd[c + "q.weight"].t()
        k_all, v_all = enc_packed @ sd[c + "k.weight"].t(), enc_packed @ sd[c + "v.weight"].t()
        att = torch.stack([_attend(_heads(q[b], n_heads), _heads(k_all[cu[b] : cu[b + 1]], n_heads), _heads(v_all[cu[b] : cu[b + 1]], n_heads))
                           for b in range(B)])
        h = h + att @ sd[c + "o.weight"].t()
        # gated-GELU feed-forward
        x = rms_norm(h, sd[p + "2.layer_norm.weight"], eps)
        f = p + "2.DenseReluDense."
        h = h + (gelu_new(x @ sd[f + "wi_0.weight"].t()) * (x @ sd[f + "wi_1.weight"].t())) @ sd[f + "wo.weight"].t()
    return rms_norm(h, sd["final_layer_norm.weight"], eps)
