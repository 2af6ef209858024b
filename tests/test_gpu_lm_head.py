"""GPU parity of SURVEY section 8 f-1's first slice: the frozen T5 output head + CrossEntropyLoss(ignore_index=-100) and a frozen
bias-free Linear (the cross-attention K / V projections) on packed rows, forward and backward to the input, through the C ABI.
Tolerances: loss rtol 1e-3 (bf16 logits), bf16 tensors relative Frobenius 2e-2 (BASELINE.json), typically 3e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_lm_head_ce_vs_reference_expression_golden():
    from oracle.golden import load_golden
    from thinkdiff_mlre_b200 import ops

    g = load_golden("lm_head_ce_small.npz")
    seq = torch.from_numpy(g["seq"]).to(torch.bfloat16).cuda()
    W = torch.from_numpy(g["weight"]).to(torch.bfloat16).cuda()
    labels = torch.from_numpy(g["labels"]).cuda()
    loss, dseq, logits = ops.lm_head_ce(seq, W, labels, keep_logits=True)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-3)
    assert _rel(logits.cpu(), torch.from_numpy(g["logits"])) < 2e-3
    assert _rel(dseq.cpu(), torch.from_numpy(g["dseq"])) < 2e-2
    assert not dseq[labels == -100].any()  # ignore_index rows: exact zeros
    # in-place gradient (the default) gives the same result as the two-buffer form
    loss2, dseq2, none = ops.lm_head_ce(seq, W, labels)
    assert none is None and torch.equal(dseq2, dseq) and float(loss2) == float(loss)
    # evaluation form: loss only
    loss3, no_grad, logits3 = ops.lm_head_ce(seq, W, labels, want_grad=False)
    assert no_grad is None and torch.equal(logits3, logits) and float(loss3) == float(loss)


def test_lm_head_ce_t5_xxl_shapes_vs_torch_and_oracle():
    """d_model 4096 -> vocab 32128 (Flan-T5-XXL): autograd module API vs the reference expression run by torch on the same GPU,
    and a sampled check of the logits against the CPU oracle."""
    import thinkdiff_mlre_b200 as td
    from oracle import t5_head_ref

    gen = torch.Generator(device="cuda").manual_seed(3)
    B, T, K, V = 6, 50, 4096, 32128
    W = (torch.randn((V, K), generator=gen, device="cuda") * 0.02)
    seq = torch.randn((B, T, K), generator=gen, device="cuda").to(torch.bfloat16)
    labels = torch.randint(0, V, (B, T), generator=gen, device="cuda")
    labels[torch.rand((B, T), generator=gen, device="cuda") < 0.35] = -100  # T5 padding -> -100 (...embed_decoder_2.py:577-581)
    s1 = seq.clone().requires_grad_(True)
    loss = td.lm_head_cross_entropy(s1, W, labels)
    (loss * 8.0).backward()
    s2 = seq.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = torch.nn.functional.linear(s2, W)
        ref = torch.nn.CrossEntropyLoss(ignore_index=-100)(logits.view(-1, V), labels.view(-1))
    (ref * 8.0).backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-3)
    assert s1.grad.dtype == torch.bfloat16 and s1.grad.shape == seq.shape
    assert _rel(s1.grad, s2.grad) < 2e-2
    assert not s1.grad[labels == -100].any()
    rows = [0, 77, 299]
    o_loss, _, o_logits = t5_head_ref.lm_head_ce_fwd_bwd(seq.view(-1, K)[rows].float().cpu().numpy(), W.cpu().numpy(),
                                                         labels.view(-1)[rows].cpu().numpy())
    _, _, mine = td.ops.lm_head_ce(seq.view(-1, K)[rows].contiguous(), W.to(torch.bfloat16), labels.view(-1)[rows].contiguous(),
                                   want_grad=False)
    assert _rel(mine.cpu(), torch.from_numpy(o_logits)) < 2e-3
    with pytest.raises(NotImplementedError):
        td.lm_head_cross_entropy(s1, W.clone().requires_grad_(True), labels)


@pytest.mark.parametrize("M", [1, 130, 2051])
def test_frozen_linear_packed_kv_projection(M):
    """[Wk; Wv] applied to PACKED aligner output rows and the gradient back to them == F.linear under autocast + autograd."""
    import thinkdiff_mlre_b200 as td
    from oracle import t5_head_ref

    gen = torch.Generator(device="cuda").manual_seed(M)
    K, inner = 512, 384
    Wkv = torch.randn((2 * inner, K), generator=gen, device="cuda") * 0.05
    y = torch.randn((M, K), generator=gen, device="cuda")  # the aligner's fp32 output under autocast
    dkv = torch.randn((M, 2 * inner), generator=gen, device="cuda").to(torch.bfloat16)
    y1 = y.clone().requires_grad_(True)
    kv = td.frozen_linear(y1, Wkv)
    kv.backward(dkv)
    y2 = y.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref = torch.nn.functional.linear(y2, Wkv)
    ref.backward(dkv)
    assert kv.dtype == torch.bfloat16 and y1.grad.dtype == torch.float32
    assert _rel(kv, ref) < 2e-3 and _rel(y1.grad, y2.grad) < 2e-2
    o = t5_head_ref.frozen_linear_fwd(y.cpu().numpy(), Wkv.cpu().numpy())
    assert _rel(kv.cpu(), torch.from_numpy(o)) < 2e-3
    odx = t5_head_ref.frozen_linear_dx(dkv.float().cpu().numpy(), Wkv.cpu().numpy())
    assert _rel(y1.grad.cpu(), torch.from_numpy(odx)) < 2e-3
