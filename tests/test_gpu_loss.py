"""GPU parity: masked cross-entropy and masked MSE (forward + gradient in one pass) vs the numpy oracle and the
reference expression's golden vector. fp32 arithmetic: rtol 1e-5 on the loss, 1e-4 on gradients (exp/log intrinsics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cross_entropy_vs_reference_golden():
    from oracle.golden import load_golden
    from thinkdiff_mlre_b200 import ops

    g = load_golden("ce_small.npz")
    loss, dz = ops.masked_ce_fwd_bwd(torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(g["labels"]).cuda())
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(dz.cpu().numpy(), g["dlogits"], rtol=1e-4, atol=1e-7)
    assert not dz[torch.from_numpy(g["labels"]).cuda() == -100].any()  # ignore_index rows: exact zeros


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cross_entropy_t5_vocab_vs_oracle(dtype):
    """V = 32128 (Flan-T5-XXL vocab), logits [B*T, V] with T5 padding -> -100."""
    from oracle import loss_ref
    import thinkdiff_mlre_b200 as td

    rng = np.random.RandomState(1)
    R, V = 96, 32128
    z = torch.from_numpy((rng.standard_normal((R, V)) * 4).astype(np.float32)).to(dtype)
    labels = rng.randint(0, V, size=R).astype(np.int64)
    labels[rng.rand(R) < 0.4] = -100
    zt = z.cuda().requires_grad_(True)
    loss = td.masked_cross_entropy(zt.view(8, 12, V), torch.from_numpy(labels).cuda().view(8, 12))
    (loss * 65536.0).backward()  # GradScaler-style upstream scalar
    ref_loss, ref_dz, n = loss_ref.cross_entropy_fwd_bwd(z.float().numpy(), labels, grad_scale=65536.0)
    np.testing.assert_allclose(loss.item(), ref_loss, rtol=1e-5)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    err = np.linalg.norm(zt.grad.float().cpu().numpy() - ref_dz) / np.linalg.norm(ref_dz)
    assert err < tol and zt.grad.dtype == dtype
    # size-independent property: each valid row's gradient sums to ~0, ignored rows are exactly 0
    rowsum = zt.grad.float().sum(1).cpu().numpy()
    assert np.abs(rowsum).max() < (1e-2 if dtype == torch.float32 else 40.0)
    assert not zt.grad[torch.from_numpy(labels).cuda() == -100].any()


def test_cross_entropy_all_ignored_is_nan_like_torch():
    from thinkdiff_mlre_b200 import ops

    z = torch.randn(4, 64, device="cuda")
    loss, dz = ops.masked_ce_fwd_bwd(z, torch.full((4,), -100, dtype=torch.int64, device="cuda"))
    assert torch.isnan(loss) and not dz.any()


@pytest.mark.parametrize("ydt,tdt", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_masked_mse_vs_oracle(ydt, tdt):
    from oracle import loss_ref
    import thinkdiff_mlre_b200 as td

    rng = np.random.RandomState(2)
    M, D = 517, 4096
    y = torch.from_numpy(rng.standard_normal((M, D)).astype(np.float32)).to(ydt)
    t = torch.from_numpy(rng.standard_normal((M, D)).astype(np.float32)).to(tdt)
    mask = (rng.rand(M) < 0.7).astype(np.int64)
    for mk in (None, mask):
        loss, dy = td.ops.masked_mse_fwd_bwd(y.cuda(), t.cuda(), None if mk is None else torch.from_numpy(mk).cuda(), grad_scale=3.0)
        ref_loss, ref_dy, n = loss_ref.masked_mse_fwd_bwd(y.float().numpy(), t.float().numpy(), None if mk is None else mk.astype(bool), grad_scale=3.0)
        np.testing.assert_allclose(loss.item(), ref_loss, rtol=1e-5)
        assert dy.dtype == ydt
        err = np.linalg.norm(dy.float().cpu().numpy() - ref_dy) / np.linalg.norm(ref_dy)
        assert err < (1e-6 if ydt == torch.float32 else 5e-3)
        if mk is not None:
            assert not dy[torch.from_numpy(mk).cuda() == 0].any()
    # autograd form on a padded [B, L, D] batch with the collater's int64 mask
    yt = y[:512].reshape(8, 64, D).cuda().requires_grad_(True)
    m2 = torch.from_numpy(mask[:512]).reshape(8, 64).cuda()
    td.masked_mse(yt, t[:512].reshape(8, 64, D).cuda(), m2).backward()
    ref_loss, ref_dy, _ = loss_ref.masked_mse_fwd_bwd(y[:512].float().numpy(), t[:512].float().numpy(), mask[:512].astype(bool))
    err = np.linalg.norm(yt.grad.float().cpu().numpy().reshape(512, D) - ref_dy) / np.linalg.norm(ref_dy)
    assert err < (1e-6 if ydt == torch.float32 else 5e-3)
