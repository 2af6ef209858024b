"""GPU parity: the tcgen05 GEMM mainloop (every operand-major / CTA-pair variant) vs an fp32 torch matmul of the same
bf16 inputs. Tolerance: fp32 accumulation of bf16 products, |err| <= 1e-3 * sqrt(K) absolute (inputs in [-1, 1])."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_mn, b_mn):
    a = A.float().t() if a_mn else A.float()
    b = B.float().t() if b_mn else B.float()
    return a @ b.t()


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 256, 64, 1), (333, 768, 1096, 1), (1000, 1024, 2048, 2), (4096, 512, 8224, 0)])
def test_gemm_variants(pair, a_mn, b_mn, M, N, K, splits):
    import thinkdiff_mlre_b200 as td

    torch.manual_seed(M + N + K)
    dev = "cuda"
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8  # an MN-major operand's row pitch (= rows * 2 bytes) must be a multiple of 16
    A = (torch.rand((K, M) if a_mn else (M, K), device=dev) * 2 - 1).to(torch.bfloat16)
    B = (torch.rand((K, N) if b_mn else (N, K), device=dev) * 2 - 1).to(torch.bfloat16)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = td.ops.gemm_f32out(A, B, a_mn, b_mn, alpha=0.5, cta_pair=pair, splits=splits)
    ref = 0.5 * _ref(A, B, a_mn, b_mn)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-3 * K**0.5


def test_gemm_rejects_bad_shapes():
    import thinkdiff_mlre_b200 as td

    A = torch.zeros((64, 64), dtype=torch.bfloat16, device="cuda")
    B = torch.zeros((40, 64), dtype=torch.bfloat16, device="cuda")  # N = 40 is not a multiple of 32
    with pytest.raises(RuntimeError, match="multiple of 32"):
        td.ops.gemm_f32out(A, B, False, False)
    with pytest.raises(RuntimeError, match="16-byte"):
        td.ops.gemm_f32out(torch.zeros((64, 68), dtype=torch.bfloat16, device="cuda")[:, :60].contiguous(),
                           torch.zeros((64, 60), dtype=torch.bfloat16, device="cuda"), False, False)


def test_linear_bf16_bias():
    import thinkdiff_mlre_b200 as td

    torch.manual_seed(0)
    x = torch.randn(300, 192, device="cuda").to(torch.bfloat16)
    W = (torch.randn(512, 192, device="cuda") / 14).to(torch.bfloat16)
    b = torch.randn(512, device="cuda").to(torch.bfloat16)
    out = td.ops.linear_bf16(x, W, b)
    ref = (x.float() @ W.float().t() + b.float()).to(torch.bfloat16)
    assert out.dtype == torch.bfloat16
    torch.testing.assert_close(out.float(), ref.float(), rtol=2e-2, atol=2e-2)
