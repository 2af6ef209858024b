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
@pytest.mark.parametrize("M,N,K,stream_k", [(128, 256, 64, False), (333, 768, 1096, True), (1000, 1024, 2048, True),
                                            (4096, 512, 8224, False), (4096, 512, 8224, True), (2600, 4096, 1024, True)])
def test_gemm_variants(pair, a_mn, b_mn, M, N, K, stream_k):
    import thinkdiff_mlre_b200 as td

    torch.manual_seed(M + N + K)
    dev = "cuda"
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8  # an MN-major operand's row pitch (= rows * 2 bytes) must be a multiple of 16
    A = (torch.rand((K, M) if a_mn else (M, K), device=dev) * 2 - 1).to(torch.bfloat16)
    B = (torch.rand((K, N) if b_mn else (N, K), device=dev) * 2 - 1).to(torch.bfloat16)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = td.ops.gemm_f32out(A, B, a_mn, b_mn, alpha=0.5, cta_pair=pair, stream_k=stream_k)
    ref = 0.5 * _ref(A, B, a_mn, b_mn)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-3 * K**0.5


def test_stream_k_tail_is_deterministic_and_close_to_the_wave_schedule():
    """The stream-K tail (leftover tiles cut along K, partial accumulators added in worker order) must give the same bits on
    every run, and differ from the plain wave schedule only by fp32 re-association of the contraction."""
    import thinkdiff_mlre_b200 as td

    torch.manual_seed(5)
    M, N, K = 4096, 3584, 1500  # the dW1 shape with a short token dimension: 224 tiles = 3 waves of 74 pairs + 2 tail tiles
    A = torch.randn((K, M), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N), device="cuda").to(torch.bfloat16)
    a = td.ops.gemm_f32out(A, B, True, True, stream_k=True)
    b = td.ops.gemm_f32out(A, B, True, True, stream_k=True)
    c = td.ops.gemm_f32out(A, B, True, True, stream_k=False)
    assert torch.equal(a, b)
    torch.testing.assert_close(a, c, rtol=1e-5, atol=1e-3)


def test_gemm_accumulate_adds_into_the_output():
    import thinkdiff_mlre_b200 as td

    torch.manual_seed(6)
    A = torch.randn((300, 256), device="cuda").to(torch.bfloat16)
    B = torch.randn((512, 256), device="cuda").to(torch.bfloat16)
    first = td.ops.gemm_f32out(A, B, False, False, alpha=0.5)
    out = first.clone()
    td.ops.gemm_f32out(A, B, False, False, alpha=0.5, out=out)
    torch.testing.assert_close(out, 2 * first, rtol=1e-6, atol=1e-6)


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
