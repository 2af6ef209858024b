"""GPU parity: the fused aligner (forward, backward, every numerical regime) through the module / C ABI vs
  (i)  the committed golden vectors produced by the reference's own build_vision_projector, and
  (ii) the closed-form CPU oracle on seeded ragged inputs.
Tolerances (north star): bf16 regime rtol 2e-2, fp32 regime rtol 1e-5 -- stated per test as relative Frobenius error
plus an elementwise assert_close whose atol is scaled by the reference tensor's magnitude (pure rtol is meaningless
for the near-zero entries of y and of the gradients)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PARAM_KEYS = ("0.weight", "0.bias", "2.weight", "2.bias", "3.weight")
ORACLE_NAME = {"0.weight": "dW1", "0.bias": "db1", "2.weight": "dW2", "2.bias": "db2", "3.weight": "dg"}
BF16_RTOL, FP32_RTOL = 2e-2, 1e-5


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def assert_close_scaled(a, b, rtol, what):
    a, b = a.detach().float().cpu(), torch.as_tensor(b).float().reshape(a.shape)
    scale = float(b.abs().max()) if rtol < 1e-3 else float(b.pow(2).mean().sqrt())
    torch.testing.assert_close(a, b, rtol=rtol, atol=rtol * 4 * scale, msg=lambda m: f"{what}: {m}")


def make_module(din, d, seed, device="cuda", dtype=None):
    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    params = aligner_ref.init_params_numpy(din, d, seed)
    m = td.ThinkDiffAligner(din, d).to(device)
    m.load_state_dict(params)
    if dtype is not None:
        m = m.to(dtype)
    return m, params


def golden_inputs(g):
    din, d, seed = int(g["din"]), int(g["d"]), int(g["seed"])
    rng = np.random.RandomState(seed + 1)
    x = rng.standard_normal(tuple(g["x_shape"])).astype(np.float32)
    if int(g["heavy_tail"]):
        x[..., rng.choice(din, size=8, replace=False)] *= 50.0
    t = rng.standard_normal(tuple(g["x_shape"][:-1]) + (d,)).astype(np.float32)
    return din, d, seed, torch.from_numpy(x), torch.from_numpy(t)


def run_train(m, x, t, autocast=True):
    import thinkdiff_mlre_b200 as td

    m.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
    else:
        y = m(x)
    loss = td.masked_mse(y, t)
    loss.backward()
    return y, loss, {k: p.grad for k, p in m.named_parameters()}


@pytest.mark.parametrize("name", ["aligner_mid_bf16.npz", "aligner_mid_bf16_heavy.npz", "aligner_cfg1_bf16.npz"])
def test_bf16_training_regime_vs_reference_golden(name):
    from oracle.golden import load_golden

    g = load_golden(name)
    din, d, seed, x, t = golden_inputs(g)
    m, _ = make_module(din, d, seed)
    y, loss, grads = run_train(m, x.cuda(), t.cuda())
    assert y.dtype == torch.float32 and y.shape == t.shape  # fp32 out under autocast (fp32 norm weight)
    ysel = y.reshape(-1, d)[torch.from_numpy(g["y_rows"]).cuda()]
    assert rel(ysel, g["y_sel"]) < BF16_RTOL
    assert_close_scaled(ysel, g["y_sel"], BF16_RTOL, "y rows")
    assert abs(float(loss) - float(g["loss"])) < BF16_RTOL * float(g["loss"])
    for k in PARAM_KEYS:
        assert grads[k].dtype == torch.float32
        if grads[k].ndim == 1:
            assert rel(grads[k], g["g_" + k]) < BF16_RTOL, k
        else:
            got = grads[k].reshape(-1)[torch.from_numpy(g["gi_" + k]).cuda()]
            assert rel(got, g["gs_" + k]) < BF16_RTOL, k
            assert abs(float(grads[k].double().norm()) - float(g["gn_" + k])) < BF16_RTOL * float(g["gn_" + k]), k


def test_small_golden_full_tensors_bf16():
    from oracle.golden import load_golden

    g = load_golden("aligner_small_bf16.npz")
    import thinkdiff_mlre_b200 as td

    m = td.ThinkDiffAligner(int(g["din"]), int(g["d"])).cuda()
    m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p_")})
    y, loss, grads = run_train(m, torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda())
    assert rel(y, g["y"]) < BF16_RTOL
    for k in PARAM_KEYS:
        assert rel(grads[k], g["g_" + k]) < BF16_RTOL, k
        assert_close_scaled(grads[k], g["g_" + k], BF16_RTOL, k)


@pytest.mark.parametrize("M", [1, 31, 128, 129, 257, 1000])
@pytest.mark.parametrize("din,d", [(192, 512), (768, 1024)])
def test_bf16_regime_vs_closed_form_oracle_ragged_rows(M, din, d):
    """Every rounding point restated on the CPU (oracle.aligner_fwd_bwd_manual); M crosses tile boundaries."""
    from oracle import aligner_ref

    m, params = make_module(din, d, seed=M + din)
    g = torch.Generator().manual_seed(M)
    x = torch.randn((M, din), generator=g).to(torch.bfloat16)
    dy = torch.randn((M, d), generator=g) / (M * d)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m.forward_packed(x.cuda())
    y.backward(dy.cuda())
    ref = aligner_ref.aligner_fwd_bwd_manual(x.float(), params, dy=dy, regime="bf16")
    assert rel(y, ref["y"]) < 5e-3  # same rounding points: far inside the 2e-2 budget
    assert_close_scaled(y, ref["y"], BF16_RTOL, "y")
    for k in PARAM_KEYS:
        assert rel(dict(m.named_parameters())[k].grad, ref[ORACLE_NAME[k]]) < BF16_RTOL, k


def test_bf16_inference_regime_output_dtype_and_rounding():
    from oracle import aligner_ref

    m, params = make_module(192, 512, seed=9, dtype=torch.bfloat16)
    x = torch.randn(77, 192).to(torch.bfloat16)
    with torch.no_grad():
        y = m(x.cuda().reshape(7, 11, 192))
    assert y.dtype == torch.bfloat16 and y.shape == (7, 11, 512)  # pure-bf16 rule
    p16 = {k: v.to(torch.bfloat16).float() for k, v in params.items()}
    ref = aligner_ref.aligner_fwd_bwd_manual(x.float(), p16, regime="bf16", out_bf16=True)
    assert rel(y.reshape(-1, 512), ref["y"]) < 5e-3
    # 1-D and 2-D call shapes of the reference's inference loops (...embed_decoder_2.py:998, :1115)
    with torch.no_grad():
        y2 = m(x.cuda())
    assert torch.equal(y2, y.reshape(-1, 512))


def test_no_grad_path_saves_nothing_and_matches_training_forward():
    m, _ = make_module(192, 512, seed=2)
    x = torch.randn(100, 192, device="cuda").to(torch.bfloat16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_train = m(x)
        with torch.no_grad():
            y_eval = m(x)
    assert y_train.requires_grad and not y_eval.requires_grad
    assert torch.equal(y_train.detach(), y_eval)


def test_backward_is_exactly_linear_in_power_of_two_loss_scale_and_propagates_inf():
    """GradScaler (runner_base.py:131-139) multiplies the loss by 65536: every rounding point scales exactly."""
    m, _ = make_module(192, 512, seed=4)
    x = torch.randn(200, 192, device="cuda").to(torch.bfloat16)
    dy = torch.randn(200, 512, device="cuda") * 1e-3
    grads = []
    for s in (1.0, 65536.0):
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        y.backward(dy * s)
        grads.append([p.grad.clone() for p in m.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a * 65536.0, b)
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    dy2 = dy.clone()
    dy2[3, 5] = float("inf")
    y.backward(dy2)
    assert not torch.isfinite(m[2].weight.grad).all()  # the scaler's inf check must see it


def test_regime_errors_match_reference_behaviour():
    m, _ = make_module(192, 512, seed=1)
    with pytest.raises(TypeError):  # bf16 input, fp32 weights, no autocast: F.linear raises in the reference too
        m(torch.zeros(4, 192, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        m(torch.zeros(4, 100, device="cuda"))
    with pytest.raises(NotImplementedError):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            m(torch.zeros(4, 192, device="cuda", requires_grad=True))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(torch.zeros(0, 192, device="cuda"))  # empty shard
    assert y.shape == (0, 512)


@pytest.mark.parametrize("name", ["aligner_small_fp32.npz", "aligner_mid_fp32.npz", "aligner_cfg1_fp32.npz"])
def test_fp32_regime_vs_reference_golden(name):
    """BASELINE config 1 (fp32, no autocast): tensor-core result via bf16x3 splitting must meet fp32 rtol 1e-5."""
    from oracle.golden import load_golden

    g = load_golden(name)
    import thinkdiff_mlre_b200 as td

    if "p_0.weight" in g:
        din, d = int(g["din"]), int(g["d"])
        m = td.ThinkDiffAligner(din, d).cuda()
        m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("p_")})
        x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    else:
        din, d, seed, x, t = golden_inputs(g)
        m, _ = make_module(din, d, seed)
    y, loss, grads = run_train(m, x.cuda(), t.cuda(), autocast=False)
    assert y.dtype == torch.float32
    if "y" in g:
        assert rel(y, g["y"]) < FP32_RTOL
        assert_close_scaled(y, g["y"], FP32_RTOL, "y")
    else:
        ysel = y.reshape(-1, d)[torch.from_numpy(g["y_rows"]).cuda()]
        assert rel(ysel, g["y_sel"]) < FP32_RTOL
        assert_close_scaled(ysel, g["y_sel"], FP32_RTOL, "y rows")
    assert abs(float(loss) - float(g["loss"])) < FP32_RTOL * float(g["loss"])
    for k in PARAM_KEYS:
        if ("g_" + k) in g:
            assert rel(grads[k], g["g_" + k]) < FP32_RTOL, k
            assert_close_scaled(grads[k], g["g_" + k], FP32_RTOL, k)
        else:
            got = grads[k].reshape(-1)[torch.from_numpy(g["gi_" + k]).cuda()]
            assert rel(got, g["gs_" + k]) < FP32_RTOL, k


def test_full_size_lvlm_shapes_sparse_probe():
    """BASELINE config 2 sizes (M = 8224 valid tokens, 3584 -> 4096). The oracle cannot run the full problem in seconds,
    so: forward rows are checked on a 48-row sample, and the backward with an upstream gradient that is non-zero on
    those rows only -- every gradient then depends on the sampled rows alone and the oracle recomputes it exactly."""
    from oracle import aligner_ref

    din, d, M = 3584, 4096, 8224
    m, params = make_module(din, d, seed=77)
    g = torch.Generator().manual_seed(5)
    x = torch.randn((M, din), generator=g).to(torch.bfloat16)
    rows = torch.sort(torch.randperm(M, generator=g)[:48]).values
    dy = torch.zeros((M, d))
    dy[rows] = torch.randn((48, d), generator=g) / 48
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m.forward_packed(x.cuda())
    y.backward(dy.cuda())
    ref = aligner_ref.aligner_fwd_bwd_manual(x[rows].float(), params, dy=dy[rows], regime="bf16", accum_dtype=torch.float32)
    assert rel(y[rows.cuda()], ref["y"]) < BF16_RTOL
    for k in PARAM_KEYS:
        assert rel(dict(m.named_parameters())[k].grad, ref[ORACLE_NAME[k]]) < BF16_RTOL, k
    # rows whose upstream gradient is zero contribute exactly nothing: rerun with those rows' features scrambled
    grads0 = [p.grad.clone() for p in m.parameters()]
    m.zero_grad(set_to_none=True)
    x2 = x.clone()
    keep = torch.zeros(M, dtype=torch.bool)
    keep[rows] = True
    x2[~keep] = torch.randn((M - 48, din), generator=g).to(torch.bfloat16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = m.forward_packed(x2.cuda())
    y2.backward(dy.cuda())
    for a, p in zip(grads0, m.parameters()):
        assert rel(p.grad, a.cpu()) < 1e-5


def test_two_shard_data_parallel_mean_equals_oracle_mean():
    """DDP semantics without a second GPU: run the two ranks' shards one after the other with grad_scale 1/2 (what
    enable_data_parallel does) and sum -- must equal the mean of the per-shard oracle gradients."""
    from oracle import aligner_ref
    from thinkdiff_mlre_b200 import ops

    m, params = make_module(192, 512, seed=6)
    g = torch.Generator().manual_seed(8)
    total = None
    want = None
    for M in (300, 77):  # ragged shards
        x = torch.randn((M, 192), generator=g).to(torch.bfloat16)
        t = torch.randn((M, 512), generator=g)
        W1b, b1b, W2b, b2b = m._bf16_params()
        y, saved = ops.aligner_fwd(x.cuda(), W1b, b1b, W2b, b2b, m[3].weight.detach(), 1e-6, False, True)
        loss, dy = ops.masked_mse_fwd_bwd(y, t.cuda())
        from thinkdiff_mlre_b200.aligner import GradBuckets

        gb = GradBuckets(192, 512, "cuda")
        bwd = ops.AlignerBackward(x.cuda(), saved, W2b, m[3].weight.detach(), dy, grad_scale=0.5)
        bwd.norm_and_linear2(gb.dW2, gb.db2, gb.dg)
        bwd.gelu_and_linear1(gb.dW1, gb.db1)
        got = [t_.clone() for t_ in gb.in_parameter_order()]
        total = got if total is None else [a + b for a, b in zip(total, got)]
        fwd = aligner_ref.aligner_fwd_bwd_manual(x.float(), params, regime="bf16")
        r = aligner_ref.aligner_fwd_bwd_manual(x.float(), params, dy=2 * (fwd["y"] - t) / t.numel(), regime="bf16")
        ref = [r[n] * 0.5 for n in ("dW1", "db1", "dW2", "db2", "dg")]
        want = ref if want is None else [a + b for a, b in zip(want, ref)]
    for a, b, n in zip(total, want, ("dW1", "db1", "dW2", "db2", "dg")):
        assert rel(a, b) < BF16_RTOL, n


def test_standalone_rmsnorm_fwd_bwd():
    from oracle import aligner_ref
    from thinkdiff_mlre_b200 import ops

    torch.manual_seed(3)
    M, D = 333, 1024
    x = (torch.randn(M, D) * 3).to(torch.bfloat16)
    gw = 1 + 0.1 * torch.randn(D)
    dy = torch.randn(M, D)
    y, rstd = ops.rmsnorm_fwd(x.cuda(), gw.cuda())
    norm = aligner_ref.T5RMSNorm(D)
    norm.weight.data.copy_(gw)
    xr = x.float().requires_grad_(True)
    yr = norm(xr)
    yr.backward(dy)
    assert rel(y, yr) < 1e-5
    dx, dg, dxsum = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), rstd, gw.cuda())
    assert dx.dtype == torch.bfloat16
    assert rel(dx, xr.grad) < 5e-3 and rel(dg, norm.weight.grad) < 1e-4
    assert rel(dxsum, dx.float().sum(0).cpu()) < 1e-5
    y16, _ = ops.rmsnorm_fwd(x.cuda(), gw.cuda(), out_bf16=True)
    ref16 = norm.to(torch.bfloat16)(x)
    assert y16.dtype == torch.bfloat16 and rel(y16, ref16.float()) < 5e-3


def test_weights_are_recast_after_a_fused_optimizer_step():
    """torch's fused AdamW updates parameters without bumping Tensor._version: the bf16 compute copies must not go stale."""
    from oracle import aligner_ref

    m, _ = make_module(192, 512, seed=12)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, fused=True)
    x = torch.randn(64, 192, device="cuda").to(torch.bfloat16)
    t = torch.randn(64, 512, device="cuda")
    for _ in range(3):
        run_train(m, x, t)
        opt.step()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    params = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ref = aligner_ref.aligner_fwd_bwd_manual(x.float().cpu(), params, regime="bf16")
    assert rel(y, ref["y"]) < 5e-3
    m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = m(x)
        m[0].weight.mul_(2.0)  # in-place edit bumps the version: the eval cache must notice
        y2 = m(x)
    assert torch.equal(y1, y.detach()) and not torch.equal(y1, y2)


@pytest.mark.parametrize("M,tdt", [(1, torch.float32), (200, torch.bfloat16), (515, torch.float32)])
def test_fused_mse_path_equals_module_boundary_path_and_oracle(M, tdt):
    """aligner.mse_loss_packed (y / dy never in HBM) vs forward -> masked_mse -> backward, and vs the closed-form oracle;
    the upstream scalar (GradScaler) is applied on the device."""
    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    m, params = make_module(192, 512, seed=21)
    g = torch.Generator().manual_seed(M)
    x = torch.randn((M, 192), generator=g).to(torch.bfloat16).cuda()
    t = torch.randn((M, 512), generator=g).to(tdt).cuda()
    # unfused, through the module boundary
    y, loss_ref_path, grads_ref = run_train(m, x, t.float())
    grads_ref = {k: v.clone() for k, v in grads_ref.items()}
    # fused
    m.zero_grad(set_to_none=True)
    loss = m.mse_loss_packed(x, t)
    scale = torch.tensor(1024.0, device="cuda")
    (loss * scale).backward()
    assert abs(float(loss) - float(loss_ref_path)) < 1e-5 * abs(float(loss_ref_path))
    for k, p in m.named_parameters():
        assert rel(p.grad / 1024.0, grads_ref[k].cpu()) < 2e-3, k  # same kernels bar the fused norm pass
    fwd = aligner_ref.aligner_fwd_bwd_manual(x.float().cpu(), params, regime="bf16")
    out = aligner_ref.aligner_fwd_bwd_manual(x.float().cpu(), params, dy=2 * (fwd["y"] - t.float().cpu()) / t.numel(), regime="bf16")
    for k in PARAM_KEYS:
        assert rel(dict(m.named_parameters())[k].grad / 1024.0, out[ORACLE_NAME[k]]) < BF16_RTOL, k
    ref_loss = float(((fwd["y"] - t.float().cpu()) ** 2).mean())
    assert abs(float(loss) - ref_loss) < 1e-3 * ref_loss


def test_fused_adamw_matches_torch_adamw_and_refreshes_bf16_copies():
    """td_adamw_step vs torch.optim.AdamW (the reference optimiser, runner_base.py:122-127) over several steps on the
    same gradients; and the bf16 compute copies it writes must equal a fresh cast of the updated parameters."""
    import copy

    import thinkdiff_mlre_b200 as td
    from thinkdiff_mlre_b200.train_step import make_reference_optimizer

    m1, _ = make_module(192, 512, seed=31)
    m2 = copy.deepcopy(m1)
    o1 = td.FusedAdamW(m1, lr=1e-3, weight_decay=0.05)
    o2 = make_reference_optimizer(m2, lr=1e-3, weight_decay=0.05)
    g = torch.Generator(device="cuda").manual_seed(0)
    for step in range(4):
        for (k, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
            p1.grad = torch.randn(p1.shape, generator=g, device="cuda") * 1e-2
            p2.grad = p1.grad.clone()
        o1.step(), o2.step()
        for (k, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
            torch.testing.assert_close(p1, p2, rtol=2e-6, atol=1e-7, msg=lambda s: f"{k} step {step}: {s}")
    for buf, p in zip(m1._bf16_buffers(), (m1[0].weight, m1[0].bias, m1[2].weight, m1[2].bias)):
        assert torch.equal(buf, p.detach().to(torch.bfloat16))
    # the next training forward must use those copies (no recast) and still be correct
    from thinkdiff_mlre_b200 import _lib as L

    x = torch.randn(50, 192, device="cuda").to(torch.bfloat16)
    n0 = L.launch_count
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = m1(x)
    assert L.launch_count - n0 == 3  # two GEMMs + norm, no cast kernels
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = m2(x)
    assert rel(y1, y2.float().cpu()) < 1e-3
    sd = o1.state_dict()
    assert len(sd["state"]) == 5 and sd["param_groups"][0]["weight_decay"] == 0.05


def test_pipelined_updates_equal_sequential_updates():
    """AlignerTrainStep(pipelined=True) applies step i's parameter updates inside step i+1 (Linear1's after the pack,
    Linear2's between the two forward GEMMs). Same arithmetic -> bit-identical parameters and losses."""
    import copy

    import thinkdiff_mlre_b200 as td

    m1, _ = make_module(192, 512, seed=41)
    m2 = copy.deepcopy(m1)
    s1 = td.AlignerTrainStep(m1, td.FusedAdamW(m1, lr=1e-3), pipelined=False)
    s2 = td.AlignerTrainStep(m2, td.FusedAdamW(m2, lr=1e-3), pipelined=True)
    batches = [td.synthetic_lvlm_batch(5, 40, 192, 512, seed=100 + j, pin=False) for j in range(3)]
    for j in range(5):
        b = batches[j % 3]
        args = (b.flat.cuda(), b.src_row_start.cuda(), b.lens.cuda(), b.total_rows, b.l_max, b.extras["flat_target"].cuda())
        l1, l2 = s1.step_device(*args), s2.step_device(*args)
        assert torch.equal(l1, l2), j
    s2.flush()
    for (k, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert torch.equal(p1, p2), k
    for b1, b2 in zip(m1._bf16_buffers(), m2._bf16_buffers()):
        assert torch.equal(b1, b2)
    assert all(p.grad is None for p in m2.parameters())


def test_no_out_of_bounds_writes_canary():
    """compute-sanitizer is closed on this GPU pool, so bounds are checked the hard way: every output / saved / workspace
    buffer of the forward and backward C-ABI calls sits between 4 KB guard zones filled with a sentinel; ragged M (not a
    multiple of any tile size) exercises the row-predicated epilogues and TMA out-of-bounds handling."""
    from thinkdiff_mlre_b200 import _lib as L

    dev = torch.device("cuda")
    M, din, d = 131, 192, 512
    guard = 4096
    sizes = {"h0": M * d * 2, "h1": M * d * 2, "h2": M * d * 2, "rstd": M * 4, "y": M * d * 4,
             "fws": L.lib().td_aligner_fwd_workspace_bytes(M, din, d), "bws": L.lib().td_aligner_bwd_workspace_bytes(M, din, d),
             "dW1": d * din * 4, "db1": d * 4, "dW2": d * d * 4, "db2": d * 4, "dg": d * 4}
    total = sum((s + 255) // 256 * 256 + guard for s in sizes.values()) + guard
    arena = torch.full((total,), 0xA5, dtype=torch.uint8, device=dev)
    off, view = guard, {}
    for k, s in sizes.items():
        view[k] = arena[off : off + s]
        off += (s + 255) // 256 * 256 + guard
    m, _ = make_module(din, d, seed=51)
    W1b, b1b, W2b, b2b = m._bf16_params()
    g = m[3].weight.detach()
    x = torch.randn(M, din, device=dev).to(torch.bfloat16)
    dy = torch.randn(M, d, device=dev)
    p = lambda t: L.ptr(t)  # noqa: E731
    L.check(L.lib().td_aligner_fwd(p(x), M, din, d, p(W1b), p(b1b), p(W2b), p(b2b), p(g), 1e-6, p(view["h0"]), p(view["h1"]),
                                   p(view["h2"]), p(view["rstd"]), p(view["y"]), L.F32, p(view["fws"]), sizes["fws"], L.stream_ptr()))
    L.check(L.lib().td_aligner_bwd(p(dy), L.F32, p(x), p(view["h0"]), p(view["h1"]), p(view["h2"]), p(view["rstd"]), p(W2b), p(g),
                                   M, din, d, 1.0, p(view["dW1"]), p(view["db1"]), p(view["dW2"]), p(view["db2"]), p(view["dg"]),
                                   p(view["bws"]), sizes["bws"], L.BWD_ALL, L.stream_ptr()))
    torch.cuda.synchronize()
    covered = torch.zeros(total, dtype=torch.bool, device=dev)
    off = guard
    for k, s in sizes.items():
        covered[off : off + s] = True
        off += (s + 255) // 256 * 256 + guard
    assert bool((arena[~covered] == 0xA5).all()), "a kernel wrote outside its buffers"
    # and the results written inside are the right ones
    y = view["y"].view(torch.float32).reshape(M, d)
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        y_ref = m(x)
    assert torch.equal(y, y_ref)


def test_fused_mse_reads_targets_through_row_index():
    """Targets left in the ragged source layout + the row index from the feature pack == packed targets."""
    import thinkdiff_mlre_b200 as td
    from thinkdiff_mlre_b200 import ops

    m, _ = make_module(192, 512, seed=61)
    b = td.synthetic_lvlm_batch(7, 60, 192, 512, seed=9, pin=False)
    flat, start, lens, tgt = b.flat.cuda(), b.src_row_start.cuda(), b.lens.cuda(), b.extras["flat_target"].cuda()
    cu = ops.cu_seqlens(lens)
    x, index = ops.pack_varlen(flat, start, cu, b.total_rows, want_index=True)
    t_packed = ops.pack_varlen(tgt, start, cu, b.total_rows)
    assert torch.equal(tgt[index].view(torch.int16), t_packed.view(torch.int16))
    l1 = m.mse_loss_packed(x, t_packed)
    l1.backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    m.zero_grad(set_to_none=True)
    l2 = m.mse_loss_packed(x, tgt, index)
    l2.backward()
    assert torch.equal(l1, l2)
    for a, p in zip(g1, m.parameters()):
        assert torch.equal(a, p.grad)
