"""GPU parity of the optimizer-side rows of SURVEY section 8 (f-3): the direct (no-autograd) train step, the device-side
GradScaler (inf check / skipped step / scale growth and back-off), global-norm clipping and gradient accumulation -- each against
the reference loop's own building blocks (thinkdiff/tasks/base_task.py:241-258: ``scaler.scale(loss).backward()``,
``scaler.unscale_`` + ``clip_grad_norm_``, ``scaler.step``, ``scaler.update``, ``accum_grad_iters``) running on the SAME
gradients (this repo's autograd path) and the SAME update arithmetic (FusedAdamW driven from the host, itself parity-tested
against torch.optim.AdamW in test_gpu_aligner), so the comparison isolates the step logic. That matters: bf16 training is
chaotic -- a 1e-7 difference in an fp32 master flips a few bf16 roundings and shows up as 1e-3 in the next step's gradients --
so multi-step comparisons are only meaningful when each step is reproduced to (almost) the bit. Tolerance: 1e-6 relative."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

DIN, D = 192, 512


def _models():
    import thinkdiff_mlre_b200 as td
    from oracle import aligner_ref

    m1 = td.ThinkDiffAligner(DIN, D).cuda()
    m1.load_state_dict(aligner_ref.init_params_numpy(DIN, D, seed=9))
    m2 = copy.deepcopy(m1)
    return td, m1, m2


def _batches(td, n, poison=None):
    out = []
    for j in range(n):
        b = td.synthetic_lvlm_batch(5, 60, DIN, D, seed=50 + j, pin=False)
        tgt = b.extras["flat_target"].cuda()
        if poison is not None and j == poison:
            tgt = tgt.clone()
            tgt[int(b.src_row_start[0]), 3] = float("inf")  # a kept row: the loss and every gradient become non-finite
        out.append((b.flat.cuda(), b.src_row_start.cuda(), b.lens.cuda(), b.total_rows, b.l_max, tgt))
    return out


def _packed(td, batch):
    flat, start, lens, total, _, tgt = batch
    cu = td.ops.cu_seqlens(lens)
    x, idx = td.ops.pack_varlen(flat, start, cu, total, want_index=True)
    return x, tgt, idx


def _assert_params_close(m1, m2, rtol=2e-6):
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        assert err <= rtol, f"{n}: {err:.3e}"


def test_direct_step_gives_the_same_bits_as_the_autograd_path():
    td, m1, m2 = _models()
    x, tgt, idx = _packed(td, _batches(td, 1)[0])
    loss1 = m1.mse_loss_packed(x, tgt, idx)
    loss1.backward()
    loss2 = m2.mse_loss_backward_packed(x, tgt, idx)
    torch.cuda.synchronize()
    assert torch.equal(loss1.detach(), loss2)
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(a.grad, b.grad)


def _reference_loop(td, m, batches, scaler=None, max_norm=0.0, accum=1):
    """The reference task loop (base_task.py:236-258) on the autograd path, host-driven optimizer."""
    opt = td.FusedAdamW(m, lr=1e-3)
    norms = []
    for i, batch in enumerate(batches):
        x, tgt, idx = _packed(td, batch)
        loss = m.mse_loss_packed(x, tgt, idx)
        (scaler.scale(loss) if scaler is not None else loss).backward()
        if (i + 1) % accum == 0:
            if max_norm > 0.0:
                if scaler is not None:
                    scaler.unscale_(opt)
                norms.append(float(torch.nn.utils.clip_grad_norm_(list(m.parameters()), max_norm)))
            if scaler is not None:
                scaler.step(opt)
                scaler.update()
            else:
                opt.step()
            opt.zero_grad()
    return norms


def test_device_grad_scaler_matches_torch_grad_scaler():
    """Six steps, the third with a non-finite loss: both loops must skip it (parameters and Adam step count unchanged), halve the
    scale, and double it again after `growth_interval` clean steps -- with no host synchronisation on the device side."""
    td, m1, m2 = _models()
    batches = _batches(td, 6, poison=2)
    ref_scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2)
    _reference_loop(td, m1, batches, scaler=ref_scaler)

    sc = td.DeviceGradScaler(torch.device("cuda"), init_scale=1024.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2)
    step = td.AlignerTrainStep(m2, td.FusedAdamW(m2, lr=1e-3), grad_scaler=sc)
    skipped_at = None
    for i, b in enumerate(batches):
        before = [p.detach().clone() for p in m2.parameters()]
        step.step_device(*b)
        if all(torch.equal(a, p) for a, p in zip(before, m2.parameters())):
            skipped_at = i
    torch.cuda.synchronize()
    st = sc.state()
    assert skipped_at == 2                       # exactly the poisoned step left every parameter untouched
    assert st["scale"] == ref_scaler.get_scale() == 2048.0
    assert st["step"] == 5.0                     # the applied-step count (Adam's bias correction) did not advance on the skip
    _assert_params_close(m1, m2, rtol=1e-6)


def test_grad_clipping_matches_clip_grad_norm():
    """One step each (so no chaos between steps), with and without a scaler: same norm as torch reports, same parameters."""
    for with_scaler in (False, True):
        td, m1, m2 = _models()
        batches = _batches(td, 1)
        ref_scaler = torch.amp.GradScaler("cuda", init_scale=256.0) if with_scaler else None
        norms = _reference_loop(td, m1, batches, scaler=ref_scaler, max_norm=0.02)
        sc = td.DeviceGradScaler(torch.device("cuda"), init_scale=256.0) if with_scaler else None
        step = td.AlignerTrainStep(m2, td.FusedAdamW(m2, lr=1e-3), grad_scaler=sc, max_grad_norm=0.02)
        step.step_device(*batches[0])
        torch.cuda.synchronize()
        got = step.grad_scaler.state()["grad_norm"]
        assert norms[0] > 0.02                    # the clip is active
        assert abs(got - norms[0]) <= 1e-5 * norms[0]
        _assert_params_close(m1, m2, rtol=2e-6)


def test_gradient_accumulation_matches_two_backward_passes():
    td, m1, m2 = _models()
    batches = _batches(td, 4)
    _reference_loop(td, m1, batches, accum=2)
    step = td.AlignerTrainStep(m2, td.FusedAdamW(m2, lr=1e-3), accum_grad_iters=2)
    for b in batches:
        step.step_device(*b)
    torch.cuda.synchronize()
    _assert_params_close(m1, m2, rtol=1e-6)


def test_fused_adamw_state_dict_round_trip_resumes_bias_correction():
    """the step counter must survive state_dict() / load_state_dict() (the reference runner restores the optimizer
    from a checkpoint, runners/runner_base.py:613/662), and a torch.optim.AdamW checkpoint (tensor `step`) must load."""
    td, m1, m2 = _models()
    batches = _batches(td, 4)
    opt1 = td.FusedAdamW(m1, lr=1e-3)
    s1 = td.AlignerTrainStep(m1, opt1)
    for b in batches[:2]:
        s1.step_device(*b)
    sd = copy.deepcopy(opt1.state_dict())
    # resume on a fresh model / optimizer pair
    m2.load_state_dict(m1.state_dict())
    opt2 = td.FusedAdamW(m2, lr=1e-3)
    opt2.load_state_dict(sd)
    assert opt2._t == 2
    s2 = td.AlignerTrainStep(m2, opt2)
    for b in batches[2:]:
        s1.step_device(*b)
        s2.step_device(*b)
    torch.cuda.synchronize()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    # a torch.optim.AdamW checkpoint: tensor steps
    from thinkdiff_mlre_b200.train_step import reference_param_groups

    topt = torch.optim.AdamW(reference_param_groups(m1, 0.05), lr=1e-3)
    x, tgt, idx = _packed(td, batches[0])
    m1.mse_loss_packed(x, tgt, idx).backward()
    topt.step()
    opt3 = td.FusedAdamW(m1, lr=1e-3)
    opt3.load_state_dict(topt.state_dict())
    assert opt3._t == 1
