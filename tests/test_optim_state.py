"""CPU tests of the optimizer / step plumbing that needs no kernel: FusedAdamW's checkpoint contract (the reference runner saves
``optimizer.state_dict()`` and restores it on every rank, thinkdiff/runners/runner_base.py:613/662), the synthetic generator's two
source layouts, and that the CPU arm of bench.py never maps the CUDA library."""
import copy
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _opt_with_state(td, m, t=3):
    opt = td.FusedAdamW(m, lr=1e-3)
    g = torch.Generator().manual_seed(0)
    for p in m.parameters():
        opt.state[p] = {"exp_avg": torch.randn(p.shape, generator=g), "exp_avg_sq": torch.rand(p.shape, generator=g), "step": t}
    opt._t = t
    return opt


def test_state_dict_round_trips_the_step_counter_and_moments():
    import thinkdiff_mlre_b200 as td

    m = td.ThinkDiffAligner(64, 128)
    opt = _opt_with_state(td, m, t=7)
    sd = opt.state_dict()
    assert sd["fused_adamw_step"] == 7
    assert all(torch.is_tensor(v["step"]) and float(v["step"]) == 7.0 for v in sd["state"].values())  # torch.optim.AdamW's format
    m2 = td.ThinkDiffAligner(64, 128)
    opt2 = td.FusedAdamW(m2)
    opt2.load_state_dict(copy.deepcopy(sd))
    assert opt2._t == 7
    for p, q in zip(m.parameters(), m2.parameters()):
        assert torch.equal(opt.state[p]["exp_avg"], opt2.state[q]["exp_avg"])
        assert torch.equal(opt.state[p]["exp_avg_sq"], opt2.state[q]["exp_avg_sq"])
        assert opt2.state[q]["step"] == 7


def test_loads_a_torch_adamw_checkpoint_and_torch_loads_ours():
    import thinkdiff_mlre_b200 as td
    from thinkdiff_mlre_b200.train_step import reference_param_groups

    m = td.ThinkDiffAligner(64, 128)
    topt = torch.optim.AdamW(reference_param_groups(m, 0.05), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    topt.step()
    topt.step()
    opt = td.FusedAdamW(m)
    opt.load_state_dict(topt.state_dict())
    assert opt._t == 2  # bias corrections continue at step 3, not at step 1
    ours = _opt_with_state(td, m, t=4).state_dict()
    ours.pop("fused_adamw_step")
    torch.optim.AdamW(reference_param_groups(m, 0.05), lr=1e-3).load_state_dict(ours)


def test_truncated_synthetic_layout_is_the_kept_rows_of_the_full_layout():
    from thinkdiff_mlre_b200.synth import lvlm_batch_tensors

    flat, start, lens, tgt = lvlm_batch_tensors(7, 30, 32, 64, seed=5)
    flat_t, start_t, lens_t, tgt_t = lvlm_batch_tensors(7, 30, 32, 64, seed=5, truncated=True)
    assert torch.equal(lens, lens_t)
    assert start_t.tolist() == [0] + torch.cumsum(lens.to(torch.int64), 0).tolist()[:-1]
    keep = torch.cat([flat[int(s) : int(s) + int(n)] for s, n in zip(start.tolist(), lens.tolist())])
    assert torch.equal(keep, flat_t)
    assert torch.equal(torch.cat([tgt[int(s) : int(s) + int(n)] for s, n in zip(start.tolist(), lens.tolist())]), tgt_t)


def test_reference_arm_of_bench_never_maps_the_cuda_library():
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "print('SO_LOADED', any('libthinkdiff' in l for l in open('/proc/self/maps')), file=sys.stderr)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert "SO_LOADED False" in r.stderr, r.stderr[-2000:]
    assert '"impl": "reference"' in r.stdout
