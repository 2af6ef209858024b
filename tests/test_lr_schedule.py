"""LR schedules (SURVEY section 8 f-3 host logic): thinkdiff_mlre_b200.lr_schedule against (a) the committed golden values produced
by the reference's own scheduler classes (oracle/make_golden.py, thinkdiff/common/optims.py) and (b), in the dev container, the
reference classes exec'd live. Bit-exact: both sides evaluate the same float64 expressions."""
import types

import numpy as np
import pytest

from oracle import ref_loader
from oracle.golden import load_golden
from oracle.make_golden import LR_CASES, lr_points
from thinkdiff_mlre_b200.lr_schedule import SCHEDULERS, get_lr_scheduler_class


def _opt():
    return types.SimpleNamespace(param_groups=[{"lr": -1.0, "weight_decay": 0.05}, {"lr": -1.0, "weight_decay": 0.0}])


@pytest.mark.parametrize("case", sorted(LR_CASES))
def test_matches_golden_values_from_the_reference_classes(case):
    g = load_golden("lr_schedule.npz")
    sched, kw = LR_CASES[case]
    opt = _opt()
    obj = get_lr_scheduler_class(sched)(optimizer=opt, **kw)
    pts = g[case + ".points"]
    assert pts.tolist() == [list(p) for p in lr_points(kw)]
    got = []
    for e, s in pts.tolist():
        lr = obj.step(cur_epoch=e, cur_step=s)
        assert opt.param_groups[0]["lr"] == opt.param_groups[1]["lr"] == lr == obj.lr_at(e, s)
        got.append(lr)
    np.testing.assert_array_equal(np.asarray(got), g[case + ".lr"])


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (dev container only)")
def test_every_iteration_of_a_short_run_matches_the_live_reference():
    ref = ref_loader.load_lr_schedulers()
    assert sorted(ref) == sorted(SCHEDULERS)
    for sched, kw in LR_CASES.values():
        if kw["max_epoch"] * kw["iters_per_epoch"] > 1000:
            kw = dict(kw, max_epoch=2, iters_per_epoch=300, warmup_steps=450)  # warm-up crosses the epoch boundary
        a, b = _opt(), _opt()
        mine, theirs = SCHEDULERS[sched](optimizer=a, **kw), ref[sched](optimizer=b, **kw)
        for e in range(kw["max_epoch"]):
            for s in range(kw["iters_per_epoch"]):
                mine.step(e, s), theirs.step(cur_epoch=e, cur_step=s)
                assert a.param_groups[0]["lr"] == b.param_groups[0]["lr"], (sched, e, s)


def test_unknown_name_is_an_error():
    with pytest.raises(KeyError):
        get_lr_scheduler_class("constant")
