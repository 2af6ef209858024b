import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:  # tests/_mp.py (the multi-rank process harness) is imported by name, also in spawned ranks
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `pytest -m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch

    # no test may hang the box: a stuck stream wait / collective sits in a C call that no signal interrupts, so the limit is
    # enforced from a watchdog thread (pytest-timeout's "thread" method: dump the stacks, end the process)
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(900, method="thread"))

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


