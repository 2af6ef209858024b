"""Loader for the committed golden vectors (tests/golden/*.npz). TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_golden(name: str) -> dict:
    return dict(np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False))
