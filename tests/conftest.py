import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ora():
    """the CPU oracle (test infrastructure, parity unpinned)"""
    from oracle import oracle
    oracle.build()
    return oracle


def jittered_ref_element(elem, seed=0, amp=0.15, scale=1.0, shift=0.0):
    """reference element corners + random displacement (still valid, positively oriented)"""
    ref = {
        "tri": [[0, 0], [1, 0], [0, 1]],
        "quad": [[0, 0], [1, 0], [1, 1], [0, 1]],
        "tet": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
        "hex": [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
        "prism": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]],
    }[elem]
    x = np.array(ref, dtype=float)
    rng = np.random.default_rng(seed)
    return scale * (x + amp * rng.uniform(-1, 1, x.shape)) + shift
