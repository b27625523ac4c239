import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def tables():
    """The two 16-bit dictionaries the golden crops were generated with (re-generated from seeds)."""
    from workloads import synth
    tab16, nrm16, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.0)
    tab16n, nrm16n, _ = synth.make_dict(16, seed=11, radius=51.0, missing_frac=0.2)
    return dict(full=(tab16, nrm16), nan20=(tab16n, nrm16n))
