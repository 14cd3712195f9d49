import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gpt_sovits_b200  # noqa: E402,F401  (alias loader for the gpt-sovits_b200/ directory)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def weights_seed0():
    from gpt_sovits_b200 import synthetic
    return synthetic.make_state_dict(seed=0)


@pytest.fixture(scope="session")
def pe_table():
    from gpt_sovits_b200 import synthetic
    return synthetic.sine_pe()
