import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_meta():
    import json

    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def eng():
    """One engine (context on cuda:0) per test module."""
    from options_model_b200 import engine as E

    e = E.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def mods():
    from options_model_b200 import _lib as L
    from options_model_b200 import engine as E
    from oracle import lsm_oracle as orc

    return L, E, orc
