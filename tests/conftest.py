import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(HERE, "golden", "golden.json")))


@pytest.fixture(scope="session")
def data_dir():
    return os.path.join(HERE, "golden", "data")


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session (GPU tests only)."""
    import fasim_b200 as fb
    eng = fb.Engine(0)
    yield eng
    eng.close()
