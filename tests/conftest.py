import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def nb():
    import nabla_jl_b200
    return nabla_jl_b200


@pytest.fixture(scope="session")
def orc():
    import nabla_oracle
    return nabla_oracle


@pytest.fixture(scope="session")
def ctx(nb):
    """Device context; creating it raises if there is no GPU (no CPU fallback)."""
    return nb.get_context()
