import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pmaf_b200  # noqa: E402,F401  (registers the package under its import alias)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    """Build the C oracle (and the reference build when /root/reference is present)."""
    from oracle import cpu_planners

    cpu_planners.build("all")
    return cpu_planners
