import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libmarl_b200.so (nvcc cross-compiles without a GPU) and the C oracle once per session.
    On the GPU box the prebuilt in-tree files are reused."""
    from dqn_marl_b200 import build as b
    if b.needs_build():
        b.build()
    import oracle
    oracle.build()
    yield
