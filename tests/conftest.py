import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand with gcc."""
    from oracle import oracle_py

    oracle_py.load()
    return oracle_py


@pytest.fixture(scope="session")
def nmi_lib():
    """The product library. Must already be built (python -m orbslam2_nmi_b200.build)."""
    from orbslam2_nmi_b200 import build, capi

    build.build_cuda()
    return capi.load()
