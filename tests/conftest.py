import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: full-size configuration (tens of seconds)")


@pytest.fixture(scope="session")
def cs():
    """The product package (compressedsensing.jl_b200/) loaded by path; needs libcsb200.so built."""
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libcsb200.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def po():
    from oracle import pursuit_oracle
    return pursuit_oracle
