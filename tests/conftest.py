import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session", autouse=True)
def _cuda_library_built():
    """The in-tree libwl_b200.so normally comes from __graft_entry__.build(); build it here if a fresh checkout has none
    (nvcc cross-compiles for sm_100a without a GPU; a no-op when the library is newer than its sources)."""
    import wl_b200

    try:
        wl_b200.build_library()
    except Exception as e:  # no nvcc: the tests that need the library will say so themselves
        print("libwl_b200.so not (re)built:", e)
    yield
