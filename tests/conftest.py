import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # a fresh checkout has no built library: compile it once (nvcc cross-compiles without a GPU)
    try:
        from petibm_b200 import build as _b

        if _b.needs_build():
            _b.build()
    except Exception as exc:  # the tests that need the library will say so themselves
        print(f"[conftest] libb200ls.so could not be built: {exc}")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
