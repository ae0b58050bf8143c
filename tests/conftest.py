import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the host pipeline calls into libpileup_b200.so (window layout is native host code): build it when missing
    from coolpuppy_b200.build import build_native

    build_native(force=False)


@pytest.fixture(scope="session")
def fixtures_dir():
    return os.path.join(ROOT, "tests", "fixtures")
