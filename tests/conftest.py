import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

_BUILD_ERROR = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "native: needs libpileup_b200.so (host-side entry points; no GPU)")
    # the host pipeline calls into libpileup_b200.so (window layout is native host code): build it when missing
    global _BUILD_ERROR
    try:
        from coolpuppy_b200.build import build_native

        build_native(force=False)
    except Exception as e:  # no nvcc on this machine: the pure-Python reader / oracle tests still run
        _BUILD_ERROR = str(e).splitlines()[0] if str(e) else repr(e)


def _device_count():
    try:
        from coolpuppy_b200 import _native

        return _native.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a machine without a CUDA device, unless PUP_REQUIRE_GPU=1 -- the GPU box
    sets nothing and has a device, so there every `-m gpu` test runs; tests that need the native library are skipped
    when it could not be built."""
    need_gpu = [it for it in items if "gpu" in it.keywords]
    if need_gpu and os.environ.get("PUP_REQUIRE_GPU", "0") != "1" and _device_count() < 1:
        skip = pytest.mark.skip(reason="no CUDA device visible (GPU tests run on the B200 box: pytest -m gpu)")
        for it in need_gpu:
            it.add_marker(skip)
    if _BUILD_ERROR is not None:
        skip = pytest.mark.skip(reason=f"libpileup_b200.so could not be built: {_BUILD_ERROR}")
        native_files = ("test_abi.py", "test_host_pipeline.py", "test_multigpu_cpu.py", "test_gpu_parity.py",
                        "test_gpu_configs.py")
        for it in items:
            if os.path.basename(str(it.fspath)) in native_files or "native" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def fixtures_dir():
    return os.path.join(ROOT, "tests", "fixtures")
