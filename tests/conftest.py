import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(target_dir, *targets):
    # one build at a time per directory: pytest-xdist workers would otherwise load a half-written library
    import fcntl

    with open(os.path.join(target_dir, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        subprocess.run(["make", "-C", target_dir, *targets], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle_api():
    """The C oracle (test infrastructure).  Built on demand; gcc exists on both boxes."""
    from uxsim_b200._capi import CApi

    path = os.path.join(ROOT, "oracle", "liboracle.so")
    _make(os.path.join(ROOT, "oracle"), "all")
    return CApi(path, "uxo_")


@pytest.fixture(scope="session")
def ref_api():
    """The reference's own C++ engine behind the harness (oracle/_ref).  Prebuilt in the dev container; skipped
    when absent (it cannot be rebuilt where /root/reference does not exist)."""
    from uxsim_b200._capi import CApi

    path = os.path.join(ROOT, "oracle", "_ref", "libuxsim_ref.so")
    if os.path.isdir("/root/reference/uxsim/trafficpp"):
        _make(os.path.join(ROOT, "oracle"), "ref")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libuxsim_ref.so not built (reference sources absent)")
    return CApi(path, "uxr_")


@pytest.fixture(scope="session")
def emu_api():
    """Host emulation build of the kernel logic (tests/emu); test-only, never the product."""
    from uxsim_b200._capi import CApi

    _make(os.path.join(ROOT, "tests", "emu"))
    return CApi(os.path.join(ROOT, "tests", "emu", "_build", "libuxsim_emu.so"), "ux_")


@pytest.fixture(scope="session")
def cuda_api():
    """The product.  Fails (not skips) when the CUDA library is missing: there is no fallback to test."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from uxsim_b200._capi import load_product

    return load_product()
