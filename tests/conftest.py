import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle's C restatement is test infrastructure: build it on demand (gcc only)."""
    if not all(os.path.exists(os.path.join(ROOT, "oracle", f)) for f in ("libsmf_cpu.so", "libsmf_cpuref_w4.so", "libsmf_cpuref_w8.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
