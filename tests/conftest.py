import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Build what the tests load (libfxg.so, bin/ tools, the oracle) when a fresh checkout has not been built yet."""
    import shutil
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "libfastx_oracle.so")):
        subprocess.call(["make", "-C", ROOT, "oracle"], stdout=subprocess.DEVNULL)          # gcc only
    need = [os.path.join(ROOT, "fastx_toolkit_b200", "libfxg.so"), os.path.join(ROOT, "bin", "fastx_b200")]
    if not all(os.path.exists(p) for p in need):
        # without nvcc the CUDA library cannot be built: the tests that need it fail on their own, the CPU-only ones still run
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            try:
                subprocess.check_call(["make", "-C", ROOT, "lib", "tools"], stdout=subprocess.DEVNULL)
            except subprocess.CalledProcessError as e:
                print("conftest: building libfxg.so / bin failed (%s); GPU-library tests will fail" % e, file=sys.stderr)
