"""CPU check of the SWAR byte primitives of the hot-path kernels (fastx_toolkit_b200/csrc/fxg_device.cuh: base validation and
complement by PRMT table, quality range and threshold tests, head masks): the header is compiled for the host
(tests/native/swar_host.cpp emulates PRMT) and compared exhaustively over byte values, lanes and every -Q with the reference's
per-character rules (fastx.c:45-84, 118-135; fastx_reverse_complement.c:43-72)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="needs g++ and the CUDA headers")
def test_swar_primitives_on_host(tmp_path):
    exe = str(tmp_path / "swar_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-Wno-attributes", "-I", CUDA_INC,
                           "-I", os.path.join(ROOT, "fastx_toolkit_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tests", "native", "swar_host.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "swar primitives ok" in r.stdout, r.stdout + r.stderr
