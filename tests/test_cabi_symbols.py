"""CPU: the C-ABI library loads and exports every symbol include/fxg.h declares; without a GPU the
product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import helpers as H

ROOT = H.ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fxg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fxg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import fastx_toolkit_b200 as F
    L = F.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libfxg.so does not export %s" % s


def test_binding_covers_header():
    import fastx_toolkit_b200._lib as B
    B.lib()
    src = open(B.__file__).read()
    for s in declared_symbols():
        assert '"%s"' % s in src, "python binding lacks %s" % s


def test_no_gpu_means_loud_failure():
    import torch
    import fastx_toolkit_b200 as F
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(F.FxgError):
        F.Context(0)
    assert b"CUDA" in F.lib().fxg_strerror(1)


def test_product_does_not_reference_oracle():
    """The product tree must never load/link the oracle."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "fastx_toolkit_b200")):
        for fn in fns:
            if fn.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                t = open(os.path.join(dp, fn), errors="replace").read()
                if "fastx_oracle" in t or "oracle/" in t:
                    bad.append(fn)
    assert not bad, bad
