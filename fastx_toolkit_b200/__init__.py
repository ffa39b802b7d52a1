"""fastx_toolkit_b200 — B200-native per-read transform loop of agordon/fastx_toolkit.

The product is the C-ABI shared library ``libfxg.so`` (hand-written sm_100a CUDA, built in-tree by
``make lib``) and the drop-in host tools in ``bin/`` (C).  This Python package is only the thin
ctypes binding used by the tests and ``bench.py``; there is no Python or CPU fallback — if the
library is missing, or no GPU is usable, every call fails loudly.
"""
from ._lib import FxgError, Context, Collapser, Comm, DCollapser, DCollapseReport, TextPipe, TextReport, Batch, BarcodeTable, ClipOpts, Stage, Report, collapse_order_dev, lib, lib_path  # noqa: F401

__all__ = ["FxgError", "Context", "Collapser", "Comm", "DCollapser", "DCollapseReport", "TextPipe", "TextReport", "Batch", "BarcodeTable", "ClipOpts", "Stage", "Report", "collapse_order_dev", "lib", "lib_path"]
