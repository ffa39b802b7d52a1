"""What the built libfxg.so contains, read from its SASS (cuobjdump, no GPU needed): the code is sm_100a only, the streaming
kernels move their tiles with TMA bulk copies (UBLKCP) completed on mbarriers (SYNCS) and keep everything in registers (no
local-memory loads / stores), K-STATS counts with shared-memory atomics (ATOMS), the clipper's integer DP uses the packed
min/max instructions (VIMNMX / VIADDMNMX).  Mnemonics as listed in /opt/skills/guides/B200_PROFILING.md."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastx_toolkit_b200", "libfxg.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("needs cuobjdump and a built libfxg.so")
    text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
    funcs, cur, archs = collections.OrderedDict(), None, set()
    for line in text.split("\n"):
        m = re.search(r"arch = (\S+)", line)
        if m:
            archs.add(m.group(1))
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return funcs, archs


def pick(funcs, pat):
    sel = {f: c for f, c in funcs.items() if re.search(pat, f)}
    assert sel, pat
    return sel


def test_only_sm_100a_code(sass):
    assert sass[1] == {"sm_100a"}


def test_streaming_kernels_use_tma_and_no_local_memory(sass):
    for pat in (r"k_scan_wI", r"k_revcomp_wI"):
        for f, c in pick(sass[0], pat).items():
            assert c["UBLKCP"] > 0 and c["SYNCS"] > 0, (f, "no TMA bulk copy / mbarrier")
            assert c["STL"] == 0 and c["LDL"] == 0, (f, "local memory traffic (spill)")
            assert c["PRMT"] > 0, f


def test_stats_kernel_counts_with_shared_atomics_behind_tma(sass):
    for f, c in pick(sass[0], r"k_stats4I").items():
        assert c["UBLKCP"] > 0 and c["SYNCS"] > 0 and c["ATOMS"] > 0 and c["PRMT"] > 0, f


def test_clipper_dp_uses_packed_min_max(sass):
    for f, c in pick(sass[0], r"k_clip_dpxI").items():
        assert c["VIMNMX"] + c["VIADDMNMX"] > 0, f
