"""CPU check of the packed-s16x2 clipper DP (fastx_toolkit_b200/csrc/fxg_clip_dpx.cuh): the header compiles for the host with
portable emulations of the DPX instructions, and its alignments must equal the oracle's literal restatement of the reference
aligner (oracle/fastx_oracle.c fxo_align) field by field."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("dpx") / "clip_dpx_host")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libfastx_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "fastx_toolkit_b200", "csrc"),
                           "-I", os.path.join(ROOT, "oracle"), "-o", exe, os.path.join(ROOT, "tests", "native", "clip_dpx_host.cpp"),
                           "-L", os.path.join(ROOT, "oracle"), "-lfastx_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


@pytest.mark.parametrize("seed,pairs,L,adapter", [
    (1, 20000, 150, "AGATCGGAAGAGC"), (2, 20000, 36, "CCTTAAGG"), (3, 20000, 50, "ACGT"), (4, 5000, 1, "AGATCGGAAGAGC"),
    (5, 20000, 7, "AGATCGGAAGAGCACA"), (6, 8000, 255, "TGGAATTCTCGG"), (7, 20000, 3, "AC"), (8, 20000, 150, "A"),
    (9, 8000, 256, "AGATCGGAAGAGCACA"), (10, 20000, 100, "TTTTTTTTTTTT"), (11, 20000, 60, "ACACACACACACA"),
    # lengths around the peeled "left ban" columns (HMAX - 4) and the 4-base word loop
    (12, 20000, 13, "AGATCGGAAGAGC"), (13, 20000, 12, "AGATCGGAAGAGCACA"), (14, 20000, 14, "AGATCGGAAGAGC"), (15, 20000, 17, "AGATCGGAAGAGCACA"),
    (16, 20000, 151, "AGATCGGAAGAGC"), (17, 20000, 149, "TGGAATTCTCGGGTGC"), (18, 20000, 19, "GGGGGGGG"), (19, 10000, 200, "ATATATATATAT"),
])
def test_dpx_alignment_equals_reference_aligner(harness, seed, pairs, L, adapter):
    r = subprocess.run([harness, str(seed), str(pairs), str(L), adapter], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
