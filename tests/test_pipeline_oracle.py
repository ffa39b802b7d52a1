"""Groundwork for SURVEY §8(f-3), fused pipelines: what a fused GPU pipeline will have to reproduce is the output of the
reference tools in a shell pipe.  This file pins that target on the CPU: a composition of the oracle's per-tool functions
(with survivors compacted between stages, and the clipper's stale-buffer semantics — SURVEY Appendix D.1 — when it runs on
mixed-length reads) must equal `tool | tool | ...` run with the reference binaries, and the digests of those outputs are
committed (tests/golden/pipeline.json, made by tests/golden/make_pipeline_golden.py) for boxes without /root/reference."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H

ADAPTER = b"AGATCGGAAGAGC"
GOLD = os.path.join(H.ROOT, "tests", "golden", "pipeline.json")
needs_ref = pytest.mark.skipif(H.ref_tool("fastx_clipper") is None, reason="oracle/_ref not built")


def compact(seq, qual, lens):
    keep = np.flatnonzero(lens >= 0)
    return np.ascontiguousarray(seq[keep]), np.ascontiguousarray(qual[keep]), lens[keep].astype(np.int32)


def stage_clip(seq, qual, lens, min_length, discard_unknown=1):
    """fastx_clipper on whatever lengths arrive: rows as the reference's aligner sees them (NUL + stale bytes of earlier,
    longer reads; matrix width = running maximum), as the host packer builds them (fxh.c stale rows)"""
    n, stride = seq.shape
    rows = np.zeros((n, stride), np.uint8)
    widths = np.zeros(n, np.int32)
    shadow = np.zeros(stride + 1, np.uint8)
    wmax = 0
    for i in range(n):
        l = int(lens[i])
        shadow[:l] = seq[i, :l]
        shadow[l] = 0
        wmax = max(wmax, l)
        rows[i, :wmax] = shadow[:wmax]
        widths[i] = wmax
    opts = H.FxoClipOpts(min_length=min_length, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=discard_unknown,
                         min_adapter_len=0)
    out_len, cls, _ = H.o_clip(rows, lens, widths, 0, stride, ADAPTER, opts)
    return np.where(cls == 0, out_len, -1).astype(np.int32)


def stage_trim(seq, qual, lens, t, l):
    out, bad = H.o_trim(seq, qual, lens, 0, seq.shape[1], 33, t, l)
    assert bad == -1
    return out


def stage_filter(seq, qual, lens, q, p):
    keep, bad = H.o_filter(seq, qual, lens, 0, seq.shape[1], 33, q, p)
    assert bad == -1
    return np.where(keep != 0, lens, -1).astype(np.int32)


def stage_collapse(seq, lens):
    first, cnt = H.o_collapse(seq, lens, 0, seq.shape[1])
    return b"".join(b">%d-%d\n" % (k + 1, int(c)) + seq[int(f), :int(lens[int(f)])].tobytes() + b"\n" for k, (f, c) in enumerate(zip(first, cnt)))


PIPELINES = {
    # name: (reference command line stages, oracle stages)
    "clip_trim_filter_collapse": (
        [["fastx_clipper", "-Q33", "-a", ADAPTER.decode(), "-l", "20"], ["fastq_quality_trimmer", "-Q33", "-t", "20", "-l", "20"],
         ["fastq_quality_filter", "-Q33", "-q", "20", "-p", "90"], ["fastx_collapser", "-Q33"]],
        [lambda s, q, l: stage_clip(s, q, l, 20), lambda s, q, l: stage_trim(s, q, l, 20, 20), lambda s, q, l: stage_filter(s, q, l, 20, 90)]),
    "trim_clip_collapse": (      # the clipper after the trimmer: mixed lengths, stale-buffer semantics
        [["fastq_quality_trimmer", "-Q33", "-t", "25", "-l", "30"], ["fastx_clipper", "-Q33", "-a", ADAPTER.decode(), "-l", "15", "-n"],
         ["fastx_collapser", "-Q33"]],
        [lambda s, q, l: stage_trim(s, q, l, 25, 30), lambda s, q, l: stage_clip(s, q, l, 15, discard_unknown=0)]),
}


def synth_input(n=6000, L=100):
    seq, qual = H.synth_slab(H.SEED_BASE + 30, n, L, H.ADAPTER)
    seq[::50, 7] = ord("N")
    seq[1::9] = seq[0::9][: len(seq[1::9])]                       # duplicates for the collapser
    qual[1::9] = qual[0::9][: len(qual[1::9])]
    return seq, qual, L


def oracle_pipeline(name):
    seq, qual, L = synth_input()
    lens = np.full(seq.shape[0], L, np.int32)
    for st in PIPELINES[name][1]:
        lens = st(seq, qual, lens)
        seq, qual, lens = compact(seq, qual, lens)
    return stage_collapse(seq, lens)


@pytest.mark.parametrize("name", sorted(PIPELINES))
def test_composed_oracle_matches_committed_digest(name):
    out = oracle_pipeline(name)
    gold = json.load(open(GOLD))[name]
    assert out.count(b"\n") // 2 == gold["unique_sequences"]
    assert hashlib.sha256(out).hexdigest() == gold["sha256"]


@needs_ref
@pytest.mark.parametrize("name", sorted(PIPELINES))
def test_composed_oracle_matches_reference_pipe(name, tmp_path):
    seq, qual, L = synth_input()
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, None, L)
    data = open(fq, "rb").read()
    for cmd in PIPELINES[name][0]:
        r = subprocess.run([H.ref_tool(cmd[0])] + cmd[1:], input=data, capture_output=True)
        assert r.returncode == 0, r.stderr
        data = r.stdout
    assert oracle_pipeline(name) == data
