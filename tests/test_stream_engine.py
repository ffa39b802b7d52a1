"""The streaming engine of the drop-in tools (fastx_toolkit_b200/csrc/host/fxh_stream.c) without a GPU: reader thread,
record-boundary splitting with carry-over, out-of-order workers, in-order writer and the hand-over to the record path, linked
with a test double of the GPU text path that copies whole records through (tests/native/fxs_harness.c).  For every input it
must behave exactly like the reference's fastx_trimmer with default arguments (an identity through libfastx)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H

REF = H.ref_tool("fastx_trimmer")
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("stream")
    exe = str(d / "fastx_trimmer")              # errx() prefixes messages with the program name
    host = os.path.join(H.ROOT, "fastx_toolkit_b200", "csrc", "host")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-I", os.path.join(H.ROOT, "include"), "-I", host, "-o", exe,
                           os.path.join(H.ROOT, "tests", "native", "fxs_harness.c"), os.path.join(host, "fxh.c"),
                           os.path.join(host, "fxh_stream.c"), "-lpthread"])
    return exe


def same(harness, args, stdin=None, env=None):
    e = dict(os.environ, FXS_HARNESS_REPORT="1")
    e.update(env or {})
    m = subprocess.run([harness] + args, input=stdin, capture_output=True, env=e, timeout=300)
    r = subprocess.run([REF] + args, input=stdin, capture_output=True)
    strip = lambda s, exe: s.replace((os.path.dirname(exe) + "/").encode(), b"")
    rep = [l for l in m.stderr.split(b"\n") if l.startswith(b"[harness]")]
    err = b"\n".join(l for l in m.stderr.split(b"\n") if not l.startswith(b"[harness]"))
    assert m.returncode == r.returncode, (args, env, m.stderr[-300:], r.stderr[-300:])
    assert m.stdout == r.stdout, (args, env, "stdout differs", len(m.stdout), len(r.stdout))
    assert strip(err, harness) == strip(r.stderr, REF), (args, env, err[-300:], r.stderr[-300:])
    return rep[0].decode() if rep else ""


ENVS = [dict(), dict(FASTX_CHUNK_BYTES="16384", FASTX_WORKERS="3"), dict(FASTX_CHUNK_BYTES="40000", FASTX_WORKERS="1"),
        dict(FASTX_CHUNK_BYTES="20000", FASTX_WORKERS="4", FASTX_WINDOW_BYTES="70000", FASTX_READ_THREADS="1")]


def test_valid_inputs_go_through_the_engine_in_order(harness, tmp_path):
    seq, qual = H.synth_slab(H.SEED_BASE + 23, 30000, 75, H.WITH_N)
    rng = np.random.default_rng(6)
    lens = H.ragged(seq, qual, rng, min_len=1)
    qual[qual == ord("@")] = ord("A")
    qual[::7, 0] = ord("@")                                   # quality lines that start with '@': the split rule must not be fooled
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, 75)
    fa = str(tmp_path / "in.fa")
    H.write_fasta(fa, seq, lens, 75, prefix="s")
    text = open(fq, "rb").read()
    for env in ENVS:
        for p in (fq, fa):
            rep = same(harness, ["-v", "-i", p], env=env)
            assert "fallback=0" in rep and "records=30000" in rep, rep
        rep = same(harness, ["-v"], stdin=text, env=env)      # a pipe: no pread, no seek
        assert "fallback=0" in rep and "records=30000" in rep, rep
    assert "chunks=1 " in same(harness, ["-i", fq])
    assert int(same(harness, ["-i", fq], env=ENVS[1]).split("chunks=")[1].split()[0]) > 100
    # last line without a newline: the engine takes all records but the last, the record path the last one
    nonl = str(tmp_path / "nonl.fq")
    open(nonl, "wb").write(text[:-1])
    for env in ENVS:
        rep = same(harness, ["-v", "-i", nonl], env=env)
        assert "records=29999" in rep and "fallback=1" in rep, rep
    # output file and gzip
    out1, out2 = str(tmp_path / "o1.fq"), str(tmp_path / "o2.fq")
    m = subprocess.run([harness, "-v", "-i", fq, "-o", out1], capture_output=True, env=dict(os.environ, **ENVS[1]))
    r = subprocess.run([REF, "-v", "-i", fq, "-o", out2], capture_output=True)
    assert (m.returncode, m.stdout) == (r.returncode, r.stdout) and open(out1, "rb").read() == open(out2, "rb").read()
    # one 5 MB block into a regular file: stored through the shared mapping by several threads (fxh.c store_parallel), after
    # some bytes written the ordinary way so that the mapping starts in the middle of a page
    big = str(tmp_path / "big.fq")
    open(big, "wb").write(text * 3)
    for wt in ("4", "1"):
        m = subprocess.run([harness, "-i", big, "-o", out1], capture_output=True, env=dict(os.environ, FASTX_WRITE_THREADS=wt, FASTX_CHUNK_BYTES=str(6 << 20)))
        r = subprocess.run([REF, "-i", big, "-o", out2], capture_output=True)
        assert m.returncode == r.returncode == 0 and open(out1, "rb").read() == open(out2, "rb").read()


def test_broken_inputs_hand_over_to_the_record_path(harness, tmp_path):
    seq, qual = H.synth_slab(H.SEED_BASE + 24, 8000, 50, H.PLAIN)
    base = str(tmp_path / "good.fq")
    H.write_fastq(base, seq, qual, None, 50)
    lines = open(base, "rb").read().split(b"\n")[:-1]

    def variant(name, edit, tail=b"\n"):
        ls = list(lines)
        edit(ls)
        p = str(tmp_path / name)
        open(p, "wb").write(b"\n".join(ls) + tail)
        return p

    def set_line(i, v):
        return lambda ls: ls.__setitem__(i, v)

    cases = [
        variant("noat.fq", set_line(4 * 5000, b"r5000")),
        variant("noat_first.fq", set_line(4 * 1, b"r1")),
        variant("emptyseq.fq", lambda ls: (ls.__setitem__(4 * 100 + 1, b""), ls.__setitem__(4 * 100 + 3, b""))),
        variant("trunc2.fq", lambda ls: ls.__delitem__(slice(4 * 7500 + 1, None))),
        variant("trunc3.fq", lambda ls: ls.__delitem__(slice(4 * 7500 + 2, None))),
        variant("trunc4.fq", lambda ls: ls.__delitem__(slice(4 * 7500 + 3, None))),
        variant("blankend.fq", lambda ls: ls.append(b"")),
        variant("blankmid.fq", lambda ls: ls.insert(4 * 3000, b"")),
        variant("qualshort.fq", set_line(4 * 3000 + 3, b"I" * 30)),
        variant("numeric_one.fq", set_line(4 * 10 + 3, b" ".join([b"40"] * 50))),          # one numeric record: valid, record path
        variant("numeric_bad.fq", set_line(4 * 10 + 3, b"40 40 x 40")),
        variant("cr_in_name.fq", set_line(4 * 4000, b"@r40\r00 tail")),
    ]
    for env in ENVS:
        for p in cases:
            same(harness, ["-i", p], env=env)
    # a record path that takes over must continue to the end of a valid file: a numeric record in the middle, then 5000 more
    rep = same(harness, ["-v", "-i", cases[9]], env=ENVS[1])
    assert "fallback=1" in rep


def test_stateful_ops_never_count_a_chunk_twice(harness, tmp_path):
    """quality stats and the collapser change state on the GPU: when a chunk is handed back to the record path (here: one record with
    numeric qualities in the middle of an ASCII file — valid input), no later chunk may already have been counted"""
    seq, qual = H.synth_slab(H.SEED_BASE + 25, 20000, 50, H.PLAIN)
    fq = str(tmp_path / "mixed.fq")
    H.write_fastq(fq, seq, qual, None, 50)
    lines = open(fq, "rb").read().split(b"\n")[:-1]
    lines[4 * 7000 + 3] = b" ".join([b"40"] * 50)
    open(fq, "wb").write(b"\n".join(lines) + b"\n")
    for env in ENVS:
        e = dict(os.environ, FXS_HARNESS_STATEFUL="1", FXS_HARNESS_REPORT="1")
        e.update(env)
        for _ in range(3):
            m = subprocess.run([harness, "-i", fq], capture_output=True, env=e, timeout=300)
            assert m.returncode == 0, m.stderr[-300:]
            assert b"[harness] counted=20000\n" in m.stderr, (env, m.stderr[-200:])
