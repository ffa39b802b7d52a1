"""Host layer of the drop-in tools without a GPU: fxh.c's block reader, batch packer and block writer are linked with stubs
for libfxg.so (tests/native/fxh_passthrough.c) into a pass-through program; for every input it must behave exactly like the
reference's fastx_trimmer with default arguments (an identity through libfastx's reader and writer): same output bytes, same
-v report, and for structurally broken input the same output prefix, message and exit status."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H

REF = H.ref_tool("fastx_trimmer")
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostlayer")
    exe = str(d / "fastx_trimmer")              # errx() prefixes messages with the program name
    host = os.path.join(H.ROOT, "fastx_toolkit_b200", "csrc", "host")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-I", os.path.join(H.ROOT, "include"), "-I", host, "-o", exe,
                           os.path.join(H.ROOT, "tests", "native", "fxh_passthrough.c"), os.path.join(host, "fxh.c"), "-lpthread"])
    return exe


def same(harness, args, stdin=None, env=None):
    e = dict(os.environ)
    e.update(env or {})
    m = subprocess.run([harness] + args, input=stdin, capture_output=True, env=e)
    r = subprocess.run([REF] + args, input=stdin, capture_output=True)
    strip = lambda s, exe: s.replace((os.path.dirname(exe) + "/").encode(), b"")
    assert m.returncode == r.returncode, (args, m.stderr[-300:], r.stderr[-300:])
    if "-z" in args:      # the reference pipes through a gzip child; this build frames the stream itself: same payload, other bytes
        import gzip
        assert gzip.decompress(m.stdout) == gzip.decompress(r.stdout), (args, "gunzip payload differs")
        return r
    assert m.stdout == r.stdout, (args, "stdout differs")
    assert strip(m.stderr, harness) == strip(r.stderr, REF), (args, m.stderr[-300:], r.stderr[-300:])
    return r


def test_valid_inputs_round_trip_like_the_reference(harness, tmp_path):
    seq, qual = H.synth_slab(H.SEED_BASE + 21, 20000, 75, H.WITH_N)
    rng = np.random.default_rng(4)
    lens = H.ragged(seq, qual, rng, min_len=1)
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, 75)
    fa = str(tmp_path / "in.fa")
    H.write_fasta(fa, seq, lens, 75, prefix="12-")            # collapsed identifiers: "N-COUNT" read counts (fastx.c:475-497)
    text = open(fq, "rb").read()
    crlf = str(tmp_path / "crlf.fq")
    open(crlf, "wb").write(text.replace(b"\n", b"\r\n"))
    nonl = str(tmp_path / "nonl.fq")
    open(nonl, "wb").write(text[:-1])                         # last line without a newline
    for env in ({}, {"FASTX_BATCH_READS": "777"}):
        for p in (fq, fa, crlf, nonl, os.path.join(H.GOLDEN, "fastx_rev_comp2.fastq"), os.path.join(H.GOLDEN, "fasta_collapser1.fasta")):
            same(harness, ["-v", "-i", p], env=env)
    same(harness, ["-Q", "64", "-v", "-i", os.path.join(H.GOLDEN, "fastq_quality_trimmer.fastq")])
    same(harness, ["-v"], stdin=text)
    same(harness, ["-z", "-i", fq])
    out1, out2 = str(tmp_path / "o1.fq"), str(tmp_path / "o2.fq")
    m = subprocess.run([harness, "-v", "-i", fq, "-o", out1], capture_output=True)
    r = subprocess.run([REF, "-v", "-i", fq, "-o", out2], capture_output=True)
    assert (m.returncode, m.stdout, m.stderr) == (r.returncode, r.stdout, r.stderr)
    assert open(out1, "rb").read() == open(out2, "rb").read()


def test_structurally_broken_inputs_fail_like_the_reference(harness, tmp_path):
    seq, qual = H.synth_slab(H.SEED_BASE + 22, 6000, 50, H.PLAIN)
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
        variant("emptyseq.fq", lambda ls: (ls.__setitem__(4 * 100 + 1, b""), ls.__setitem__(4 * 100 + 3, b""))),
        variant("trunc2.fq", lambda ls: ls.__delitem__(slice(4 * 5500 + 1, None))),
        variant("trunc3.fq", lambda ls: ls.__delitem__(slice(4 * 5500 + 2, None))),
        variant("trunc4.fq", lambda ls: ls.__delitem__(slice(4 * 5500 + 3, None))),
        variant("blankend.fq", lambda ls: ls.append(b"")),
        variant("qualshort.fq", set_line(4 * 3000 + 3, b"I" * 30)),
        variant("quallong.fq", set_line(4 * 3000 + 3, b"I" * 60)),
        variant("numeric_bad.fq", set_line(4 * 10 + 3, b"40 40 x 40")),
        variant("numeric_range.fq", set_line(4 * 10 + 3, b" ".join([b"40"] * 49 + [b"120"]))),
        # chomp() ends a line at its FIRST carriage return (chomp.c:34-44)
        variant("cr_in_name.fq", set_line(4 * 4000, b"@r40\r00 tail")),
        variant("cr_in_name2.fq", set_line(4 * 4000 + 2, b"+r40\r00")),
        variant("cr_in_seq.fq", set_line(4 * 2500 + 1, b"ACGTACGT\rCGT" + b"A" * 38)),
        variant("cr_in_qual.fq", set_line(4 * 2500 + 3, b"I" * 20 + b"\r" + b"I" * 29)),
    ]
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").write(b"")
    junk = str(tmp_path / "junk.txt")
    open(junk, "wb").write(b"hello\nworld\n")
    fa_bad = str(tmp_path / "bad.fa")
    open(fa_bad, "wb").write(b">a\nACGT\nACGT\n>b\nAC\n")      # a second sequence line where an identifier is expected
    fa_blank = str(tmp_path / "blank.fa")
    open(fa_blank, "wb").write(b">a\nACGT\n\n>b\nACGT\n")     # a blank line is not a nucleotide string: "expecting FASTA prefix"
    for env in ({}, {"FASTX_BATCH_READS": "1000"}):
        for p in cases + [empty, junk, fa_bad, fa_blank, str(tmp_path / "missing.fq")]:
            same(harness, ["-i", p], env=env)
