"""Drop-in executables (bin/, host C + libfxg.so) vs the unmodified reference binaries (oracle/_ref/).
CPU part: everything that is decided before the GPU is touched (usage, flag errors, format sniffing).
GPU part (-m gpu): byte-identical stdout / stderr / exit status on fixtures, synthetic and broken inputs."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H

BIN = os.path.join(H.ROOT, "bin")
TOOLS = ["fastq_quality_trimmer", "fastq_quality_filter", "fastx_reverse_complement", "fastx_clipper",
         "fastx_collapser", "fastx_quality_stats"]

needs_ref = pytest.mark.skipif(H.ref_tool("fastq_quality_trimmer") is None, reason="oracle/_ref not built")
needs_bin = pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "fastx_clipper")), reason="bin/ not built (make tools)")


def run_tool(exe, args, stdin=None):
    r = subprocess.run([exe] + args, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr


def both(tool, args, stdin=None):
    m, r = run_tool(os.path.join(BIN, tool), args, stdin), run_tool(H.ref_tool(tool), args, stdin)
    # getopt prints argv[0] verbatim: strip the two install directories before comparing
    m = (m[0], m[1], m[2].replace((BIN + "/").encode(), b""))
    r = (r[0], r[1], r[2].replace((H.REF_BIN + "/").encode(), b""))
    return m, r


def assert_same(tool, args, stdin=None):
    mine, ref = both(tool, args, stdin)
    assert mine[0] == ref[0], (tool, args, mine[2][-300:], ref[2][-300:])
    if "-z" in args and mine[0] == 0:      # gzip framing differs (GPU DEFLATE blocks vs the reference's gzip child): the payload is the contract
        assert gzip.decompress(mine[1]) == gzip.decompress(ref[1]), (tool, args, "gunzip payload differs")
    else:
        assert mine[1] == ref[1], (tool, args, "stdout differs")
    assert mine[2] == ref[2], (tool, args, mine[2][-300:], ref[2][-300:])
    return ref


# ----------------------------------------------------------------------------------------- CPU part
@needs_ref
@needs_bin
@pytest.mark.parametrize("tool", TOOLS)
def test_usage_and_flag_errors_match_reference(tool, tmp_path):
    assert_same(tool, ["-h"])
    assert_same(tool, ["-Z"])                      # unknown flag: getopt message + "use '-h'..." + exit 1
    assert_same(tool, ["-i", str(tmp_path / "does_not_exist.fq")] + (["-t", "5"] if tool == "fastq_quality_trimmer" else []))
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    assert_same(tool, ["-i", str(empty)] + (["-t", "5"] if tool == "fastq_quality_trimmer" else []))
    junk = tmp_path / "junk.txt"
    junk.write_bytes(b"hello\nworld\n")
    assert_same(tool, ["-i", str(junk)] + (["-t", "5"] if tool == "fastq_quality_trimmer" else []))


@needs_ref
@needs_bin
def test_tool_specific_argument_errors(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a\nACGT\n")
    assert_same("fastq_quality_trimmer", ["-i", str(fa)])                   # missing -t
    assert_same("fastq_quality_trimmer", ["-t", "20", "-i", str(fa)])       # FASTA to a FASTQ-only tool
    assert_same("fastq_quality_filter", ["-p", "0", "-i", str(fa)])
    assert_same("fastq_quality_filter", ["-p", "101", "-i", str(fa)])
    assert_same("fastq_quality_filter", ["-q", "20", "-p", "50", "-i", str(fa)])
    assert_same("fastx_clipper", ["-M", "0", "-i", str(fa)])


# ----------------------------------------------------------------------------------------- GPU part
gpu = pytest.mark.gpu


def synth_fastq(path, n, L, kind, lens_rng=None, crlf=False):
    seq, qual = H.synth_slab(H.SEED_BASE + 11, n, L, kind)
    lens = H.ragged(seq, qual, lens_rng, min_len=6) if lens_rng is not None else None
    H.write_fastq(path, seq, qual, lens, L)
    if crlf:
        data = open(path, "rb").read().replace(b"\n", b"\r\n")
        open(path, "wb").write(data)
    return seq, qual, lens


@gpu
@needs_ref
def test_golden_fixtures_through_the_binaries():
    G = H.GOLDEN
    cases = [
        ("fastq_quality_trimmer", ["-Q", "64", "-t", "30", "-l", "16"], "fastq_quality_trimmer.fastq", "fastq_quality_trimmer.out"),
        ("fastq_quality_filter", ["-Q", "64", "-q", "33", "-p", "100"], "fastq_qual_filter1.fastq", "fastq_qual_filter1a.out"),
        ("fastq_quality_filter", ["-Q", "64", "-q", "20", "-p", "80"], "fastq_qual_filter1.fastq", "fastq_qual_filter1b.out"),
        ("fastx_clipper", ["-Q", "64", "-l", "15", "-a", "CAATTGGTTAATCCCCCTATATA", "-d", "0", "-n", "-c"], "fastx_clipper1.fastq", "fastx_clipper1a.out"),
        ("fastx_reverse_complement", [], "fastx_rev_comp1.fasta", "fastx_reverse_complement1.out"),
        ("fastx_reverse_complement", ["-Q", "64"], "fastx_rev_comp2.fastq", "fastx_reverse_complement2.out"),
        ("fastx_quality_stats", ["-Q", "64"], "fastq_stats1.fastq", "fastq_stats1.out"),
    ]
    for tool, args, fin, fout in cases:
        rc, out, errs = run_tool(os.path.join(BIN, tool), args + ["-i", os.path.join(G, fin)])
        assert rc == 0, errs
        assert out == open(os.path.join(G, fout), "rb").read(), (tool, args)
        assert_same(tool, args + ["-v", "-i", os.path.join(G, fin)])
    # collapser: the fixture pins only the distinct-count ranks; the binary pins the rest
    assert_same("fastx_collapser", ["-v", "-i", os.path.join(G, "fasta_collapser1.fasta")])


@gpu
@needs_ref
@pytest.mark.parametrize("L,kind,ragged,crlf", [(100, H.PLAIN, False, False), (150, H.WITH_N, False, False),
                                                 (75, H.WITH_N, True, False), (50, H.PLAIN, False, True)])
def test_trim_filter_revcomp_stats_binaries(tmp_path, L, kind, ragged, crlf):
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 30000, L, kind, np.random.default_rng(L) if ragged else None, crlf)
    os.environ["FASTX_BATCH_READS"] = "7001"          # several batches, with a ragged last one
    try:
        assert_same("fastq_quality_trimmer", ["-t", "20", "-l", "20", "-v", "-i", fq])
        assert_same("fastq_quality_trimmer", ["-t", "35", "-i", fq])
        assert_same("fastq_quality_filter", ["-q", "20", "-p", "90", "-v", "-i", fq])
        assert_same("fastq_quality_filter", ["-q", "30", "-i", fq])          # -p defaults to 0
        assert_same("fastx_reverse_complement", ["-v", "-i", fq])
        assert_same("fastx_quality_stats", ["-i", fq])
        assert_same("fastx_quality_stats", ["-N", "-i", fq])
        assert_same("fastq_quality_trimmer", ["-t", "20"], stdin=open(fq, "rb").read())   # stdin -> stdout
    finally:
        os.environ.pop("FASTX_BATCH_READS", None)
    # FASTA: a blank line where a '>' is expected is "expecting FASTA prefix", not the multi-line FASTA message
    fa = str(tmp_path / "blank.fa")
    open(fa, "wb").write(b">a\nACGT\n\n>b\nACGT\n")
    assert_same("fastx_reverse_complement", ["-i", fa])
    assert_same("fastx_collapser", ["-i", fa])
    ml = str(tmp_path / "multiline.fa")
    open(ml, "wb").write(b">a\nACGT\nACGT\n>b\nACGT\n")
    assert_same("fastx_reverse_complement", ["-i", ml])


@gpu
@needs_ref
def test_small_windows_and_chunks(tmp_path):
    """tiny reader window + tiny GPU text chunks: records straddle every boundary; output must not change"""
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 40000, 100, H.WITH_N, np.random.default_rng(5))
    fa = str(tmp_path / "in.fa")
    seq, _ = H.synth_slab(H.SEED_BASE + 4, 40000, 60, H.DUPS)
    H.write_fasta(fa, seq, None, 60)
    env_sets = [dict(FASTX_WINDOW_BYTES="70000", FASTX_CHUNK_BYTES="20000"), dict(FASTX_WINDOW_BYTES="300000", FASTX_CHUNK_BYTES="100000"),
                dict(FASTX_WINDOW_BYTES="70000", FASTX_TEXT_PATH="0", FASTX_BATCH_READS="777")]
    for env in env_sets:
        os.environ.update(env)
        try:
            assert_same("fastq_quality_trimmer", ["-t", "25", "-l", "30", "-v", "-i", fq])
            assert_same("fastq_quality_filter", ["-q", "20", "-p", "80", "-v", "-i", fq])
            assert_same("fastx_reverse_complement", ["-i", fq])
            assert_same("fastx_quality_stats", ["-i", fq])
            assert_same("fastx_clipper", ["-a", "AGATCGGAAGAGC", "-l", "10", "-n", "-v", "-i", fq])
            assert_same("fastx_collapser", ["-v", "-i", fa])
            assert_same("fastx_reverse_complement", ["-i", fa])
            assert_same("fastq_quality_trimmer", ["-t", "25"], stdin=open(fq, "rb").read())
        finally:
            for k in env:
                os.environ.pop(k, None)


@gpu
@needs_ref
def test_clipper_text_path_and_fallback_transition(tmp_path):
    """fastx_clipper on the GPU text path (equal-length reads), then a file whose read lengths change part-way: the
    record path that takes over must see the aligner's stale query buffer exactly as the reference leaves it"""
    fq = str(tmp_path / "uni.fq")
    synth_fastq(fq, 30000, 100, H.ADAPTER, np.random.default_rng(11))
    rng = np.random.default_rng(12)
    mixed = str(tmp_path / "mixed.fq")
    with open(fq, "rb") as f:
        lines = f.read().split(b"\n")[:-1]
    recs = [lines[i:i + 4] for i in range(0, len(lines), 4)]
    with open(mixed, "wb") as f:
        for k, r in enumerate(recs):
            if k >= 12000 and rng.random() < 0.5:                      # shorter reads after a long uniform prefix
                L = int(rng.integers(5, 100))
                r = [r[0], r[1][:L], r[2], r[3][:L]]
            f.write(b"\n".join(r) + b"\n")
    opts = [["-a", "AGATCGGAAGAGC", "-l", "10", "-v"], ["-a", "AGATCGGAAGAGC", "-n", "-c", "-v"], ["-a", "AGATCGGAAGAGC", "-C", "-v"],
            ["-a", "AGATCGGAAGAGC", "-k", "-v"], ["-a", "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC", "-d", "3", "-M", "4", "-n", "-v"]]
    for env in [dict(), dict(FASTX_WINDOW_BYTES="400000", FASTX_CHUNK_BYTES="150000"), dict(FASTX_TEXT_PATH="0")]:
        os.environ.update(env)
        try:
            for o in opts:
                assert_same("fastx_clipper", o + ["-i", fq])
                assert_same("fastx_clipper", o + ["-i", mixed])
            assert_same("fastx_clipper", ["-a", "AGATCGGAAGAGC", "-v"], stdin=open(mixed, "rb").read())
        finally:
            for k in env:
                os.environ.pop(k, None)


@gpu
@needs_ref
def test_clipper_fallback_right_after_a_text_chunk_sees_the_stale_bases(tmp_path):
    """the GPU text chunk ends exactly where the read length drops: the record path that takes over starts with a SHORT read
    and must find, behind its end, the bases of the last long read (the reference aligner's grow-only query buffer) — here
    an adapter, so with -C the short reads are dropped as 'clipped' exactly as the reference drops them"""
    rng = np.random.default_rng(77)
    ad = b"AGATCGGAAGAGC"
    n1, L = 300, 100
    first = b""
    for k in range(n1):
        seq = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
        if k == n1 - 1:
            seq = seq[:60] + ad + seq[60 + len(ad):]
        first += b"@r%d\n%s\n+\n%s\n" % (k, seq, b"I" * L)
    rest = b""
    for k in range(200):
        Ls = int(rng.integers(20, 50))
        rest += b"@s%d\n%s\n+\n%s\n" % (k, bytes(rng.choice(list(b"ACGT"), size=Ls).astype(np.uint8)), b"I" * Ls)
    assert len(first) >= 16384
    fq = str(tmp_path / "switch.fq")
    open(fq, "wb").write(first + rest)
    os.environ.update(FASTX_CHUNK_BYTES=str(len(first)), FASTX_WINDOW_BYTES=str(4 * len(first)))
    try:
        for o in (["-C", "-v"], ["-c", "-v"], ["-v"], ["-n", "-l", "5", "-v"]):
            assert_same("fastx_clipper", ["-a", ad.decode()] + o + ["-i", fq])
    finally:
        os.environ.pop("FASTX_CHUNK_BYTES"); os.environ.pop("FASTX_WINDOW_BYTES")


@gpu
@needs_ref
def test_output_file_report_stream_and_gzip(tmp_path):
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 5000, 100, H.PLAIN)
    o1, o2 = str(tmp_path / "mine.fq"), str(tmp_path / "ref.fq")
    m = run_tool(os.path.join(BIN, "fastq_quality_trimmer"), ["-t", "20", "-l", "20", "-v", "-i", fq, "-o", o1])
    r = run_tool(H.ref_tool("fastq_quality_trimmer"), ["-t", "20", "-l", "20", "-v", "-i", fq, "-o", o2])
    assert m == r and open(o1, "rb").read() == open(o2, "rb").read()       # report goes to stdout with -o
    z1, z2 = str(tmp_path / "mine.gz"), str(tmp_path / "ref.gz")
    m = run_tool(os.path.join(BIN, "fastq_quality_filter"), ["-q", "20", "-p", "80", "-z", "-i", fq, "-o", z1])
    r = run_tool(H.ref_tool("fastq_quality_filter"), ["-q", "20", "-p", "80", "-z", "-i", fq, "-o", z2])
    assert m[0] == r[0] == 0
    import time
    time.sleep(0.5)   # the reference does not wait for its gzip child
    assert gzip.open(z1).read() == gzip.open(z2).read()


@gpu
@needs_ref
def test_gzip_output_is_deflated_on_the_gpu(tmp_path):
    """-z: the emitted text leaves the GPU as DEFLATE blocks (fxg_deflate.cu), the writer frames the gzip stream; gunzip must
    give exactly the reference's payload, for several chunk sizes, tools and a host-path tail"""
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 60000, 150, H.WITH_N)
    text = open(fq, "rb").read()
    for env in (dict(), dict(FASTX_CHUNK_BYTES="300000", FASTX_WORKERS="3"), dict(FASTX_CHUNK_BYTES="70000")):
        os.environ.update(env)
        try:
            for tool, args in (("fastq_quality_trimmer", ["-t", "20", "-l", "20"]), ("fastq_quality_filter", ["-q", "20", "-p", "80"]),
                               ("fastx_reverse_complement", []), ("fastx_clipper", ["-a", "AGATCGGAAGAGC", "-l", "20"])):
                z1, z2 = str(tmp_path / "mine.gz"), str(tmp_path / "ref.gz")
                m = run_tool(os.path.join(BIN, tool), args + ["-z", "-i", fq, "-o", z1])
                r = run_tool(H.ref_tool(tool), args + ["-z", "-i", fq, "-o", z2])
                assert m[0] == r[0] == 0, m[2][-300:]
                import time
                time.sleep(0.3)   # the reference does not wait for its gzip child
                a, b = gzip.open(z1).read(), gzip.open(z2).read()
                assert a == b, (tool, env, len(a), len(b))
                assert os.path.getsize(z1) < 0.62 * len(a), (tool, os.path.getsize(z1), len(a))      # really compressed
            # stdout, and a file whose last record has no newline (host path writes the tail as a stored block)
            m = run_tool(os.path.join(BIN, "fastq_quality_trimmer"), ["-t", "20", "-l", "20", "-z"], stdin=text[:-1])
            r = run_tool(H.ref_tool("fastq_quality_trimmer"), ["-t", "20", "-l", "20", "-z"], stdin=text[:-1])
            assert m[0] == r[0] == 0 and gzip.decompress(m[1]) == gzip.decompress(r[1])
        finally:
            for k in env:
                os.environ.pop(k, None)
    # the standard tool must accept the stream, too
    import subprocess
    assert subprocess.run(["gzip", "-t", z1]).returncode == 0


@gpu
@needs_ref
def test_numeric_quality_and_fasta_inputs(tmp_path):
    G = H.GOLDEN
    num = os.path.join(G, "fastx_rev_comp2.fastq")
    assert_same("fastq_quality_trimmer", ["-t", "20", "-l", "5", "-i", num])
    assert_same("fastq_quality_filter", ["-q", "10", "-p", "50", "-v", "-i", num])
    assert_same("fastx_quality_stats", ["-i", num])
    assert_same("fastx_clipper", ["-a", "AGATCGG", "-l", "5", "-v", "-i", num])
    assert_same("fastx_collapser", ["-i", num])
    fa = str(tmp_path / "in.fa")
    seq, _ = H.synth_slab(H.SEED_BASE + 4, 20000, 40, H.DUPS)
    H.write_fasta(fa, seq, None, 40, prefix="5-")          # ids "5-<i>": collapsed-style counts (get_reads_count)
    assert_same("fastx_reverse_complement", ["-v", "-i", fa])
    assert_same("fastx_collapser", ["-v", "-i", fa])
    assert_same("fastx_clipper", ["-a", "ACGTACGT", "-v", "-n", "-i", fa])
    # FASTA into quality_stats: the reference walks off its (empty) quality tables; the old format stays inside the
    # cycle's own entries and is reproduced, "-N" on FASTA runs off the end of the reference's static table (undefined)
    assert_same("fastx_quality_stats", ["-i", fa])
    # all of these take the GPU text path (K-NUMQ / 2-line FASTA records), not the host parser: FASTX_PATH_REPORT=1 says so
    big = str(tmp_path / "numeric.fq")
    s2, q2 = H.synth_slab(H.SEED_BASE + 5, 30000, 60, H.PLAIN)
    with open(big, "wb") as f:
        for i in range(30000):
            f.write(b"@n%d\n%s\n+\n%s\n" % (i, s2[i, :60].tobytes(), b" ".join(b"%d" % (int(v) - 33) for v in q2[i, :60])))
    os.environ["FASTX_PATH_REPORT"] = "1"
    try:
        for tool, args, path, nrec in (("fastq_quality_trimmer", ["-t", "20", "-l", "5"], big, 30000), ("fastq_quality_filter", ["-q", "20", "-p", "50"], big, 30000),
                                       ("fastx_reverse_complement", [], big, 30000), ("fastx_quality_stats", [], big, 30000),
                                       ("fastx_collapser", [], big, 30000), ("fastx_reverse_complement", [], fa, 20000),
                                       ("fastx_collapser", [], fa, 20000), ("fastx_clipper", ["-a", "ACGTACGT", "-n"], fa, 20000)):
            m = run_tool(os.path.join(BIN, tool), args + ["-i", path])
            r = run_tool(H.ref_tool(tool), args + ["-i", path])
            assert m[0] == r[0] == 0 and m[1] == r[1], (tool, path)
            assert ("[path] gpu_text_records=%d fallback=0" % nrec).encode() in m[2], (tool, path, m[2][-200:])
    finally:
        os.environ.pop("FASTX_PATH_REPORT", None)


@gpu
@needs_ref
def test_clipper_binaries(tmp_path):
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 20000, 60, H.ADAPTER)
    for args in (["-a", "AGATCGGAAGAGC", "-l", "20", "-v"], ["-a", "AGATCGGAAGAGC", "-l", "20", "-n", "-c", "-v"],
                 ["-a", "AGATCGGAAGAGC", "-C", "-v"], ["-a", "AGATCGGAAGAGC", "-d", "3", "-n"], ["-a", "AGATCGGAAGAGC", "-k", "-v"],
                 ["-a", "AGATCGGAAGAGC", "-M", "8", "-v"], ["-v"]):
        assert_same("fastx_clipper", args + ["-i", fq])
    # mixed lengths: the reference's grow-only matrix reads stale bytes of earlier reads
    fq2 = str(tmp_path / "mixed.fq")
    synth_fastq(fq2, 20000, 60, H.ADAPTER, np.random.default_rng(3))
    os.environ["FASTX_BATCH_READS"] = "3001"
    try:
        for args in (["-a", "AGATCGGAAGAGC", "-l", "5", "-C", "-n", "-v"], ["-a", "AGATCGGAAGAGC", "-l", "5", "-c", "-v"],
                     ["-a", "AGATCGGAAGAGC", "-l", "5", "-v"]):
            assert_same("fastx_clipper", args + ["-i", fq2])
    finally:
        os.environ.pop("FASTX_BATCH_READS", None)


@gpu
@needs_ref
def test_collapser_binary_large(tmp_path):
    fa = str(tmp_path / "in.fa")
    seq, _ = H.synth_slab(H.SEED_BASE + 4, 300000, 50, H.DUPS)
    H.write_fasta(fa, seq, None, 50)
    assert_same("fastx_collapser", ["-v", "-i", fa])
    fq = str(tmp_path / "in.fq")
    synth_fastq(fq, 50000, 36, H.DUPS)
    assert_same("fastx_collapser", ["-i", fq])


@gpu
@needs_ref
def test_broken_inputs_fail_like_the_reference(tmp_path):
    """prefix of the output, then the reference's message and exit status 1"""
    seq, qual = H.synth_slab(H.SEED_BASE + 12, 12000, 50, H.PLAIN)
    base = str(tmp_path / "good.fq")
    H.write_fastq(base, seq, qual, None, 50)
    lines = open(base, "rb").read().split(b"\n")[:-1]

    def variant(name, edit):
        ls = list(lines)
        edit(ls)
        p = str(tmp_path / name)
        open(p, "wb").write(b"\n".join(ls) + b"\n")
        return p

    def set_line(i, v):
        return lambda ls: ls.__setitem__(i, v)

    bad = [
        variant("badbase.fq", set_line(4 * 9000 + 1, b"ACGTACGTxCGT" + b"A" * 38)),
        variant("lower.fq", set_line(4 * 10 + 1, b"acgt" * 12 + b"AC")),
        variant("badqual_low.fq", set_line(4 * 7000 + 3, b"I" * 49 + b"\x05")),
        variant("badqual_hi.fq", set_line(4 * 3 + 3, b"I" * 20 + b"\x7f" + b"I" * 29)),
        variant("noat.fq", set_line(4 * 5000, b"r5000")),
        variant("emptyseq.fq", lambda ls: (ls.__setitem__(4 * 100 + 1, b""), ls.__setitem__(4 * 100 + 3, b""))),
        variant("trunc3.fq", lambda ls: ls.__delitem__(slice(4 * 11000 + 2, None))),
        variant("trunc4.fq", lambda ls: ls.__delitem__(slice(4 * 11000 + 3, None))),
        variant("blankend.fq", lambda ls: ls.append(b"")),
        variant("qualshort.fq", set_line(4 * 8000 + 3, b"I" * 30)),
        variant("two_errors.fq", lambda ls: (ls.__setitem__(4 * 6000 + 3, b"I" * 49 + b"\x01"), ls.__setitem__(4 * 2000 + 1, b"N" * 49 + b"U"))),
        variant("base_and_trunc.fq", lambda ls: (ls.__setitem__(4 * 11000 + 1, b"ACGU" + b"A" * 46), ls.__delitem__(slice(4 * 11000 + 3, None)))),
        # chomp() ends a line at its FIRST carriage return (chomp.c:34-44): inside a name the rest is dropped, inside the
        # sequence / quality line the record is short
        variant("cr_in_name.fq", set_line(4 * 4000, b"@r40\r00 tail")),
        variant("cr_in_name2.fq", set_line(4 * 4000 + 2, b"+r40\r00")),
        variant("cr_in_seq.fq", set_line(4 * 9500 + 1, b"ACGTACGT\rCGT" + b"A" * 38)),
        variant("cr_in_qual.fq", set_line(4 * 9500 + 3, b"I" * 20 + b"\r" + b"I" * 29)),
    ]
    os.environ["FASTX_BATCH_READS"] = "2500"
    try:
        for p in bad:
            assert_same("fastq_quality_trimmer", ["-t", "20", "-l", "10", "-i", p])
            assert_same("fastq_quality_filter", ["-q", "20", "-p", "50", "-i", p])
            assert_same("fastx_reverse_complement", ["-i", p])
            assert_same("fastx_clipper", ["-a", "AGATCGGAAGAGC", "-n", "-i", p])
            assert_same("fastx_quality_stats", ["-i", p])
            assert_same("fastx_collapser", ["-i", p])
    finally:
        os.environ.pop("FASTX_BATCH_READS", None)
