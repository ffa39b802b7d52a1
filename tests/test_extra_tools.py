"""fastx_trimmer, fastq_masker, fastx_artifacts_filter — oracle pinned to the reference's
fixtures and binaries (CPU); kernels vs oracle and binaries vs reference binaries (-m gpu)."""
import os

import numpy as np
import pytest

import helpers as H
from helpers import GOLDEN
from test_oracle_golden import emit, golden
from test_tools_cli import BIN, assert_same, needs_bin, needs_ref, run_tool

EXTRA = ["fastx_trimmer", "fastq_masker", "fastx_artifacts_filter", "fastq_to_fasta"]


def fastq_to_fasta_expected(recs, has_n, keep_n, rename):
    """src/fastq_to_fasta/fastq_to_fasta.c:79-88 over parsed records, with the oracle's N flags"""
    out, k = [], 0
    for (name, s, _n2, _q), f in zip(recs, has_n):
        if f and not keep_n:
            continue
        k += 1
        out.append(b">" + (b"%d" % k if rename else name) + b"\n" + s + b"\n")
    return b"".join(out)


def fastx_trimmer_expected(recs, first, last, t, m, q_offset):
    out = []
    for name, s, name2, q in recs:
        nl, st = H.o_fastx_trimmer(len(s), first, last, t, m)
        if nl < 0:
            continue
        if q is None:
            out.append(b">" + name + b"\n" + s[st:st + nl] + b"\n")
        else:
            ql = q[st:st + nl] if len(q) == len(s) else b" ".join(b"%d" % int(t) for t in q.split()[st:st + nl])
            out.append(b"@" + name + b"\n" + s[st:st + nl] + b"\n+" + name2 + b"\n" + ql + b"\n")
    return b"".join(out)


def test_oracle_golden_extra():
    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_trimmer1.fasta"))
    assert fastx_trimmer_expected(recs, 5, 36, 0, 0, 33) == golden("fastx_trimmer1.out")
    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_trimmer2.fastq"))
    assert fastx_trimmer_expected(recs, 1, 27, 0, 0, 33) == golden("fastx_trimmer2.out")
    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_trimmer_from_end1.fasta"))
    assert fastx_trimmer_expected(recs, 1, 0, 2, 16, 33) == golden("fastx_trimmer_from_end1.out")
    recs = H.read_fastx(os.path.join(GOLDEN, "fastq_masker.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    oseq, flag, mr, mb = H.o_mask(seq, qual, lens, 0, stride, 64, 29, ord("x"))
    recs2 = [(r[0], oseq[i, :lens[i]].tobytes(), r[2], r[3]) for i, r in enumerate(recs)]
    assert emit(recs2, lens, 64) == golden("fastq_masker.out")
    # fastq_to_fasta: the .xml's two tests are (default flags) and (-n -r) (galaxy/tools/fastx_toolkit/fastq_to_fasta.xml:28-43)
    recs = H.read_fastx(os.path.join(GOLDEN, "fastq_to_fasta1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    fl = H.o_has_n(seq, lens, 0, stride)
    assert fastq_to_fasta_expected(recs, fl, False, False) == golden("fastq_to_fasta1a.out")
    assert fastq_to_fasta_expected(recs, fl, True, True) == golden("fastq_to_fasta1b.out")
    for fin, fout in (("fastx_artifacts1.fasta", "fastx_artifacts1.out"), ("fastx_artifacts2.fastq", "fastx_artifacts2.out")):
        recs = H.read_fastx(os.path.join(GOLDEN, fin))
        seq, qual, lens, stride, _ = H.slab_from_records(recs, 33)
        keep = H.o_artifacts(seq, lens, 0, stride)
        assert emit(recs, np.where(keep != 0, lens, -1), 33, fastq=recs[0][3] is not None) == golden(fout)


@needs_ref
@needs_bin
@pytest.mark.parametrize("tool", EXTRA)
def test_extra_usage_and_flag_errors(tool, tmp_path):
    assert_same(tool, ["-h"])
    assert_same(tool, ["-Z"])
    assert_same(tool, ["-i", str(tmp_path / "missing.fq")])
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a\nACGT\n")
    if tool == "fastx_trimmer":
        for a in (["-f", "0"], ["-l", "25000"], ["-t", "0"], ["-m", "0"], ["-f", "2", "-t", "3"]):
            assert_same(tool, a + ["-i", str(fa)])
    if tool == "fastq_to_fasta":
        assert_same(tool, ["-i", str(fa)])      # FASTA into a FASTQ-only tool
        assert_same(tool, ["-x"])
    if tool == "fastq_masker":
        assert_same(tool, ["-q", "-41", "-i", str(fa)])
        assert_same(tool, ["-r", "xy", "-i", str(fa)])
        assert_same(tool, ["-i", str(fa)])      # FASTA into a FASTQ-only tool


# ------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@gpu
def test_extra_kernels_vs_oracle():
    torch = pytest.importorskip("torch")
    import fastx_toolkit_b200 as F
    ctx = F.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for L, n, ragged in ((150, 30001, False), (36, 9000, True), (100, 5000, True), (7, 3000, False), (250, 2000, True)):
        seq, qual = H.synth_slab(H.SEED_BASE + 15, n, L, H.WITH_N)
        rng = np.random.default_rng(L)
        # make some artifacts: long homopolymers with a few other bases
        for i in rng.choice(n, n // 4, replace=False):
            b = b"ACGT"[int(rng.integers(0, 4))]
            seq[i, :L] = b
            for _ in range(int(rng.integers(0, 6))):
                seq[i, int(rng.integers(0, L))] = b"ACGTN"[int(rng.integers(0, 5))]
        lens = H.ragged(seq, qual, rng, min_len=1) if ragged else None
        stride = seq.shape[1]
        dseq, dqual = torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda()
        dlens = None if lens is None else torch.from_numpy(lens).cuda()
        b = ctx.batch(dseq, dqual, n, stride, L if lens is None else 0, dlens)
        keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        ctx.report_reset()
        ctx.artifacts_dev(b, 33, keep)
        rep = ctx.sync()
        exp = H.o_artifacts(seq, lens, L, stride)
        assert np.array_equal(keep.cpu().numpy(), exp) and rep.n_out == int(exp.sum()) and rep.first_bad_read == -1
        for q, ch in ((20, ord("N")), (29, ord("x")), (-40, ord(".")), (41, ord("n"))):
            oseq = torch.full((n, stride), 0xEE, dtype=torch.uint8, device="cuda")
            flag = torch.empty(n, dtype=torch.uint8, device="cuda")
            ctx.report_reset()
            ctx.mask_dev(b, 33, q, ch, oseq, flag)
            rep = ctx.sync()
            eseq, eflag, mr, mb = H.o_mask(seq, qual, lens, L, stride, 33, q, ch)
            assert np.array_equal(oseq.cpu().numpy(), eseq) and np.array_equal(flag.cpu().numpy(), eflag)
            assert rep.n_out == mr and rep.aux[0] == mb and rep.first_bad_read == -1
        # validation alone (and through the host pipeline), FASTA form too
        s2, q2 = seq.copy(), qual.copy()
        bad_i = n // 2
        ll = L if lens is None else int(lens[bad_i])
        s2[bad_i, ll - 1] = ord("u")
        ds2 = torch.from_numpy(s2).cuda()
        ctx.report_reset()
        ctx.validate_dev(ctx.batch(ds2, None, n, stride, L if lens is None else 0, dlens), 33)
        assert ctx.sync().first_bad_read == bad_i
        rep = ctx.validate_host(ctx.batch(s2, q2, n, stride, L if lens is None else 0, lens), 33)
        assert rep.first_bad_read == bad_i
        rep = ctx.validate_host(ctx.batch(seq, qual, n, stride, L if lens is None else 0, lens), 33)
        assert rep.first_bad_read == -1 and rep.n_in == n
        kh = np.empty(n, np.uint8)
        rep = ctx.artifacts_host(ctx.batch(seq, qual, n, stride, L if lens is None else 0, lens), 33, kh)
        assert np.array_equal(kh, exp)
        # K-HASN (fastq_to_fasta's discard test), junk in the padding must not count
        jseq = seq.copy()
        if lens is not None:
            for i in range(0, n, 3):
                jseq[i, int(lens[i]):] = ord("N")
        djs = torch.from_numpy(jseq).cuda()
        fl = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
        ctx.report_reset()
        ctx.has_n_dev(ctx.batch(djs, dqual, n, stride, L if lens is None else 0, dlens), 33, fl)
        rep = ctx.sync()
        efl = H.o_has_n(jseq, lens, L, stride)
        assert np.array_equal(fl.cpu().numpy(), efl) and rep.n_out == int(efl.sum()) and rep.first_bad_read == -1
        fh = np.empty(n, np.uint8)
        rep = ctx.has_n_host(ctx.batch(jseq, None, n, stride, L if lens is None else 0, lens), 33, fh)
        assert np.array_equal(fh, efl) and rep.n_in == n
    ctx.close()


@gpu
@needs_ref
def test_extra_binaries_vs_reference(tmp_path):
    G = GOLDEN
    for tool, args, fin, fout in (
            ("fastx_trimmer", ["-f", "5", "-l", "36"], "fastx_trimmer1.fasta", "fastx_trimmer1.out"),
            ("fastx_trimmer", ["-Q", "64", "-f", "1", "-l", "27"], "fastx_trimmer2.fastq", "fastx_trimmer2.out"),
            ("fastx_trimmer", ["-t", "2", "-m", "16"], "fastx_trimmer_from_end1.fasta", "fastx_trimmer_from_end1.out"),
            ("fastq_masker", ["-Q", "64", "-q", "29", "-r", "x"], "fastq_masker.fastq", "fastq_masker.out"),
            ("fastx_artifacts_filter", [], "fastx_artifacts1.fasta", "fastx_artifacts1.out"),
            ("fastx_artifacts_filter", [], "fastx_artifacts2.fastq", "fastx_artifacts2.out"),
            ("fastq_to_fasta", ["-Q", "64"], "fastq_to_fasta1.fastq", "fastq_to_fasta1a.out"),
            ("fastq_to_fasta", ["-Q", "64", "-n", "-r"], "fastq_to_fasta1.fastq", "fastq_to_fasta1b.out")):
        rc, out, errs = run_tool(os.path.join(BIN, tool), args + ["-i", os.path.join(G, fin)])
        assert rc == 0 and out == open(os.path.join(G, fout), "rb").read(), (tool, args, errs)
        assert_same(tool, args + ["-v", "-i", os.path.join(G, fin)])
    fq = str(tmp_path / "in.fq")
    seq, qual = H.synth_slab(H.SEED_BASE + 16, 30000, 75, H.WITH_N)
    rng = np.random.default_rng(8)
    for i in rng.choice(30000, 6000, replace=False):
        seq[i, :75] = b"ACGT"[int(rng.integers(0, 4))]
        seq[i, int(rng.integers(0, 75))] = ord("N")
    lens = H.ragged(seq, qual, rng, min_len=3)
    H.write_fastq(fq, seq, qual, lens, 75)
    fa = str(tmp_path / "in.fa")
    H.write_fasta(fa, seq, lens, 75, prefix="7-")
    os.environ["FASTX_BATCH_READS"] = "7001"
    try:
        for args in (["-f", "3", "-l", "40", "-v"], ["-f", "10"], ["-l", "20", "-v"], ["-t", "5", "-m", "30", "-v"], ["-t", "1"], ["-v"]):
            assert_same("fastx_trimmer", args + ["-i", fq])
            assert_same("fastx_trimmer", args + ["-i", fa])
        for args in (["-v"], ["-q", "25", "-r", ".", "-v"], ["-q", "-40"], ["-q", "41", "-r", "n", "-v"]):
            assert_same("fastq_masker", args + ["-i", fq])
        assert_same("fastx_artifacts_filter", ["-v", "-i", fq])
        assert_same("fastx_artifacts_filter", ["-v", "-i", fa])
        for args in (["-v"], ["-n", "-v"], ["-r"], ["-n", "-r", "-v"], ["-z", "-r"]):
            assert_same("fastq_to_fasta", args + ["-i", fq])
        assert_same("fastq_to_fasta", ["-v", "-r", "-i", fq, "-o", str(tmp_path / "o.fa")])
        assert_same("fastq_to_fasta", ["-n", "-i", os.path.join(G, "fastx_rev_comp2.fastq")])                # numeric qualities
        assert_same("fastx_trimmer", ["-f", "2", "-l", "9", "-i", os.path.join(G, "fastx_rev_comp2.fastq")])   # numeric qualities
        assert_same("fastq_masker", ["-q", "20", "-v", "-i", os.path.join(G, "fastx_rev_comp2.fastq")])
        # a broken record: prefix of the output, then the reference's message
        lines = open(fq, "rb").read().split(b"\n")[:-1]
        lines[4 * 20000 + 1] = b"ACGTxACGT"; lines[4 * 20000 + 3] = b"IIIIIIIII"
        bad = str(tmp_path / "bad.fq")
        open(bad, "wb").write(b"\n".join(lines) + b"\n")
        for tool, args in (("fastx_trimmer", ["-f", "2"]), ("fastq_masker", []), ("fastx_artifacts_filter", []), ("fastq_to_fasta", ["-r"])):
            assert_same(tool, args + ["-i", bad])
    finally:
        os.environ.pop("FASTX_BATCH_READS", None)
