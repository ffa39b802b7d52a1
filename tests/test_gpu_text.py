"""GPU parity (-m gpu): the text path (K-LINES / K-RECS / K-PACK / op / K-EMIT, fxg_text_run_host) — output text must be
byte-identical to what the reference writer would emit for the oracle's decisions; anything unusual must be flagged."""
import numpy as np
import pytest

import helpers as H
from test_gpu_parity import ctx  # noqa: F401
from test_oracle_golden import emit

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def fastq_bytes(seq, qual, lens, L, crlf=False, plus_names=False):
    out = []
    nl = b"\r\n" if crlf else b"\n"
    for i in range(seq.shape[0]):
        l = int(lens[i]) if lens is not None else L
        name = b"r%d some comment" % i
        out.append(b"@" + name + nl + seq[i, :l].tobytes() + nl + (b"+" + name if plus_names and i % 3 == 0 else b"+") + nl + qual[i, :l].tobytes() + nl)
    return b"".join(out)


def expected(text, q_offset, op, a0, a1):
    import io, tempfile, os
    p = tempfile.mktemp(suffix=".fq")
    open(p, "wb").write(text)
    recs = H.read_fastx(p)
    os.unlink(p)
    seq, qual, lens, stride, _ = H.slab_from_records(recs, q_offset)
    if op == 0:
        out, bad = H.o_trim(seq, qual, lens, 0, stride, q_offset, a0, a1)
    else:
        keep, bad = H.o_filter(seq, qual, lens, 0, stride, q_offset, a0, a1)
        out = np.where(keep != 0, lens, -1)
    return emit(recs, out, q_offset), int((out >= 0).sum()), len(recs)


@pytest.mark.parametrize("L,ragged,crlf", [(150, False, False), (100, True, False), (50, False, True), (37, True, True)])
def test_text_path_matches_reference_writer(ctx, L, ragged, crlf):
    import fastx_toolkit_b200 as F
    n = 20000
    seq, qual = H.synth_slab(H.SEED_BASE + 13, n, L, H.WITH_N)
    lens = H.ragged(seq, qual, np.random.default_rng(L), min_len=1) if ragged else None
    text = fastq_bytes(seq, qual, lens, L, crlf, plus_names=True)
    tp = F.TextPipe(ctx, len(text) + 4096)
    for op, a0, a1 in ((0, 20, 20), (0, 35, 0), (1, 20, 90), (1, 30, 50)):
        got, rep = tp.run(op, text, 33, a0, a1)
        exp, kept, nrec = expected(text, 33, op, a0, a1)
        assert rep.anomaly == 0 and rep.n_records == nrec == n and rep.consumed_bytes == len(text)
        assert rep.n_out_records == kept and rep.out_bytes == len(exp)
        assert got == exp
    # a chunk that ends in the middle of a record: only the complete records are consumed
    cut = len(text) - 57
    got, rep = tp.run(0, text[:cut], 33, 20, 20)
    assert rep.n_records == n - 1 and text[rep.consumed_bytes - 1:rep.consumed_bytes] == b"\n"
    exp, kept, _ = expected(text[:rep.consumed_bytes], 33, 0, 20, 20)
    assert got == exp and rep.n_out_records == kept
    # last line without a trailing newline: that record is left to the caller
    got, rep = tp.run(0, text.rstrip(b"\r\n"), 33, 20, 20)
    assert rep.n_records == n - 1
    assert tp.L.fxg_text_launches(tp.h) > 0
    tp.close()


def test_text_path_revcomp_and_stats(ctx):
    import ctypes as C
    import fastx_toolkit_b200 as F
    n, L = 15000, 75
    seq, qual = H.synth_slab(H.SEED_BASE + 14, n, L, H.WITH_N)
    lens = H.ragged(seq, qual, np.random.default_rng(4), min_len=1)
    text = fastq_bytes(seq, qual, lens, L, crlf=False, plus_names=True)
    tp = F.TextPipe(ctx, len(text) + 4096)
    got, rep = tp.run(2, text, 33, 0, 0)
    eseq, equal = H.o_revcomp(seq, qual, lens, 0, seq.shape[1])
    exp = fastq_bytes(eseq, equal, lens, L, plus_names=True)
    assert rep.anomaly == 0 and rep.n_out_records == n and got == exp
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    r2 = F.TextReport()
    src = np.frombuffer(text, np.uint8)
    rc = tp.L.fxg_text_stats_host(tp.h, src.ctypes.data, src.size, 33, hist.data_ptr(), L, C.byref(r2))
    assert rc == 0 and r2.anomaly == 0 and r2.n_records == n and r2.max_len == int(lens.max())
    eh, _ = H.o_stats_hist(seq, qual, lens, 0, seq.shape[1], 33, L)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), eh)
    tp.close()


def test_text_path_flags_anomalies(ctx):
    import fastx_toolkit_b200 as F
    n, L = 5000, 60
    seq, qual = H.synth_slab(H.SEED_BASE + 13, n, L, H.PLAIN)
    lines = fastq_bytes(seq, qual, None, L).split(b"\n")[:-1]
    tp = F.TextPipe(ctx, 4 << 20)

    def run(edit):
        ls = list(lines)
        edit(ls)
        return tp.run(0, b"\n".join(ls) + b"\n", 33, 20, 20)[1]

    rep = run(lambda ls: ls.__setitem__(4 * 1234 + 3, b"40 " * 59 + b"40"))        # ONE numeric quality line in an ASCII chunk
    assert (rep.anomaly, rep.anomaly_record) == (3, 1234)
    rep = run(lambda ls: ls.__setitem__(4 * 77, b"r77"))                             # no '@'
    assert (rep.anomaly, rep.anomaly_record) == (1, 77)
    rep = run(lambda ls: (ls.__setitem__(4 * 9 + 1, b""), ls.__setitem__(4 * 9 + 3, b"")))
    assert (rep.anomaly, rep.anomaly_record) == (2, 9)
    rep = run(lambda ls: ls.__setitem__(4 * 4000 + 1, b"ACGU" + b"A" * 56))          # illegal base -> op kernel
    assert (rep.anomaly, rep.anomaly_record) == (5, 4000)
    rep = run(lambda ls: ls.__setitem__(4 * 4001 + 3, b"I" * 59 + b"\x07"))          # illegal quality
    assert (rep.anomaly, rep.anomaly_record) == (5, 4001)
    rep = run(lambda ls: (ls.__setitem__(4 * 10 + 3, b"\x07" * 60), ls.__setitem__(4 * 5, b"x")))   # earliest wins per stage
    assert (rep.anomaly, rep.anomaly_record) == (1, 5)
    rep = run(lambda ls: None)
    assert rep.anomaly == 0 and rep.n_records == n
    tp.close()


def test_text_path_clipper(ctx):
    """fxg_text_clip_host (fastx_clipper.cpp:257-320 on whole text chunks): equal-length chunks are clipped and emitted on
    the GPU, anything else is handed back as FXG_TEXT_MIXED_LEN"""
    import tempfile, os
    import fastx_toolkit_b200 as F
    n, L = 20000, 100
    adapter = b"AGATCGGAAGAGC"
    seq, qual = H.synth_slab(H.SEED_BASE + 2, n, L, H.ADAPTER)
    text = fastq_bytes(seq, qual, None, L, plus_names=True)
    p = tempfile.mktemp(suffix=".fq")
    open(p, "wb").write(text)
    recs = H.read_fastx(p)
    os.unlink(p)
    tp = F.TextPipe(ctx, len(text) + 4096)
    for kw, k in ((dict(min_length=20), 0), (dict(min_length=5, discard_non_clipped=1, discard_unknown=0), 0),
                  (dict(min_length=5, discard_clipped=1), 0), (dict(min_length=10, keep_delta=16, min_adapter_len=6), 0),
                  (dict(min_length=5), 1)):
        full = dict(min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
        full.update(kw)
        e_len, e_cls, _ = H.o_clip(seq, None, None, L, seq.shape[1], adapter, H.FxoClipOpts(**full))
        if k:
            out = np.where(e_cls == 1, L, -1)
        else:
            out = np.where(e_cls == 0, e_len, -1)
        exp = emit(recs, out, 33)
        got, rep = tp.clip(text, 33, F.ClipOpts(adapter=adapter, **full), k, 0)
        assert rep.anomaly == 0 and rep.n_records == n and rep.min_len == rep.max_len == L
        assert [rep.clip_class[c] for c in range(6)] == [int((e_cls == c).sum()) for c in range(6)]
        assert rep.n_out_records == int((out >= 0).sum()) and got == exp
    o = F.ClipOpts(adapter=adapter, min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    # earlier reads were longer / shorter than this chunk's: host path
    assert tp.clip(text, 33, o, 0, L + 1)[1].anomaly == 7
    assert tp.clip(text, 33, o, 0, L)[1].anomaly == 0
    # one short read inside the chunk
    lens = np.full(n, L, np.int32); lens[n // 2] = L - 3
    assert tp.clip(fastq_bytes(seq, qual, lens, L), 33, o, 0, 0)[1].anomaly == 7
    # illegal base is still reported by the op kernel
    s2 = seq.copy(); s2[321, 5] = ord("x")
    rep = tp.clip(fastq_bytes(s2, qual, None, L), 33, o, 0, 0)[1]
    assert (rep.anomaly, rep.anomaly_record) == (5, 321)
    tp.close()


def numeric_fastq(seq, qual, lens, L, q_offset=33, style=0):
    """4-line records with NUMERIC quality lines (fastx.c:137-167): style 1 writes signs / leading zeros / tabs the way strtol
    accepts them; the reference prints plain %d back"""
    out = []
    for i in range(seq.shape[0]):
        l = int(lens[i]) if lens is not None else L
        vals = [int(v) - q_offset for v in qual[i, :l]]
        if style == 0:
            ql = b" ".join(b"%d" % v for v in vals)
        else:
            ql = b"".join((b" " if k else b"") + (b"\t" if k % 7 == 3 else b"") + (b"+%d" % v if (v >= 0 and k % 5 == 0) else b"%03d" % v if (v >= 0 and k % 11 == 1) else b"%d" % v)
                          for k, v in enumerate(vals))
        out.append(b"@n%d\n" % i + seq[i, :l].tobytes() + b"\n+n%d\n" % i + ql + b"\n")
    return b"".join(out)


def test_text_path_numeric_qualities(ctx):
    """records with numeric quality lines are parsed (K-NUMQ), processed and written back in numeric form on the GPU"""
    import tempfile, os
    import fastx_toolkit_b200 as F
    n, L = 12000, 60
    seq, qual = H.synth_slab(H.SEED_BASE + 15, n, L, H.WITH_N)
    rng = np.random.default_rng(8)
    qual[:, :L] = (rng.integers(-15, 94, size=(n, L)) + 33).astype(np.uint8)          # the whole legal range, negatives included
    lens = H.ragged(seq, qual, rng, min_len=2)
    lens[lens == 1] = 2
    for style in (0, 1):
        text = numeric_fastq(seq, qual, lens, L, 33, style)
        p = tempfile.mktemp(suffix=".fq")
        open(p, "wb").write(text)
        recs = H.read_fastx(p)
        os.unlink(p)
        tp = F.TextPipe(ctx, len(text) + 4096)
        for op, a0, a1 in ((0, 20, 10), (1, 10, 60), (0, -5, 3)):
            got, rep = tp.run(op, text, 64, a0, a1)           # -Q is irrelevant for numeric records
            if op == 0:
                out, bad = H.o_trim(seq, qual, lens, 0, seq.shape[1], 33, a0, a1)
            else:
                keep, bad = H.o_filter(seq, qual, lens, 0, seq.shape[1], 33, a0, a1)
                out = np.where(keep != 0, lens, -1)
            exp = emit(recs, out, 33)
            assert rep.anomaly == 0 and rep.n_records == n and bad == -1
            assert rep.n_out_records == int((out >= 0).sum()) and got == exp, (style, op)
        got, rep = tp.run(2, text, 33, 0, 0)
        eseq, equal = H.o_revcomp(seq, qual, lens, 0, seq.shape[1])
        exp = numeric_fastq(eseq, equal, lens, L, 33, 0)
        assert rep.anomaly == 0 and got == exp
        assert tp.numeric_chunks() == 4 and tp.fasta_chunks() == 0
        tp.close()
    # malformed numbers, a value out of range, a missing value, trailing blank: the host parser words the message
    lines = numeric_fastq(seq[:3000], qual[:3000], lens[:3000], L).split(b"\n")[:-1]
    tp = F.TextPipe(ctx, 4 << 20)
    nb = int(lens[1500])
    for bad_line in (b"40 40 x 40", b" ".join([b"40"] * (nb - 1) + [b"94"]), b" ".join([b"40"] * (nb - 1) + [b"-16"]), b" ".join([b"40"] * (nb - 1)),
                     b" ".join([b"40"] * nb) + b" ", b" ".join([b"40"] * (nb + 1))):
        if len(bad_line) == nb:
            continue                      # same length as the sequence: an ASCII quality line by the reader's rule
        ls = list(lines)
        ls[4 * 1500 + 3] = bad_line
        rep = tp.run(0, b"\n".join(ls) + b"\n", 33, 20, 20)[1]
        assert rep.anomaly != 0 and rep.anomaly_record == 1500, bad_line
    tp.close()


def fasta_bytes(seq, lens, L, names=None):
    return b"".join(b">" + (names[i] if names else b"s%d" % i) + b"\n" + seq[i, :(int(lens[i]) if lens is not None else L)].tobytes() + b"\n"
                    for i in range(seq.shape[0]))


def test_text_path_fasta_and_collapser(ctx):
    """2-line FASTA records on the GPU text path: reverse complement, quality stats (weights from "N-COUNT" identifiers) and the
    collapser fed straight from text (FASTA and FASTQ)"""
    import fastx_toolkit_b200 as F
    n, L = 30000, 50
    seq, qual = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    rng = np.random.default_rng(3)
    lens = H.ragged(seq, qual, rng, min_len=1)
    names = [b"%d-%d" % (i, rng.integers(1, 40)) if i % 3 else b"plain%d" % i for i in range(n)]
    weights = np.array([int(nm.split(b"-")[1]) if b"-" in nm else 1 for nm in names], np.int32)
    text = fasta_bytes(seq, lens, L, names)
    tp = F.TextPipe(ctx, len(text) + 4096)
    tp.set_format(True)
    got, rep = tp.run(2, text, 33, 0, 0)
    eseq, _ = H.o_revcomp(seq, qual, lens, 0, seq.shape[1])
    assert rep.anomaly == 0 and rep.n_records == n and rep.n_reads == int(weights.sum()) == rep.n_out_reads
    assert got == fasta_bytes(eseq, lens, L, names)
    # quality stats of FASTA: only counts, weighted by the collapsed-read counts
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    rep = tp.stats(text, 33, hist, L)
    assert rep.anomaly == 0
    exp = np.zeros((L, 5), np.int64)
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3, ord("N"): 4}
    for i in range(n):
        for c in range(int(lens[i])):
            exp[c, code[int(seq[i, c])]] += int(weights[i])
    assert np.array_equal(hist.sum(dim=2).cpu().numpy(), exp)
    # the collapser: counts are the identifiers' read counts; order is the reference's
    col = F.Collapser(0, n, seq.shape[1])
    half = text[: text.index(b">", len(text) // 2)]
    r1 = tp.collapse(half, 33, col)
    r2 = tp.collapse(text[len(half):], 33, col)
    assert r1.anomaly == 0 and r2.anomaly == 0 and r1.n_records + r2.n_records == n
    u = col.finish(True)
    oseq, olen, ocnt, ofirst = np.zeros((u, seq.shape[1]), np.uint8), np.zeros(u, np.int32), np.zeros(u, np.uint64), np.zeros(u, np.int64)
    col.fetch(oseq, olen, ocnt, ofirst, None)
    col.close()
    O = H.oracle()
    oc = O.fxo_collapser_new()
    import ctypes as C
    for i in range(n):
        row = np.ascontiguousarray(seq[i])
        O.fxo_collapser_add(oc, H._p(row, H.u8p), int(lens[i]), int(weights[i]))
    eu = O.fxo_collapser_unique(oc)
    efirst, ecnt = np.empty(eu, np.int64), np.empty(eu, np.uint64)
    O.fxo_collapser_order(oc, H._p(efirst, H.i64p), H._p(ecnt, H.u64p))
    O.fxo_collapser_free(oc)
    assert u == eu and np.array_equal(ocnt, ecnt) and np.array_equal(ofirst, efirst)
    assert all(oseq[k, :olen[k]].tobytes() == seq[efirst[k], :lens[efirst[k]]].tobytes() for k in (0, 1, u // 2, u - 1))
    assert tp.fasta_chunks() == 4
    # a FASTA chunk with a problem is handed back
    bad = text.replace(b">plain0\n", b"plain0\n", 1)
    assert tp.run(2, bad, 33, 0, 0)[1].anomaly == 1
    s2 = seq.copy(); s2[777, 0] = ord("x")
    col = F.Collapser(0, n, seq.shape[1])
    rep = tp.collapse(fasta_bytes(s2, lens, L, names), 33, col)
    assert (rep.anomaly, rep.anomaly_record) == (5, 777)
    col.close()
    tp.close()
    # FASTQ into the collapser: qualities are validated as the reader validates them
    fq = fastq_bytes(seq, qual, lens, L)
    tq = F.TextPipe(ctx, len(fq) + 4096)
    col = F.Collapser(0, n, seq.shape[1])
    assert tq.collapse(fq, 33, col).anomaly == 0
    u2 = col.finish(True)
    efirst2, ecnt2 = H.o_collapse(seq, lens, 0, seq.shape[1])
    c2, f2 = np.zeros(u2, np.uint64), np.zeros(u2, np.int64)
    col.fetch(None, None, c2, f2, None)
    assert u2 == len(ecnt2) and np.array_equal(c2, ecnt2) and np.array_equal(f2, efirst2)
    col.close()
    q2 = qual.copy(); q2[4321, 0] = 7
    col = F.Collapser(0, n, seq.shape[1])
    rep = tq.collapse(fastq_bytes(seq, q2, lens, L), 33, col)
    assert (rep.anomaly, rep.anomaly_record) == (5, 4321)
    col.close(); tq.close()


def test_text_path_deflate_blocks(ctx):
    """fxg_text_set_deflate: the emitted text leaves the GPU as byte-aligned DEFLATE blocks; framed as gzip they must inflate to
    exactly the plain output, and the chained block CRCs must be the CRC-32 of that text"""
    import struct
    import zlib
    import fastx_toolkit_b200 as F
    n, L = 40000, 150
    seq, qual = H.synth_slab(H.SEED_BASE + 16, n, L, H.WITH_N)
    text = fastq_bytes(seq, qual, None, L, plus_names=True)
    tp = F.TextPipe(ctx, len(text) + 4096)
    L_ = F.lib()
    for op, a0, a1 in ((0, 20, 20), (1, 20, 90), (2, 0, 0), (0, 41, 1)):
        tp.set_deflate(False)
        plain, rep0 = tp.run(op, text, 33, a0, a1)
        tp.set_deflate(True)
        z, rep = tp.run(op, text, 33, a0, a1)
        assert rep.anomaly == 0 and rep.n_out_records == rep0.n_out_records and rep.raw_out_bytes == len(plain)
        if not plain:
            assert rep.out_bytes == 0
            continue
        assert rep.deflated == 1 and rep.out_bytes == len(z) < 0.6 * len(plain)
        crc = L_.fxg_crc32_finish(rep.out_crc32_pure, len(plain))
        assert crc == zlib.crc32(plain) & 0xFFFFFFFF
        gz = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + z + b"\x01\x00\x00\xff\xff" + struct.pack("<II", crc, len(plain) & 0xFFFFFFFF)
        import gzip
        assert gzip.decompress(gz) == plain
        # two chunks concatenate: blocks end on byte boundaries
        half = text[: text.index(b"\n@", len(text) // 2) + 1]
        z1, r1 = tp.run(op, half, 33, a0, a1)
        z2, r2 = tp.run(op, text[len(half):], 33, a0, a1)
        c12 = L_.fxg_crc32_concat(r1.out_crc32_pure, r2.out_crc32_pure, r2.raw_out_bytes)
        assert L_.fxg_crc32_finish(c12, len(plain)) == crc
        assert zlib.decompress(z1 + z2 + b"\x01\x00\x00\xff\xff", -15) == plain
    # tiny and skewed inputs: one literal dominates, a single record, every byte value in the names
    tp.set_deflate(True)
    odd = b"@" + bytes(range(33, 127)) + b"\nACGT\n+\nIIII\n" + b"@x\n" + b"A" * 150 + b"\n+\n" + b"I" * 150 + b"\n"
    z, rep = tp.run(2, odd, 33, 0, 0)
    tp.set_deflate(False)
    plain, _ = tp.run(2, odd, 33, 0, 0)
    assert zlib.decompress(z + b"\x01\x00\x00\xff\xff", -15) == plain
    tp.close()


def test_text_path_decisions_only(ctx):
    """fxg_text_decide_host: the per-record decision and the line table instead of any text"""
    import ctypes as C
    import fastx_toolkit_b200 as F
    n, L = 20000, 100
    seq, qual = H.synth_slab(H.SEED_BASE + 17, n, L, H.PLAIN)
    lens = H.ragged(seq, qual, np.random.default_rng(2), min_len=1)
    text = fastq_bytes(seq, qual, lens, L, plus_names=True)
    src = np.frombuffer(text, np.uint8)
    tp = F.TextPipe(ctx, len(text) + 4096)
    dec, starts = np.empty(n, np.int32), np.empty(4 * n, np.uint32)
    rep = F.TextReport()
    for op, a0, a1 in ((0, 20, 20), (1, 20, 80)):
        rc = tp.L.fxg_text_decide_host(tp.h, op, src.ctypes.data, src.size, 33, a0, a1, dec.ctypes.data, starts.ctypes.data, C.byref(rep))
        assert rc == 0 and rep.anomaly == 0 and rep.n_records == n
        if op == 0:
            exp, _ = H.o_trim(seq, qual, lens, 0, seq.shape[1], 33, a0, a1)
        else:
            keep, _ = H.o_filter(seq, qual, lens, 0, seq.shape[1], 33, a0, a1)
            exp = np.where(keep != 0, lens, -1)
        assert np.array_equal(dec, exp) and rep.n_out_records == int((exp >= 0).sum())
        # the line table lets the caller cut the records out of its own copy of the text
        for r in (0, 1, n // 2, n - 1):
            s1 = int(starts[4 * r + 1])
            assert text[int(starts[4 * r])] == ord("@") and text[s1:s1 + int(lens[r])] == seq[r, :lens[r]].tobytes()
    tp.close()
