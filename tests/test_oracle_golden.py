"""Pin the CPU oracle (oracle/fastx_oracle.c) against
  (1) every hot-path golden fixture the reference ships (galaxy/test-data, copied to
      tests/golden/reference_fixtures/; flags from the Galaxy tool XMLs), and
  (2) the unmodified reference binaries in oracle/_ref/ on seeded synthetic input.
CPU only.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from helpers import GOLDEN, u8p, _p


def emit(recs, keep_len, q_offset, fastq=True):
    """Text the reference writer would produce (fastx.c:440-473) for kept records."""
    out = []
    for (name, s, name2, q), kl in zip(recs, keep_len):
        if kl < 0:
            continue
        if not fastq or q is None:
            out.append(b">" + name + b"\n" + s[:kl] + b"\n")
            continue
        if len(q) == len(s):
            ql = q[:kl]
        else:
            ql = b" ".join(b"%d" % int(t) for t in q.split()[:kl])
        out.append(b"@" + name + b"\n" + s[:kl] + b"\n+" + name2 + b"\n" + ql + b"\n")
    return b"".join(out)


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def test_golden_trimmer():
    recs = H.read_fastx(os.path.join(GOLDEN, "fastq_quality_trimmer.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    out, bad = H.o_trim(seq, qual, lens, 0, stride, 64, 30, 16)
    assert bad == -1
    assert emit(recs, out, 64) == golden("fastq_quality_trimmer.out")


@pytest.mark.parametrize("q,p,outname", [(33, 100, "fastq_qual_filter1a.out"), (20, 80, "fastq_qual_filter1b.out")])
def test_golden_filter(q, p, outname):
    recs = H.read_fastx(os.path.join(GOLDEN, "fastq_qual_filter1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    keep, bad = H.o_filter(seq, qual, lens, 0, stride, 64, q, p)
    assert bad == -1
    kl = np.where(keep != 0, lens, -1)
    assert emit(recs, kl, 64) == golden(outname)


def test_golden_clipper():
    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_clipper1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    opts = H.FxoClipOpts(min_length=15, keep_delta=0, discard_non_clipped=1, discard_clipped=0,
                         discard_unknown=0, min_adapter_len=0)
    out_len, cls, cut = H.o_clip(seq, lens, None, 0, stride, b"CAATTGGTTAATCCCCCTATATA", opts)
    kl = np.where(cls == 0, out_len, -1)
    assert emit(recs, kl, 64) == golden("fastx_clipper1a.out")


def test_golden_revcomp_fasta_and_numeric_fastq():
    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_rev_comp1.fasta"))
    seq, _, lens, stride, _ = H.slab_from_records(recs)
    oseq, _ = H.o_revcomp(seq, None, lens, 0, stride)
    out = b"".join(b">" + r[0] + b"\n" + oseq[i, :lens[i]].tobytes() + b"\n" for i, r in enumerate(recs))
    assert out == golden("fastx_reverse_complement1.out")

    recs = H.read_fastx(os.path.join(GOLDEN, "fastx_rev_comp2.fastq"))
    seq, qual, lens, stride, asc = H.slab_from_records(recs, 64)
    assert not any(asc)  # numeric-quality fixture (includes "-1" and "-0" tokens)
    oseq, oqual = H.o_revcomp(seq, qual, lens, 0, stride)
    out = []
    for i, r in enumerate(recs):
        l = lens[i]
        nums = b" ".join(b"%d" % (int(np.int8(v)) - 64) for v in oqual[i, :l])
        out.append(b"@" + r[0] + b"\n" + oseq[i, :l].tobytes() + b"\n+" + r[2] + b"\n" + nums + b"\n")
    assert b"".join(out) == golden("fastx_reverse_complement2.out")


def test_golden_stats_old_format(tmp_path):
    recs = H.read_fastx(os.path.join(GOLDEN, "fastq_stats1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    O = H.oracle()
    s = O.fxo_stats_new(64)
    O.fxo_stats_add_batch(s, _p(seq, u8p), _p(qual, u8p), _p(lens, H.i32p), 0, stride, len(recs), 64)
    p = str(tmp_path / "o.txt")
    O.fxo_stats_print_path(s, p.encode(), 0)
    O.fxo_stats_free(s)
    assert open(p, "rb").read() == golden("fastq_stats1.out")


def test_golden_collapser_counts():
    """The fixture predates unordered_map: only the distinct-count ranks are pinned."""
    recs = H.read_fastx(os.path.join(GOLDEN, "fasta_collapser1.fasta"))
    seq, _, lens, stride, _ = H.slab_from_records(recs)
    first, cnt = H.o_collapse(seq, lens, 0, stride)
    exp = H.read_fastx(os.path.join(GOLDEN, "fasta_collapser1.out"))
    exp_cnt = [int(r[0].split(b"-")[1]) for r in exp]
    assert list(cnt) == exp_cnt
    got = [(int(c), recs[int(f)][1]) for f, c in zip(first, cnt)]
    assert sorted(got) == sorted((c, r[1]) for c, r in zip(exp_cnt, exp))
    for k in (0, 1, 2, 3):  # distinct counts => position pinned
        assert got[k][1] == exp[k][1]


# --------------------------------------------------------------------------- vs real binaries

needs_ref = pytest.mark.skipif(H.ref_tool("fastq_quality_trimmer") is None, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("L,kind", [(100, H.PLAIN), (150, H.WITH_N), (37, H.PLAIN)])
def test_ref_trimmer_filter_revcomp(tmp_path, L, kind):
    n = 20000
    seq, qual = H.synth_slab(H.SEED_BASE, n, L, kind)
    stride = seq.shape[1]
    rng = np.random.default_rng(L)
    lens = H.ragged(seq, qual, rng) if L == 37 else None
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, L)
    recs = H.read_fastx(fq)
    ll = lens if lens is not None else np.full(n, L, np.int32)

    out, bad = H.o_trim(seq, qual, lens, L, stride, 33, 20, 20)
    r = H.run([H.ref_tool("fastq_quality_trimmer"), "-Q33", "-t", "20", "-l", "20", "-i", fq])
    assert bad == -1 and emit(recs, out, 33) == r.stdout

    for q, p in ((20, 90), (30, 50), (2, 100), (41, 1)):
        keep, bad = H.o_filter(seq, qual, lens, L, stride, 33, q, p)
        r = H.run([H.ref_tool("fastq_quality_filter"), "-Q33", "-q", str(q), "-p", str(p), "-i", fq])
        assert emit(recs, np.where(keep != 0, ll, -1), 33) == r.stdout

    oseq, oqual = H.o_revcomp(seq, qual, lens, L, stride)
    r = H.run([H.ref_tool("fastx_reverse_complement"), "-Q33", "-i", fq])
    got = b"".join(b"@" + rec[0] + b"\n" + oseq[i, :ll[i]].tobytes() + b"\n+\n" + oqual[i, :ll[i]].tobytes() + b"\n"
                   for i, rec in enumerate(recs))
    assert got == r.stdout


@needs_ref
@pytest.mark.parametrize("new_format", [0, 1])
def test_ref_stats(tmp_path, new_format):
    n, L = 30000, 75
    seq, qual = H.synth_slab(H.SEED_BASE + 3, n, L, H.WITH_N)
    lens = H.ragged(seq, qual, np.random.default_rng(5), min_len=60)
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, L)
    O = H.oracle()
    s = O.fxo_stats_new(L)
    O.fxo_stats_add_batch(s, _p(seq, u8p), _p(qual, u8p), _p(lens, H.i32p), 0, seq.shape[1], n, 33)
    p = str(tmp_path / "o.txt")
    O.fxo_stats_print_path(s, p.encode(), new_format)
    O.fxo_stats_free(s)
    r = H.run([H.ref_tool("fastx_quality_stats"), "-Q33", "-i", fq] + (["-N"] if new_format else []))
    assert open(p, "rb").read() == r.stdout


CLIP_CASES = [
    (b"AGATCGGAAGAGC", ["-l", "20"], dict(min_length=20)),
    (b"AGATCGGAAGAGC", ["-l", "20", "-n"], dict(min_length=20, discard_unknown=0)),
    (b"AGATCGGAAGAGC", ["-l", "5", "-c"], dict(min_length=5, discard_non_clipped=1)),
    (b"AGATCGGAAGAGC", ["-l", "5", "-C", "-n"], dict(min_length=5, discard_clipped=1, discard_unknown=0)),
    (b"AGATCGGAAGAGC", ["-l", "10", "-d", "3", "-n"], dict(min_length=10, keep_delta=3 + 13, discard_unknown=0)),
    (b"AGATCGGAAGAGC", ["-l", "10", "-M", "8"], dict(min_length=10, min_adapter_len=8)),
    (b"AGNTCGGAAGNGCTTGA", ["-l", "12", "-n"], dict(min_length=12, discard_unknown=0)),
    (b"CCTTAAGG", [], dict()),
]


@needs_ref
@pytest.mark.parametrize("case", range(len(CLIP_CASES)))
@pytest.mark.parametrize("kind", [H.ADAPTER, H.WITH_N])
def test_ref_clipper_uniform(tmp_path, case, kind):
    adapter, flags, kw = CLIP_CASES[case]
    n, L = 4000, 60
    seq, qual = H.synth_slab(H.SEED_BASE + 2, n, L, kind)
    if kind == H.WITH_N:  # plant adapters by hand as well
        rng = np.random.default_rng(11)
        for i in rng.choice(n, n // 3, replace=False):
            st = int(rng.integers(0, L))
            m = min(len(adapter), L - st)
            seq[i, st:st + m] = np.frombuffer(adapter[:m], np.uint8)
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, None, L)
    recs = H.read_fastx(fq)
    base = dict(min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    base.update(kw)
    opts = H.FxoClipOpts(**base)
    out_len, cls, cut = H.o_clip(seq, None, None, L, seq.shape[1], adapter, opts)
    r = H.run([H.ref_tool("fastx_clipper"), "-Q33", "-a", adapter.decode(), "-i", fq] + flags)
    assert emit(recs, np.where(cls == 0, out_len, -1), 33) == r.stdout


@needs_ref
def test_ref_clipper_mixed_length_stale_tail(tmp_path):
    """with mixed-length input the reference's DP runs over the widest query
    seen so far and reads NUL + stale bytes of earlier reads.  The oracle reproduces that when fed
    the shadow row + running width."""
    n, L = 3000, 60
    seq, qual = H.synth_slab(H.SEED_BASE + 7, n, L, H.ADAPTER)
    lens = H.ragged(seq, qual, np.random.default_rng(3), min_len=8)
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, L)
    recs = H.read_fastx(fq)
    stride = seq.shape[1]
    rows = np.zeros_like(seq)
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
    for flags, kw in ((["-l", "5", "-C", "-n"], dict(discard_clipped=1, discard_unknown=0)),
                      (["-l", "5", "-c"], dict(discard_non_clipped=1))):
        base = dict(min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
        base.update(kw)
        out_len, cls, cut = H.o_clip(rows, lens, widths, 0, stride, b"AGATCGGAAGAGC", H.FxoClipOpts(**base))
        r = H.run([H.ref_tool("fastx_clipper"), "-Q33", "-a", "AGATCGGAAGAGC", "-i", fq] + flags)
        assert emit(recs, np.where(cls == 0, out_len, -1), 33) == r.stdout


@needs_ref
@pytest.mark.parametrize("n,L", [(300000, 50), (40000, 23)])
def test_ref_collapser_order(tmp_path, n, L):
    seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    fa = str(tmp_path / "in.fa")
    H.write_fasta(fa, seq, None, L)
    first, cnt = H.o_collapse(seq, None, L, seq.shape[1])
    got = b"".join(b">%d-%d\n" % (k + 1, int(c)) + seq[int(f), :L].tobytes() + b"\n"
                   for k, (f, c) in enumerate(zip(first, cnt)))
    r = H.run([H.ref_tool("fastx_collapser"), "-i", fa])
    assert got == r.stdout
    assert 0.5 * n < len(first) < 0.95 * n


def test_hash_bytes_known_answers():
    """std::hash<std::string> values taken from g++ 13.3 (libstdc++), via oracle/_ref probing:
    recomputed independently in pure Python here."""
    def py_hash(b, seed=0xc70f6907):
        M = (1 << 64) - 1
        mul = 0xc6a4a7935bd1e995
        h = (seed ^ (len(b) * mul)) & M
        n8 = len(b) & ~7
        for p in range(0, n8, 8):
            d = (int.from_bytes(b[p:p + 8], "little") * mul) & M
            d ^= d >> 47
            d = (d * mul) & M
            h ^= d
            h = (h * mul) & M
        if len(b) & 7:
            h ^= int.from_bytes(b[n8:], "little")
            h = (h * mul) & M
        h ^= h >> 47
        h = (h * mul) & M
        h ^= h >> 47
        return h
    O = H.oracle()
    for s in (b"", b"A", b"ACGTACG", b"ACGTACGT", b"ACGTACGTA", b"N" * 50, bytes(range(65, 91)) * 3):
        buf = C.create_string_buffer(s, len(s) + 1)
        assert O.fxo_hash_bytes(buf, len(s), 0xc70f6907) == py_hash(s)


def _stale_rows(seq, lens):
    """the query buffer of the reference's aligner after each read (grow-only, keeps stale bytes: SURVEY Appendix D.1)"""
    n, stride = seq.shape
    rows = np.zeros_like(seq)
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
    return rows, widths


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_ref_clipper_fuzz(tmp_path, seed):
    """random adapters (1..25 characters, sometimes with N), reads carrying whole, truncated and mutated copies of them,
    random flag sets incl. -k / -d / -M, equal-length and mixed-length inputs: oracle vs the reference binary
    (fastx_clipper.cpp:159-241,280-319; sequence_alignment.cpp:340-650)"""
    rng = np.random.default_rng(7000 + seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    alen = int(rng.integers(1, 26))
    adapter = alpha[rng.integers(0, 4, alen)].copy()
    if seed % 4 == 3 and alen > 2:
        adapter[rng.integers(0, alen, max(1, alen // 6))] = ord("N")
    adapter = adapter.tobytes()
    n, L = 1500, int(rng.integers(20, 81))
    stride = (L + 1 + 15) // 16 * 16
    seq = np.zeros((n, stride), np.uint8)
    seq[:, :L] = alpha[rng.integers(0, 4, (n, L))]
    qual = np.zeros((n, stride), np.uint8)
    qual[:, :L] = 33 + rng.integers(2, 41, (n, L))
    for i in range(n):
        r = rng.random()
        if r < 0.55:                                      # a copy of the adapter, possibly mutated, possibly cut by the read end
            a = np.frombuffer(adapter, np.uint8).copy()
            for _ in range(int(rng.integers(0, 3))):
                k = int(rng.integers(0, len(a)))
                m = rng.random()
                if m < 0.5: a[k] = alpha[rng.integers(0, 4)]
                elif m < 0.75 and len(a) > 1: a = np.delete(a, k)
                else: a = np.insert(a, k, alpha[rng.integers(0, 4)])
            st = int(rng.integers(0, L))
            m = min(len(a), L - st)
            seq[i, st:st + m] = a[:m]
        if rng.random() < 0.1:
            seq[i, rng.integers(0, L, int(rng.integers(1, 4)))] = ord("N")
    mixed = seed % 2 == 1
    lens = H.ragged(seq, qual, rng, min_len=6) if mixed else None
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, L)
    recs = H.read_fastx(fq)
    for _ in range(3):
        kw = dict(min_length=int(rng.integers(0, 30)), keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
        flags = ["-l", str(kw["min_length"])]
        if rng.random() < 0.4:
            d = int(rng.integers(1, 6)); kw["keep_delta"] = d + len(adapter); flags += ["-d", str(d)]
        c = rng.random()
        if c < 0.25: kw["discard_non_clipped"] = 1; flags.append("-c")
        elif c < 0.5: kw["discard_clipped"] = 1; flags.append("-C")
        if rng.random() < 0.5: kw["discard_unknown"] = 0; flags.append("-n")
        if rng.random() < 0.4:
            kw["min_adapter_len"] = int(rng.integers(1, len(adapter) + 3)); flags += ["-M", str(kw["min_adapter_len"])]
        adapter_only = rng.random() < 0.25
        if adapter_only: flags.append("-k")
        if mixed:
            rows, widths = _stale_rows(seq, lens)
            out_len, cls, cut = H.o_clip(rows, lens, widths, 0, stride, adapter, H.FxoClipOpts(**kw))
            full = lens
        else:
            out_len, cls, cut = H.o_clip(seq, None, None, L, stride, adapter, H.FxoClipOpts(**kw))
            full = np.full(n, L, np.int32)
        r = H.run([H.ref_tool("fastx_clipper"), "-Q33", "-a", adapter.decode(), "-i", fq] + flags)
        exp_len = np.where(cls == 1, full, -1) if adapter_only else np.where(cls == 0, out_len, -1)
        assert emit(recs, exp_len, 33) == r.stdout, (seed, adapter, flags)


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_ref_trim_filter_stats_fuzz(tmp_path, seed):
    """random -Q (33 / 64), qualities over the whole legal range (q = -15 .. 62, fastx.h:28-29), ragged lengths, random
    thresholds: trimmer, filter and both quality-stats formats, oracle vs the reference binaries"""
    rng = np.random.default_rng(9100 + seed)
    Q = 33 if seed % 2 == 0 else 64
    n, L = 4000, int(rng.integers(10, 121))
    stride = (L + 15) // 16 * 16
    alpha = np.frombuffer(b"ACGTN", np.uint8)
    seq = np.zeros((n, stride), np.uint8)
    seq[:, :L] = alpha[rng.choice(5, (n, L), p=[0.24, 0.24, 0.24, 0.24, 0.04])]
    qual = np.zeros((n, stride), np.uint8)
    qual[:, :L] = Q + rng.integers(-15, 63, (n, L))
    qual[:, :L][qual[:, :L] > 126] = 126
    lens = H.ragged(seq, qual, rng, min_len=1)
    fq = str(tmp_path / "in.fq")
    H.write_fastq(fq, seq, qual, lens, L)
    recs = H.read_fastx(fq)
    for _ in range(3):
        t, l = int(rng.integers(1, 46)), int(rng.integers(0, L + 2))
        out, bad = H.o_trim(seq, qual, lens, 0, stride, Q, t, l)
        r = H.run([H.ref_tool("fastq_quality_trimmer"), "-Q%d" % Q, "-t", str(t), "-l", str(l), "-i", fq])
        assert bad == -1 and emit(recs, out, Q) == r.stdout, (seed, t, l)
        q, p = int(rng.integers(-10, 50)), int(rng.integers(1, 101))
        keep, bad = H.o_filter(seq, qual, lens, 0, stride, Q, q, p)
        r = H.run([H.ref_tool("fastq_quality_filter"), "-Q%d" % Q, "-q", str(q), "-p", str(p), "-i", fq])
        assert emit(recs, np.where(keep != 0, lens, -1), Q) == r.stdout, (seed, q, p)
    O = H.oracle()
    for new_format in (0, 1):
        s = O.fxo_stats_new(L)
        O.fxo_stats_add_batch(s, _p(seq, u8p), _p(qual, u8p), _p(lens, H.i32p), 0, stride, n, Q)
        p = str(tmp_path / ("o%d.txt" % new_format))
        O.fxo_stats_print_path(s, p.encode(), new_format)
        O.fxo_stats_free(s)
        r = H.run([H.ref_tool("fastx_quality_stats"), "-Q%d" % Q, "-i", fq] + (["-N"] if new_format else []))
        assert open(p, "rb").read() == r.stdout, (seed, new_format)
