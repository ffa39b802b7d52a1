"""GPU parity (-m gpu): K-STATS histograms and K-CLIP decisions vs the CPU oracle, through the C ABI."""
import os

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import ctx, dev  # noqa: F401  (module-scoped fixture + helper)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def gpu_hist(ctx, seq, qual, lens, L, Q, max_cycles, weight=None, host=False):
    import fastx_toolkit_b200 as F  # noqa: F401
    n, stride = seq.shape
    hist = torch.zeros((max_cycles, 5, 109), dtype=torch.int64, device="cuda")
    ctx.report_reset()
    if host:
        torch.cuda.synchronize()   # the *_host calls run on the library's own side streams
        rep = ctx.stats_accum_host(ctx.batch(seq, qual, n, stride, L, lens), Q, hist, max_cycles, weight)
    else:
        keep = [dev(seq), dev(qual), dev(lens), dev(weight)]   # device copies must outlive the launch
        ctx.stats_accum_dev(ctx.batch(keep[0], keep[1], n, stride, L, keep[2]), Q, hist, max_cycles, keep[3])
        rep = ctx.sync()
    return hist.cpu().numpy().astype(np.uint64), rep


@pytest.mark.parametrize("L,n,kind", [(150, 50001, H.WITH_N), (100, 20000, H.PLAIN), (36, 9000, H.WITH_N), (250, 6000, H.WITH_N),
                                      (1000, 700, H.WITH_N), (3, 500, H.PLAIN), (161, 3000, H.WITH_N)])
def test_stats_hist_uniform(ctx, L, n, kind):
    seq, qual = H.synth_slab(H.SEED_BASE + 3, n, L, kind)
    exp, cyc = H.o_stats_hist(seq, qual, None, L, seq.shape[1], 33, L)
    got, rep = gpu_hist(ctx, seq, qual, None, L, 33, L)
    assert cyc == L and rep.first_bad_read == -1 and rep.n_in == n
    assert np.array_equal(got, exp)
    assert int(got.sum()) == n * L


def test_stats_full_quality_range_and_ragged(ctx):
    n, L = 30000, 150
    seq, qual = H.synth_slab(H.SEED_BASE + 3, n, L, H.WITH_N)
    rng = np.random.default_rng(7)
    for Q in (33, 64):
        lo, hi = Q - 15, min(Q + 93, 127)
        qual2 = np.zeros_like(qual)
        qual2[:, :L] = rng.integers(lo, hi + 1, size=(n, L), dtype=np.uint8)
        seq2 = seq.copy()
        lens = H.ragged(seq2, qual2, rng, min_len=1)
        exp, cyc = H.o_stats_hist(seq2, qual2, lens, 0, seq.shape[1], Q, L)
        got, rep = gpu_hist(ctx, seq2, qual2, lens, 0, Q, L)
        assert rep.first_bad_read == -1 and np.array_equal(got, exp)
        got2, rep2 = gpu_hist(ctx, seq2, qual2, lens, 0, Q, L, host=True)
        assert np.array_equal(got2, exp) and rep2.n_in == n
        # max_cycles smaller than the reads: cycles beyond it are dropped
        got3, _ = gpu_hist(ctx, seq2, qual2, lens, 0, Q, 40)
        assert np.array_equal(got3, exp[:40])


def test_stats_fasta_weights_and_bad_read(ctx):
    n, L = 5000, 50
    seq, qual = H.synth_slab(H.SEED_BASE + 4, n, L, H.WITH_N)
    w = np.random.default_rng(2).integers(1, 9, size=n).astype(np.int32)
    got, rep = gpu_hist(ctx, seq, None, None, L, 33, L, weight=w)
    exp = np.zeros((L, 5, 109), np.uint64)
    for k, ch in enumerate(b"ACGTN"):
        exp[:, k, 15] = ((seq[:, :L] == ch) * w[:, None]).sum(axis=0)
    assert np.array_equal(got, exp)
    s2 = seq.copy(); s2[4000, 17] = ord("X")
    q2 = qual.copy(); q2[1234, 49] = 5
    _, rep = gpu_hist(ctx, s2, q2, None, L, 33, L)
    assert rep.first_bad_read == 1234


def test_stats_golden_fixture_hist(ctx):
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastq_stats1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    exp, cyc = H.o_stats_hist(seq, qual, lens, 0, stride, 64, 36)
    got, rep = gpu_hist(ctx, seq, qual, lens, 0, 64, 36)
    assert cyc == 36 and np.array_equal(got, exp) and rep.first_bad_read == -1


# ------------------------------------------------------------------------------------------ clipper

def gpu_clip(ctx, seq, qual, lens, widths, L, adapter, opts_kw, host=False):
    import fastx_toolkit_b200 as F
    n, stride = seq.shape
    base = dict(min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    base.update(opts_kw)
    o = F.ClipOpts(adapter=adapter, **base)
    oo = H.FxoClipOpts(**base)
    exp = H.o_clip(seq, lens, widths, L, stride, adapter, oo)
    ctx.report_reset()
    if host:
        out_len, out_cls = np.empty(n, np.int32), np.empty(n, np.uint8)
        rep = ctx.clip_host(ctx.batch(seq, qual, n, stride, L, lens), widths, 33, o, out_len, out_cls)
        out_cut = None
    else:
        d_len = torch.empty(n, dtype=torch.int32, device="cuda")
        d_cls = torch.empty(n, dtype=torch.uint8, device="cuda")
        d_cut = torch.empty(n, dtype=torch.int32, device="cuda")
        keep = [dev(seq), dev(qual), dev(lens), dev(widths)]
        ctx.clip_dev(ctx.batch(keep[0], keep[1], n, stride, L, keep[2]), keep[3], 33, o, d_len, d_cls, d_cut)
        rep = ctx.sync()
        out_len, out_cls, out_cut = d_len.cpu().numpy(), d_cls.cpu().numpy(), d_cut.cpu().numpy()
    e_len, e_cls, e_cut = exp
    assert np.array_equal(out_cls, e_cls)
    if out_cut is not None:
        assert np.array_equal(out_cut, e_cut)
    assert np.array_equal(out_len, np.where(e_cls == 0, e_len, -1))
    assert rep.n_out == int((e_cls == 0).sum())
    for c in range(1, 6):
        assert rep.aux[c] == int((e_cls == c).sum())
    return rep


CASES = [
    (b"AGATCGGAAGAGC", dict(min_length=20)),
    (b"AGATCGGAAGAGC", dict(min_length=20, discard_unknown=0)),
    (b"AGATCGGAAGAGC", dict(min_length=5, discard_non_clipped=1)),
    (b"AGATCGGAAGAGC", dict(min_length=5, discard_clipped=1, discard_unknown=0)),
    (b"AGATCGGAAGAGC", dict(min_length=10, keep_delta=3 + 13, discard_unknown=0)),
    (b"AGATCGGAAGAGC", dict(min_length=10, min_adapter_len=8)),
    (b"AGNTCGGAAGNGCTTGA", dict(min_length=12, discard_unknown=0)),
    (b"CCTTAAGG", dict()),
    (b"CAATTGGTTAATCCCCCTATATA", dict(min_length=15, discard_unknown=0, discard_non_clipped=1)),
    (b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTATGCC", dict(min_length=10)),   # 52 nt: local-memory column path
    (b"A", dict(min_length=1)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_clip_uniform(ctx, case):
    adapter, kw = CASES[case]
    for L, n, kind in ((60, 6000, H.ADAPTER), (150, 4000, H.ADAPTER), (33, 3000, H.WITH_N)):
        seq, qual = H.synth_slab(H.SEED_BASE + 2, n, L, kind)
        rng = np.random.default_rng(case * 7 + L)
        for i in rng.choice(n, n // 3, replace=False):   # plant (possibly truncated / mutated) adapters
            st = int(rng.integers(0, L))
            m = min(len(adapter), L - st)
            ad = np.frombuffer(adapter[:m], np.uint8).copy()
            if m > 4 and rng.random() < 0.4:
                ad[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
            seq[i, st:st + m] = ad
        gpu_clip(ctx, seq, qual, None, None, L, adapter, kw)


def test_clip_tiny_and_host_and_fasta(ctx):
    for L in (1, 2, 3, 5):
        seq, qual = H.synth_slab(H.SEED_BASE + 8, 2000, L, H.WITH_N)
        gpu_clip(ctx, seq, qual, None, None, L, b"AGATCGGAAGAGC", dict(min_length=1, discard_unknown=0))
        gpu_clip(ctx, seq, None, None, None, L, b"CCTTAAGG", dict(min_length=0))
    seq, qual = H.synth_slab(H.SEED_BASE + 2, 300000, 150, H.ADAPTER)
    gpu_clip(ctx, seq, qual, None, None, 150, b"AGATCGGAAGAGC", dict(min_length=20), host=True)
    s2 = seq[:5000].copy(); q2 = qual[:5000].copy()
    s2[77, 3] = ord("n"); q2[12, 100] = 3
    import fastx_toolkit_b200 as F
    o = F.ClipOpts(adapter=b"AGATCGGAAGAGC", min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0,
                   discard_unknown=1, min_adapter_len=0)
    d_len = torch.empty(5000, dtype=torch.int32, device="cuda")
    ctx.report_reset()
    ds2, dq2 = dev(s2), dev(q2)
    ctx.clip_dev(ctx.batch(ds2, dq2, 5000, 160, 150), None, 33, o, d_len)
    assert ctx.sync().first_bad_read == 12


def test_clip_mixed_length_stale_tail(ctx):
    """Bug-compatible mode: rows carry NUL + stale bytes, widths = running max."""
    n, L = 5000, 60
    seq, qual = H.synth_slab(H.SEED_BASE + 7, n, L, H.ADAPTER)
    lens = H.ragged(seq, qual, np.random.default_rng(3), min_len=8)
    stride = seq.shape[1]
    rows = np.zeros_like(seq)
    widths = np.zeros(n, np.int32)
    shadow = np.zeros(stride + 1, np.uint8)
    wmax = 0
    for i in range(n):
        l = int(lens[i])
        shadow[:l] = seq[i, :l]; shadow[l] = 0
        wmax = max(wmax, l)
        rows[i, :wmax] = shadow[:wmax]
        widths[i] = wmax
    gpu_clip(ctx, rows, qual, lens, widths, 0, b"AGATCGGAAGAGC", dict(min_length=5, discard_clipped=1, discard_unknown=0))
    gpu_clip(ctx, rows, qual, lens, widths, 0, b"AGATCGGAAGAGC", dict(min_length=5, discard_non_clipped=1))
    gpu_clip(ctx, seq, qual, lens, None, 0, b"AGATCGGAAGAGC", dict(min_length=5))     # clean mode: width = len


def test_clip_golden_fixture(ctx):
    from test_oracle_golden import emit, golden
    import fastx_toolkit_b200 as F
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastx_clipper1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    o = F.ClipOpts(adapter=b"CAATTGGTTAATCCCCCTATATA", min_length=15, keep_delta=0, discard_non_clipped=1, discard_clipped=0,
                   discard_unknown=0, min_adapter_len=0)
    n = len(recs)
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.report_reset()
    keep = [dev(seq), dev(qual), dev(lens)]
    ctx.clip_dev(ctx.batch(keep[0], keep[1], n, stride, 0, keep[2]), None, 64, o, d_len)
    rep = ctx.sync()
    assert rep.first_bad_read == -1
    assert emit(recs, d_len.cpu().numpy(), 64) == golden("fastx_clipper1a.out")


# ------------------------------------------------------------------- k_stats4 geometry / fallback coverage

@pytest.mark.parametrize("L", [1, 2, 4, 5, 31, 32, 33, 63, 64, 65, 127, 128, 129, 131, 159, 160, 200, 320, 321])
def test_stats_length_sweep(ctx, L):
    """k_stats4: A region (8 chunks, fewer for short reads), unmasked and masked B region, tail bytes, multi-pass (> 160 cycles)."""
    n = 4099
    seq, qual = H.synth_slab(H.SEED_BASE + 11, n, L, H.WITH_N)
    exp, cyc = H.o_stats_hist(seq, qual, None, L, seq.shape[1], 33, L)
    got, rep = gpu_hist(ctx, seq, qual, None, L, 33, L)
    assert cyc == L and rep.first_bad_read == -1 and np.array_equal(got, exp)
    # ragged lengths over the same rows, junk in the padding
    rng = np.random.default_rng(L)
    seq2, qual2 = seq.copy(), qual.copy()
    lens = H.ragged(seq2, qual2, rng, min_len=1)
    exp2, _ = H.o_stats_hist(seq2, qual2, lens, 0, seq.shape[1], 33, L)
    got2, rep2 = gpu_hist(ctx, seq2, qual2, lens, 0, 33, L)
    assert rep2.first_bad_read == -1 and np.array_equal(got2, exp2)


def test_stats_full_range_bad_reads_and_fallback_kernel(ctx):
    """the whole legal quality range (q' >= 64 leaves the shared histogram for the global table), bad reads at the corners of the
    layout, and the inputs only the global-atomics fallback takes (-Q 90)."""
    n, L = 40001, 150
    seq, qual = H.synth_slab(H.SEED_BASE + 12, n, L, H.WITH_N)
    rng = np.random.default_rng(5)
    qual[:, :L] = rng.integers(33 - 15, 127, size=(n, L), dtype=np.uint8)     # full legal range: q' >= 64 goes to the global table
    exp, _ = H.o_stats_hist(seq, qual, None, L, seq.shape[1], 33, L)
    got2, _ = gpu_hist(ctx, seq, qual, None, L, 33, L)
    assert np.array_equal(got2, exp)
    for pos, (arr, val) in enumerate([(seq, ord("a")), (qual, 17), (seq, 0), (qual, 200), (seq, ord("X"))]):
        a2 = arr.copy()
        r, c = 1000 + 7 * pos, [0, 37, 148, 149, 75][pos]
        a2[r, c] = val
        s2, q2 = (a2, qual) if arr is seq else (seq, a2)
        _, rep = gpu_hist(ctx, s2, q2, None, L, 33, L)
        assert rep.first_bad_read == r
    # Q = 64 and an offset too large for the packed range test (the global-atomics kernel)
    for Q in (64, 90):
        q3 = np.zeros_like(qual)
        q3[:, :L] = rng.integers(Q - 15, min(Q + 93, 127) + 1, size=(n, L), dtype=np.uint8)
        exp3, _ = H.o_stats_hist(seq, q3, None, L, seq.shape[1], Q, L)
        got3, rep3 = gpu_hist(ctx, seq, q3, None, L, Q, L)
        assert rep3.first_bad_read == -1 and np.array_equal(got3, exp3)


@pytest.mark.parametrize("adapter", [b"A", b"ACGT", b"CCTTA", b"CCTTAAGG", b"TGGAATTCTCGG", b"AGATCGGAAGAGC", b"AGATCGGAAGAGCACA"])
def test_clip_dpx_sweep(ctx, adapter):
    """Integer (packed s16x2) clipper path: every adapter-length bucket, odd batch sizes, reads with N (second pass)."""
    for L, n, kind in ((150, 5001, H.ADAPTER), (36, 3333, H.WITH_N), (256, 801, H.WITH_N), (255, 500, H.PLAIN), (1, 77, H.PLAIN),
                       (17, 1, H.PLAIN)):
        seq, qual = H.synth_slab(H.SEED_BASE + 13, n, L, kind)
        rng = np.random.default_rng(len(adapter) * 1000 + L)
        for i in rng.choice(n, max(1, n // 2), replace=False):
            st = int(rng.integers(0, L))
            m = min(len(adapter), L - st)
            ad = np.frombuffer(adapter[:m], np.uint8).copy()
            if m > 2 and rng.random() < 0.5:
                ad[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
            if m > 5 and rng.random() < 0.3:          # a deletion inside the adapter copy
                cut = int(rng.integers(1, m - 1))
                ad = np.concatenate([ad[:cut], ad[cut + 1:]])
            seq[i, st:st + len(ad)] = ad
        for kw in (dict(min_length=5), dict(min_length=1, discard_unknown=0, discard_clipped=1)):
            gpu_clip(ctx, seq, qual, None, None, L, adapter, kw)


def test_clip_generations_agree(ctx, monkeypatch):
    """FXG_CLIP_V=1 (fp32 kernel for everything) equals the default (integer path + fp32 second pass)."""
    n, L = 20001, 100
    seq, qual = H.synth_slab(H.SEED_BASE + 14, n, L, H.ADAPTER)
    seq[::7, 50] = ord("N")
    monkeypatch.setenv("FXG_CLIP_V", "1")
    gpu_clip(ctx, seq, qual, None, None, L, b"AGATCGGAAGAGC", dict(min_length=20))
    monkeypatch.delenv("FXG_CLIP_V")
    gpu_clip(ctx, seq, qual, None, None, L, b"AGATCGGAAGAGC", dict(min_length=20))
    gpu_clip(ctx, seq, None, None, None, L, b"AGATCGGAAGAGC", dict(min_length=20, discard_unknown=0))     # FASTA


def test_clip_config_c_100m(ctx):
    """BASELINE config (c) at full size: 100 M x 150 bp, 30 % of the reads carry the adapter, fastx_clipper -a AGATCGGAAGAGC -l 20.
    Prefix and tail equal the oracle; every read lands in exactly one class; lengths stay in range."""
    import fastx_toolkit_b200 as F
    n = int(os.environ.get("FXG_FULL_N", 100_000_000))
    L, stride = 150, 160
    free = torch.cuda.mem_get_info()[0]
    if free < (n * (stride * 2 + 16)) * 1.05:
        n = int(free / 1.05 / (stride * 2 + 16)) // 1024 * 1024
    dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    dqual = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dqual, n, L, stride, H.SEED_BASE + 2, H.ADAPTER, 33)
    kw = dict(min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    o = F.ClipOpts(adapter=b"AGATCGGAAGAGC", **kw)
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_cls = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.report_reset()
    ctx.clip_dev(ctx.batch(dseq, dqual, n, stride, L), None, 33, o, d_len, d_cls, None)
    rep = ctx.sync()
    assert rep.first_bad_read == -1 and rep.n_in == n
    assert rep.n_out + sum(rep.aux[c] for c in range(1, 6)) == n
    hist = torch.bincount(d_cls.to(torch.int64), minlength=6).cpu().numpy()
    assert hist[0] == rep.n_out and all(hist[c] == rep.aux[c] for c in range(1, 6))
    kept = d_len[d_cls == 0]
    assert int(kept.min().item()) >= 20 and int(kept.max().item()) <= L and bool((d_len[d_cls != 0] == -1).all().item())
    assert 0.25 * n < n - int((d_len == L).sum().item()) < 0.8 * n          # the lenient rules clip many adapter-free tails too
    oo = H.FxoClipOpts(**kw)
    for first, m in ((0, 100000), (n - 20000, 20000)):
        seq, _ = H.synth_slab(H.SEED_BASE + 2, m, L, H.ADAPTER, first=first)
        e_len, e_cls, _ = H.o_clip(seq, None, None, L, stride, b"AGATCGGAAGAGC", oo)
        assert np.array_equal(d_cls[first:first + m].cpu().numpy(), e_cls)
        assert np.array_equal(d_len[first:first + m].cpu().numpy(), np.where(e_cls == 0, e_len, -1))


def test_stats_config_d_share_100m(ctx):
    """BASELINE config (d), one GPU's worth and more: fastx_quality_stats over 100 M x 150 bp.  hist.sum() == n*L, every cycle
    holds n bases, the histogram of a 500 K-read prefix equals the oracle, and the whole equals the sum of its two halves
    (the additivity the multi-GPU all-reduce relies on)."""
    n = int(os.environ.get("FXG_FULL_N", 100_000_000))
    L, stride = 150, 160
    free = torch.cuda.mem_get_info()[0]
    if free < (n * stride * 2) * 1.05:
        n = int(free / 1.05 / (stride * 2)) // 1024 * 1024
    dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    dqual = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dqual, n, L, stride, H.SEED_BASE + 3, H.WITH_N, 33)
    whole = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.report_reset()
    ctx.stats_accum_dev(ctx.batch(dseq, dqual, n, stride, L), 33, whole, L)
    rep = ctx.sync()
    assert rep.first_bad_read == -1
    assert int(whole.sum().item()) == n * L and bool((whole.sum(dim=(1, 2)) == n).all().item())
    h = n // 2 + 12345
    halves = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.stats_accum_dev(ctx.batch(dseq, dqual, h, stride, L), 33, halves, L)
    ctx.stats_accum_dev(ctx.batch(dseq[h:], dqual[h:], n - h, stride, L), 33, halves, L, None, h)
    ctx.sync()
    assert torch.equal(whole, halves)
    # 12 M reads: every CTA goes through more than 65 280 reads, i.e. through a flush of its 16-bit counters; checked against an
    # independent count made with library ops (torch.bincount), cycle range by cycle range
    k12 = min(n, 12_000_000)
    h12 = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.stats_accum_dev(ctx.batch(dseq, dqual, k12, stride, L), 33, h12, L)
    ctx.sync()
    lut = torch.full((256,), 4, dtype=torch.int64, device="cuda")
    for ch, v in ((ord("A"), 0), (ord("C"), 1), (ord("G"), 2), (ord("T"), 3), (ord("N"), 4)):
        lut[ch] = v
    for c0 in range(0, L, 25):
        c1 = min(L, c0 + 25)
        nuc = lut[dseq[:k12, c0:c1].long()]
        qp = dqual[:k12, c0:c1].long() - (33 - 15)
        cyc = torch.arange(c0, c1, device="cuda").view(1, -1).expand(k12, -1)
        ref = torch.bincount(((cyc * 5 + nuc) * 109 + qp).reshape(-1), minlength=L * 5 * 109).view(L, 5, 109)
        assert torch.equal(ref[c0:c1], h12[c0:c1]), (c0, c1)
        del nuc, qp, cyc, ref
    m = 500000
    pre = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.stats_accum_dev(ctx.batch(dseq, dqual, m, stride, L), 33, pre, L)
    ctx.sync()
    seq, qual = H.synth_slab(H.SEED_BASE + 3, m, L, H.WITH_N)
    eh, _ = H.o_stats_hist(seq, qual, None, L, stride, 33, L)
    assert np.array_equal(pre.cpu().numpy().astype(np.uint64), eh)
