"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(libfxg.so via ctypes), must be bit-identical to the CPU oracle on the same seeded inputs, to the
reference's golden fixtures, and must satisfy size-independent properties at BASELINE.json sizes.
"""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    import fastx_toolkit_b200 as F
    c = F.Context(0)
    # run the library on torch's current stream so that torch allocations/fills and our kernels are ordered
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_trim(ctx, seq, qual, lens, L, Q, t, l, with_seq=True):
    n, stride = qual.shape
    dseq, dqual, dlens = dev(seq) if with_seq else None, dev(qual), dev(lens)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.report_reset()
    ctx.trim_dev(ctx.batch(dseq, dqual, n, stride, L, dlens), Q, t, l, out)
    rep = ctx.sync()
    return out.cpu().numpy(), rep


def run_filter(ctx, seq, qual, lens, L, Q, q, p, with_seq=True):
    n, stride = qual.shape
    dseq, dqual, dlens = dev(seq) if with_seq else None, dev(qual), dev(lens)
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.report_reset()
    ctx.filter_dev(ctx.batch(dseq, dqual, n, stride, L, dlens), Q, q, p, out)
    rep = ctx.sync()
    return out.cpu().numpy(), rep


def run_revcomp(ctx, seq, qual, lens, L, Q=33):
    n, stride = seq.shape
    dseq, dqual, dlens = dev(seq), dev(qual), dev(lens)
    oseq = torch.full((n, stride), 0xEE, dtype=torch.uint8, device="cuda")
    oqual = torch.full((n, stride), 0xEE, dtype=torch.uint8, device="cuda") if qual is not None else None
    ctx.report_reset()
    ctx.revcomp_dev(ctx.batch(dseq, dqual, n, stride, L, dlens), Q, oseq, oqual)
    rep = ctx.sync()
    return oseq.cpu().numpy(), None if oqual is None else oqual.cpu().numpy(), rep


GEOMS = [(100, 20011), (150, 30001), (50, 10007), (128, 5003), (36, 7001), (250, 4099), (1000, 1203), (16, 999), (1, 300), (2500, 97)]


@pytest.mark.parametrize("L,n", GEOMS)
def test_trim_filter_uniform(ctx, L, n):
    seq, qual = H.synth_slab(H.SEED_BASE, n, L, H.WITH_N)
    stride = seq.shape[1]
    for t, l in ((20, 20), (30, 0), (2, 1), (41, 5), (-100, 0), (200, 0)):
        exp, bad = H.o_trim(seq, qual, None, L, stride, 33, t, l)
        got, rep = run_trim(ctx, seq, qual, None, L, 33, t, l)
        assert np.array_equal(got, exp), (L, t, l)
        assert rep.first_bad_read == bad == -1
        assert rep.n_out == int((exp >= 0).sum()) and rep.n_in == n
    for q, p in ((20, 90), (30, 50), (2, 100), (41, 1), (25, 0), (94, 0), (-20, 100)):
        exp, bad = H.o_filter(seq, qual, None, L, stride, 33, q, p)
        got, rep = run_filter(ctx, seq, qual, None, L, 33, q, p)
        assert np.array_equal(got, exp), (L, q, p)
        assert rep.n_out == int(exp.sum())


@pytest.mark.parametrize("L,n", [(150, 20000), (100, 9999), (75, 5000), (400, 3000)])
def test_trim_filter_revcomp_ragged(ctx, L, n):
    seq, qual = H.synth_slab(H.SEED_BASE + 1, n, L, H.WITH_N)
    stride = seq.shape[1]
    lens = H.ragged(seq, qual, np.random.default_rng(L))
    # padding must be ignored: fill it with junk
    junk = np.random.default_rng(1).integers(0, 256, size=seq.shape, dtype=np.uint8)
    pad = np.arange(stride)[None, :] >= lens[:, None]
    seqj, qualj = np.where(pad, junk, seq), np.where(pad, junk, qual)
    exp, _ = H.o_trim(seq, qual, lens, 0, stride, 33, 20, 20)
    got, rep = run_trim(ctx, seqj, qualj, lens, 0, 33, 20, 20)
    assert np.array_equal(got, exp) and rep.first_bad_read == -1
    exp, _ = H.o_filter(seq, qual, lens, 0, stride, 33, 20, 90)
    got, rep = run_filter(ctx, seqj, qualj, lens, 0, 33, 20, 90)
    assert np.array_equal(got, exp) and rep.first_bad_read == -1
    eseq, equal = H.o_revcomp(seq, qual, lens, 0, stride)
    gseq, gqual, rep = run_revcomp(ctx, seqj, qualj, lens, 0)
    assert np.array_equal(gseq, eseq) and np.array_equal(gqual, equal) and rep.first_bad_read == -1


@pytest.mark.parametrize("L,n", GEOMS)
def test_revcomp_uniform(ctx, L, n):
    seq, qual = H.synth_slab(H.SEED_BASE + 5, n, L, H.WITH_N)
    eseq, equal = H.o_revcomp(seq, qual, None, L, seq.shape[1])
    gseq, gqual, rep = run_revcomp(ctx, seq, qual, None, L)
    assert np.array_equal(gseq, eseq) and np.array_equal(gqual, equal)
    assert rep.first_bad_read == -1 and rep.n_in == n
    # FASTA form (no qualities)
    gseq2, _, _ = run_revcomp(ctx, seq, None, None, L)
    assert np.array_equal(gseq2, eseq)
    # involution
    back, backq, _ = run_revcomp(ctx, gseq, gqual, None, L)
    assert np.array_equal(back, seq) and np.array_equal(backq, qual)


def test_every_stride_bucket(ctx):
    """every 16-byte stride from 16 to 512 (plus a few longer): each picks its own tile / shared-memory plan, several of
    which sit exactly on the 48 KB opt-in boundary (stride 96 with qualities needs 48 KB + the barrier words)"""
    strides = list(range(16, 513, 16)) + [640, 768, 1024, 1536, 2048]
    rng = np.random.default_rng(3)
    for S in strides:
        L = S - int(rng.integers(0, 16))
        n = 1500 + int(rng.integers(0, 200))
        seq, qual = H.synth_slab(H.SEED_BASE + S, n, L, H.WITH_N)
        assert seq.shape[1] == S
        for pass_ in range(2):
            ln, LL = (None, L) if pass_ == 0 else (H.ragged(seq, qual, np.random.default_rng(S), min_len=1), 0)
            exp, _ = H.o_trim(seq, qual, ln, LL, S, 33, 22, 10)
            got, rep = run_trim(ctx, seq, qual, ln, LL, 33, 22, 10)
            assert np.array_equal(got, exp), ("trim", S, ln is None)
            exp, _ = H.o_filter(seq, qual, ln, LL, S, 33, 22, 70)
            got, rep = run_filter(ctx, seq, qual, ln, LL, 33, 22, 70)
            assert np.array_equal(got, exp), ("filter", S, ln is None)
            eseq, equal = H.o_revcomp(seq, qual, ln, LL, S)
            gseq, gqual, rep = run_revcomp(ctx, seq, qual, ln, LL)
            assert np.array_equal(gseq, eseq) and np.array_equal(gqual, equal), ("revcomp", S, ln is None)
            gseq, _, rep = run_revcomp(ctx, seq, None, ln, LL)
            assert np.array_equal(gseq, eseq), ("revcomp fasta", S, ln is None)


def test_decide_only_variant(ctx):
    seq, qual = H.synth_slab(H.SEED_BASE, 12345, 150)
    exp, _ = H.o_trim(None, qual, None, 150, 160, 33, 20, 20)
    got, rep = run_trim(ctx, None, qual, None, 150, 33, 20, 20, with_seq=False)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("Q", [33, 64])
def test_validation_first_bad_read(ctx, Q):
    n, L = 50000, 150
    seq, qual = H.synth_slab(H.SEED_BASE + 9, n, L, H.WITH_N, q_offset=Q)
    stride = seq.shape[1]
    rng = np.random.default_rng(Q)
    cases = []
    # illegal bases / qualities at assorted positions (first, last, interior, tail chunk)
    for bad_byte, slab in ((ord("a"), "s"), (ord("E"), "s"), (0, "s"), (200, "s"), (ord("U"), "s"),
                           (Q - 16, "q"), (Q + 94 if Q + 94 < 256 else 255, "q"), (250, "q"), (10, "q")):
        for pos in (0, L - 1, 77, 144, 15, 16):
            cases.append((bad_byte, slab, pos))
    for bad_byte, slab, pos in cases:
        i = int(rng.integers(0, n))
        s2, q2 = seq.copy(), qual.copy()
        (s2 if slab == "s" else q2)[i, pos] = bad_byte
        exp_bad = H.o_trim(s2, q2, None, L, stride, Q, 20, 20)[1]
        if Q == 64 and slab == "q" and bad_byte == Q + 94:  # 158 > 127: a negative char
            assert exp_bad == i
        _, rep = run_trim(ctx, s2, q2, None, L, Q, 20, 20)
        assert rep.first_bad_read == exp_bad, (bad_byte, slab, pos, i)
        _, rep = run_filter(ctx, s2, q2, None, L, Q, 20, 90)
        assert rep.first_bad_read == exp_bad
        _, _, rep = run_revcomp(ctx, s2, q2, None, L, Q)
        assert rep.first_bad_read == exp_bad
    # two bad reads: the smaller index wins; legal extremes are accepted
    s2, q2 = seq.copy(), qual.copy()
    s2[40000, 3] = ord("x"); q2[123, 9] = 255
    q2[7, 0] = Q - 15; q2[7, 1] = min(Q + 93, 127)
    _, rep = run_trim(ctx, s2, q2, None, L, Q, 20, 20)
    assert rep.first_bad_read == 123
    # zero-length read is fatal (fastx.c:361-362)
    lens = np.full(n, L, np.int32); lens[4321] = 0
    _, rep = run_trim(ctx, seq, qual, lens, 0, Q, 20, 20)
    assert rep.first_bad_read == 4321


def test_golden_fixtures_on_gpu(ctx):
    from test_oracle_golden import emit, golden
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastq_quality_trimmer.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    got, rep = run_trim(ctx, seq, qual, lens, 0, 64, 30, 16)
    assert emit(recs, got, 64) == golden("fastq_quality_trimmer.out")
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastq_qual_filter1.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    for q, p, name in ((33, 100, "fastq_qual_filter1a.out"), (20, 80, "fastq_qual_filter1b.out")):
        keep, _ = run_filter(ctx, seq, qual, lens, 0, 64, q, p)
        assert emit(recs, np.where(keep != 0, lens, -1), 64) == golden(name)
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastx_rev_comp1.fasta"))
    seq, _, lens, stride, _ = H.slab_from_records(recs)
    oseq, _, _ = run_revcomp(ctx, seq, None, lens, 0)
    out = b"".join(b">" + r[0] + b"\n" + oseq[i, :lens[i]].tobytes() + b"\n" for i, r in enumerate(recs))
    assert out == golden("fastx_reverse_complement1.out")
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fastx_rev_comp2.fastq"))
    seq, qual, lens, stride, _ = H.slab_from_records(recs, 64)
    oseq, oqual, rep = run_revcomp(ctx, seq, qual, lens, 0, 64)
    assert rep.first_bad_read == -1
    out = []
    for i, r in enumerate(recs):
        l = lens[i]
        nums = b" ".join(b"%d" % (int(np.int8(v)) - 64) for v in oqual[i, :l])
        out.append(b"@" + r[0] + b"\n" + oseq[i, :l].tobytes() + b"\n+" + r[2] + b"\n" + nums + b"\n")
    assert b"".join(out) == golden("fastx_reverse_complement2.out")


def test_synth_dev_matches_numpy(ctx):
    for kind, L in ((H.PLAIN, 150), (H.WITH_N, 100), (H.ADAPTER, 150), (H.DUPS, 50)):
        n = 4096 + 17
        stride = ((L + 15) // 16) * 16
        dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
        dqual = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
        ctx.synth_dev(dseq, dqual, n, L, stride, H.SEED_BASE + kind, kind, 33, first_read=1000, n_total=10**6)
        ctx.sync()
        seq, qual = H.synth_slab(H.SEED_BASE + kind, n, L, kind, first=1000, n_total=10**6)
        assert np.array_equal(dseq.cpu().numpy(), seq) and np.array_equal(dqual.cpu().numpy(), qual)


def test_host_pipeline_matches_device(ctx):
    n, L = 700001, 150
    seq, qual = H.synth_slab(H.SEED_BASE, n, L)
    stride = seq.shape[1]
    pseq, pqual = torch.from_numpy(seq).pin_memory(), torch.from_numpy(qual).pin_memory()
    out = torch.empty(n, dtype=torch.int32).pin_memory()
    rep = ctx.trim_host(ctx.batch(pseq, pqual, n, stride, L), 33, 20, 20, out)
    exp, _ = H.o_trim(seq, qual, None, L, stride, 33, 20, 20)
    assert np.array_equal(out.numpy(), exp) and rep.n_out == int((exp >= 0).sum()) and rep.first_bad_read == -1
    keep = torch.empty(n, dtype=torch.uint8).pin_memory()
    rep = ctx.filter_host(ctx.batch(pseq, pqual, n, stride, L), 33, 20, 90, keep)
    exp, _ = H.o_filter(seq, qual, None, L, stride, 33, 20, 90)
    assert np.array_equal(keep.numpy(), exp)
    oseq, oqual = torch.empty_like(pseq).pin_memory(), torch.empty_like(pqual).pin_memory()
    rep = ctx.revcomp_host(ctx.batch(pseq, pqual, n, stride, L), 33, oseq, oqual)
    eseq, equal = H.o_revcomp(seq, qual, None, L, stride)
    assert np.array_equal(oseq.numpy(), eseq) and np.array_equal(oqual.numpy(), equal)
    # ragged lengths + a bad read far into the batch, unpinned memory
    lens = H.ragged(seq, qual, np.random.default_rng(0), min_len=30)
    qual[650000, 3] = 7
    out2 = np.empty(n, np.int32)
    rep = ctx.trim_host(ctx.batch(seq, qual, n, stride, 0, lens), 33, 20, 20, out2)
    exp, bad = H.o_trim(seq, qual, lens, 0, stride, 33, 20, 20)
    assert rep.first_bad_read == bad == 650000
    ok = np.arange(n) != 650000
    assert np.array_equal(out2[ok], exp[ok])


def test_full_size_properties(ctx):
    """BASELINE.json configs[1]-sized slab (100 M x 150 bp, generated on the device):
    prefix equals the oracle; kept-count equals the number of non-negative results; revcomp is an
    involution; filter with p=100/q=min keeps everything."""
    n = int(os.environ.get("FXG_FULL_N", 100_000_000))
    L, stride = 150, 160
    free = torch.cuda.mem_get_info()[0]
    need = n * stride * 2 + n * 8
    if free < need * 1.05:
        n = int(free / 1.05 / (stride * 2 + 8)) // 1024 * 1024
    dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    dqual = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dqual, n, L, stride, H.SEED_BASE + 1, H.PLAIN, 33)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.report_reset()
    ctx.trim_dev(ctx.batch(dseq, dqual, n, stride, L), 33, 20, 20, out)
    rep = ctx.sync()
    assert rep.first_bad_read == -1 and rep.n_in == n
    assert rep.n_out == int((out >= 0).sum().item())
    assert int(out.max().item()) <= L and int(out[out >= 0].min().item()) >= 20
    m = 200000
    seq, qual = H.synth_slab(H.SEED_BASE + 1, m, L)
    exp, _ = H.o_trim(seq, qual, None, L, stride, 33, 20, 20)
    assert np.array_equal(out[:m].cpu().numpy(), exp)
    tail, _ = H.o_trim(*H.synth_slab(H.SEED_BASE + 1, 1000, L, first=n - 1000)[0:2], None, L, stride, 33, 20, 20)
    assert np.array_equal(out[n - 1000:].cpu().numpy(), tail)
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.report_reset()
    ctx.filter_dev(ctx.batch(dseq, dqual, n, stride, L), 33, 2, 100, keep)
    rep = ctx.sync()
    assert rep.n_out == n and bool(keep.all().item())
    ctx.report_reset()
    ctx.filter_dev(ctx.batch(dseq, dqual, n, stride, L), 33, 20, 90, keep)
    rep = ctx.sync()
    expk, _ = H.o_filter(seq, qual, None, L, stride, 33, 20, 90)
    assert np.array_equal(keep[:m].cpu().numpy(), expk) and rep.n_out == int(keep.sum().item())
    del out, keep
    # revcomp involution on a slice that fits beside the input
    k = min(n, 20_000_000)
    o1s = torch.empty((k, stride), dtype=torch.uint8, device="cuda"); o1q = torch.empty_like(o1s)
    o2s = torch.empty_like(o1s); o2q = torch.empty_like(o1s)
    ctx.revcomp_dev(ctx.batch(dseq, dqual, k, stride, L), 33, o1s, o1q)
    ctx.revcomp_dev(ctx.batch(o1s, o1q, k, stride, L), 33, o2s, o2q)
    rep = ctx.sync()
    assert rep.first_bad_read == -1
    assert torch.equal(o2s, dseq[:k]) and torch.equal(o2q, dqual[:k])
    eseq, equal = H.o_revcomp(seq, qual, None, L, stride)
    assert np.array_equal(o1s[:m].cpu().numpy(), eseq) and np.array_equal(o1q[:m].cpu().numpy(), equal)
