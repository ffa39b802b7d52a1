"""GPU parity (-m gpu): K-HASH, exact dedup and the libstdc++-order emulation vs the CPU oracle (which itself
is pinned to the real fastx_collapser binary in test_oracle_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import ctx, dev  # noqa: F401

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def gpu_collapse(seq, lens, L, weight=None, first=None, splits=1):
    import fastx_toolkit_b200 as F
    n, stride = seq.shape
    col = F.Collapser(0, max(n, 1), stride)
    bounds = np.linspace(0, n, splits + 1).astype(np.int64)
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b == a:
            continue
        bt = F.Batch(seq[a:b].ctypes.data, None, None if lens is None else lens[a:b].ctypes.data, L, stride, int(b - a))
        col.add(bt, None if weight is None else weight[a:b], None if first is None else first[a:b], int(a))
    u = col.finish(True)
    oseq = np.zeros((u, stride), np.uint8); olen = np.zeros(u, np.int32)
    ocnt = np.zeros(u, np.uint64); ofirst = np.zeros(u, np.int64); ohash = np.zeros(u, np.uint64)
    col.fetch(oseq, olen, ocnt, ofirst, ohash)
    assert col.first_bad == -1 and col.launches() > 0
    col.close()
    return oseq, olen, ocnt, ofirst, ohash


def check_against_oracle(seq, lens, L, splits=1):
    n, stride = seq.shape
    efirst, ecnt = H.o_collapse(seq, lens, L, stride)
    oseq, olen, ocnt, ofirst, ohash = gpu_collapse(seq, lens, L, splits=splits)
    assert len(ocnt) == len(ecnt)
    assert np.array_equal(ocnt, ecnt)
    assert np.array_equal(ofirst, efirst), "output order differs from the reference's unordered_map order"
    ll = lens if lens is not None else np.full(n, L, np.int32)
    assert np.array_equal(olen, ll[efirst])
    for k in (0, len(efirst) // 2, len(efirst) - 1):
        assert oseq[k, :olen[k]].tobytes() == seq[efirst[k], :olen[k]].tobytes()
    return len(ecnt)


def test_hash_matches_std_hash(ctx):
    n, L = 20000, 50
    seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, H.WITH_N)
    lens = H.ragged(seq, None, np.random.default_rng(1), min_len=1)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    dseq, dlens = dev(seq), dev(lens)
    ctx.hash_dev(ctx.batch(dseq, None, n, seq.shape[1], 0, dlens), out)
    ctx.sync()
    got = out.cpu().numpy().astype(np.uint64)
    O = H.oracle()
    for i in list(range(50)) + [n - 1, n // 2]:
        row = np.ascontiguousarray(seq[i])
        assert got[i] == O.fxo_hash_bytes(row.ctypes.data_as(C.c_void_p), int(lens[i]), 0xc70f6907)


@pytest.mark.parametrize("n,L", [(300000, 50), (40000, 23), (2000000, 36), (5000, 150)])
def test_collapse_order_matches_reference_map(ctx, n, L):
    seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    u = check_against_oracle(seq, None, L)
    assert 0.5 * n < u < 0.95 * n


@pytest.mark.parametrize("u", [1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 10273, 10274, 20754])
def test_collapse_epoch_boundaries(ctx, u):
    """exactly u distinct keys (+ some repeats): bucket-count ladder edges 13/29/59/... are where rehashes fire"""
    L = 40
    base, _ = H.synth_slab(H.SEED_BASE + 6, u, L, H.PLAIN)
    rng = np.random.default_rng(u)
    extra = base[rng.integers(0, u, size=u // 2 + 3)]
    seq = np.concatenate([base[: u // 2 + 1], extra[: u // 3], base[u // 2 + 1:], extra[u // 3:]])
    check_against_oracle(np.ascontiguousarray(seq), None, L, splits=3)


def test_collapse_ragged_prefix_keys_and_weights(ctx):
    L = 30
    rng = np.random.default_rng(9)
    seq, _ = H.synth_slab(H.SEED_BASE + 6, 4000, L, H.WITH_N)
    seq[1000:2000] = seq[:1000]            # same bytes, different lengths => different keys
    lens = rng.integers(1, L + 1, size=4000).astype(np.int32)
    lens[1000:1500] = lens[:500]           # ... except these, which are true duplicates
    mask = np.arange(seq.shape[1])[None, :] >= lens[:, None]
    seq[mask] = 0
    check_against_oracle(seq, lens, 0, splits=2)
    # weights and explicit first indices (what an owner GPU receives from its peers)
    n = 4000
    w = rng.integers(1, 50, size=n).astype(np.int32)
    f = rng.permutation(n).astype(np.int64) * 3
    oseq, olen, ocnt, ofirst, ohash = gpu_collapse(seq, lens, 0, weight=w, first=f)
    keys = {}
    for i in range(n):
        k = seq[i, :lens[i]].tobytes()
        c, fm = keys.get(k, (0, 1 << 62))
        keys[k] = (c + int(w[i]), min(fm, int(f[i])))
    got = {oseq[k, :olen[k]].tobytes(): (int(ocnt[k]), int(ofirst[k])) for k in range(len(ocnt))}
    assert got == keys
    assert all(ocnt[k] >= ocnt[k + 1] for k in range(len(ocnt) - 1))


def test_collapse_golden_fixture(ctx):
    recs = H.read_fastx(os.path.join(H.GOLDEN, "fasta_collapser1.fasta"))
    seq, _, lens, stride, _ = H.slab_from_records(recs)
    check_against_oracle(seq, lens, 0)
    exp = H.read_fastx(os.path.join(H.GOLDEN, "fasta_collapser1.out"))
    oseq, olen, ocnt, ofirst, _ = gpu_collapse(seq, lens, 0)
    assert [int(c) for c in ocnt] == [int(r[0].split(b"-")[1]) for r in exp]
    for k in range(4):
        assert oseq[k, :olen[k]].tobytes() == exp[k][1]


def test_collapse_bad_read_detected(ctx):
    import fastx_toolkit_b200 as F
    seq, _ = H.synth_slab(H.SEED_BASE + 4, 3000, 50, H.DUPS)
    seq[1234, 7] = ord("x")
    col = F.Collapser(0, 3000, seq.shape[1])
    col.add(F.Batch(seq.ctypes.data, None, None, 50, seq.shape[1], 3000))
    col.finish(False)
    assert col.first_bad == 1234
    col.close()


def test_collapse_sharded_like_multi_gpu(ctx):
    """The multi-GPU algorithm on one GPU: G shards dedup locally, uniques are routed to owner = hash mod G,
    owners merge (weights + first indices), the triples are gathered and ordered once.  Must equal the
    single-table result (and therefore the reference)."""
    import fastx_toolkit_b200 as F
    n, L, G = 400000, 50, 3
    seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    stride = seq.shape[1]
    efirst, ecnt = H.o_collapse(seq, None, L, stride)
    bounds = np.linspace(0, n, G + 1).astype(np.int64)
    parts = []
    for g in range(G):
        a, b = int(bounds[g]), int(bounds[g + 1])
        col = F.Collapser(0, b - a, stride)
        col.add(F.Batch(seq[a:b].ctypes.data, None, None, L, stride, b - a), None, None, a)   # first = global index
        u = col.finish(False)
        s = np.zeros((u, stride), np.uint8); ln = np.zeros(u, np.int32)
        c = np.zeros(u, np.uint64); f = np.zeros(u, np.int64); h = np.zeros(u, np.uint64)
        col.fetch(s, ln, c, f, h)
        col.close()
        parts.append((s, ln, c, f, h))
    trip = []
    for owner in range(G):
        s = np.concatenate([p[0][p[4] % np.uint64(G) == owner] for p in parts])
        ln = np.concatenate([p[1][p[4] % np.uint64(G) == owner] for p in parts])
        c = np.concatenate([p[2][p[4] % np.uint64(G) == owner] for p in parts]).astype(np.int32)
        f = np.concatenate([p[3][p[4] % np.uint64(G) == owner] for p in parts])
        col = F.Collapser(0, max(len(ln), 1), stride)
        col.add(F.Batch(s.ctypes.data, None, ln.ctypes.data, 0, stride, len(ln)), c, f, 0)
        u = col.finish(False)
        oc = np.zeros(u, np.uint64); of = np.zeros(u, np.int64); oh = np.zeros(u, np.uint64)
        col.fetch(None, None, oc, of, oh)
        col.close()
        trip.append((oh, of, oc))
    hh = torch.from_numpy(np.concatenate([t[0] for t in trip]).view(np.int64)).cuda()
    ff = torch.from_numpy(np.concatenate([t[1] for t in trip])).cuda()
    cc = torch.from_numpy(np.concatenate([t[2] for t in trip]).view(np.int64)).cuda()
    U = hh.numel()
    perm = torch.empty(U, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    F.collapse_order_dev(0, hh, ff, cc, U, perm)
    p = perm.cpu().numpy().astype(np.int64)
    assert U == len(ecnt)
    assert np.array_equal(ff.cpu().numpy()[p], efirst)
    assert np.array_equal(cc.cpu().numpy()[p].astype(np.uint64), ecnt)


def test_dcollapse_one_rank_native_exchange(ctx):
    """fxg_dcollapse_* (the multi-GPU collapser) on a communicator of ONE GPU: every phase runs — K-ROUTE, the NCCL
    send/recv exchange (to itself), owner dedup, triple gather, K-ORDER — and the result must be the reference's."""
    import fastx_toolkit_b200 as F
    for n, L, kind in ((300000, 50, H.DUPS), (50000, 150, H.DUPS), (777, 36, H.WITH_N)):
        seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, kind)
        stride = seq.shape[1]
        dseq = dev(seq)
        comm = F.Comm.all([0])
        assert comm.nranks == 1 and comm.nlocal == 1
        dc = F.DCollapser(comm, stride)
        for _ in range(2):                 # a second run reuses the buffers
            rep = dc.run([F.Batch(dseq.data_ptr(), None, None, L, stride, n)], [0])
        U = rep.n_unique
        efirst, ecnt = H.o_collapse(seq, None, L, stride)
        assert U == len(ecnt) == rep.n_unique_local and rep.first_bad_read == -1 and rep.rows_received == n
        po, pi = np.empty(U, np.int32), np.empty(U, np.uint32)
        of, oc = np.empty(U, np.int64), np.empty(U, np.uint64)
        dc.fetch_order(po, pi, of, oc)
        assert np.array_equal(of, efirst) and np.array_equal(oc, ecnt) and not po.any()
        rows, lens = np.zeros((U, stride), np.uint8), np.zeros(U, np.int32)
        cnt, first = np.zeros(U, np.uint64), np.zeros(U, np.int64)
        dc.fetch_local(0, rows, lens, cnt, first, None)
        assert np.array_equal(first[pi], efirst) and np.array_equal(cnt[pi], ecnt) and (lens == L).all()
        assert np.array_equal(rows[pi][:, :L], seq[efirst][:, :L])
        assert comm.collectives() >= 7 and dc.launches() > 0
        dc.close(); comm.close()
    # a bad read is reported with its global index; ragged lengths and weights travel in the 16-byte records
    seq, _ = H.synth_slab(H.SEED_BASE + 6, 4000, 30, H.WITH_N)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 31, size=4000).astype(np.int32)
    seq[np.arange(seq.shape[1])[None, :] >= lens[:, None]] = 0
    seq[2000:3000] = seq[:1000]; lens[2000:3000] = lens[:1000]
    w = rng.integers(1, 9, size=4000).astype(np.int32)
    comm = F.Comm.all([0]); dc = F.DCollapser(comm, seq.shape[1])
    dseq, dlens, dw = dev(seq), dev(lens), dev(w)
    rep = dc.run([F.Batch(dseq.data_ptr(), None, dlens.data_ptr(), 0, seq.shape[1], 4000)], [1000], [dw])
    U = rep.n_unique
    rows, ol, cnt, first = np.zeros((U, seq.shape[1]), np.uint8), np.zeros(U, np.int32), np.zeros(U, np.uint64), np.zeros(U, np.int64)
    dc.fetch_local(0, rows, ol, cnt, first, None)
    keys = {}
    for i in range(4000):
        k = seq[i, :lens[i]].tobytes()
        c, fm = keys.get(k, (0, 1 << 62))
        keys[k] = (c + int(w[i]), min(fm, 1000 + i))
    assert {rows[k, :ol[k]].tobytes(): (int(cnt[k]), int(first[k])) for k in range(U)} == keys
    seq[1234, 0] = ord("x")
    dseq = dev(seq)
    rep = dc.run([F.Batch(dseq.data_ptr(), None, dlens.data_ptr(), 0, seq.shape[1], 4000)], [1000])
    assert rep.first_bad_read == 1000 + 1234
    dc.close(); comm.close()


def test_collapse_24m_reads_full_order_vs_oracle(ctx):
    """BASELINE config (e) shape, scaled to what the oracle finishes in a minute: 24 M x 50 bp, ~16 M uniques — the map
    grows through bucket counts 5 967 347, 12 117 689 and 24 607 243 (three rehash epochs beyond the 2 M-read test) and the
    FULL ordered output must equal the reference's (src/fastx_collapser/fastx_collapser.cpp:112-122)."""
    import fastx_toolkit_b200 as F
    n, L, stride = int(os.environ.get("FXG_COLLAPSE_N", 24_000_000)), 50, 64
    dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    dq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dq, n, L, stride, H.SEED_BASE + 4, H.DUPS, 33)
    ctx.sync()
    del dq
    col = F.Collapser(0, n, stride)
    col.add(F.Batch(dseq.data_ptr(), None, None, L, stride, n))
    U = col.finish(True)
    assert U > 12_117_689 and col.first_bad == -1
    ocnt, ofirst = np.zeros(U, np.uint64), np.zeros(U, np.int64)
    col.fetch(None, None, ocnt, ofirst, None)
    col.close()
    seq = dseq.cpu().numpy()
    del dseq
    efirst, ecnt = H.o_collapse(seq, None, L, stride)
    assert U == len(ecnt) and int(ocnt.sum()) == n
    assert np.array_equal(ocnt, ecnt)
    assert np.array_equal(ofirst, efirst), "output order differs from the reference's unordered_map order"


def test_collapse_config_e_200m_properties(ctx):
    """BASELINE config (e) at full size on one GPU: 200 M x 50 bp, ~40 % repeats.  Checked without the oracle: the counts sum
    to n, the number of uniques equals an independent host recount from the generator (reads that are not repeats + distinct
    pool members hit), the order is count-descending and the first indices are exactly the reads that start a key."""
    import fastx_toolkit_b200 as F
    n, L, stride = int(os.environ.get("FXG_FULL_COLLAPSE_N", 200_000_000)), 50, 64
    free = torch.cuda.mem_get_info()[0]
    if free < n * 64 * 2.6 + (8 << 30):
        pytest.skip("not enough device memory")
    seed = H.SEED_BASE + 4
    dseq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    dq = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dq, n, L, stride, seed, H.DUPS, 33)
    ctx.sync()
    del dq
    torch.cuda.empty_cache()
    col = F.Collapser(0, n, stride)
    col.add(F.Batch(dseq.data_ptr(), None, None, L, stride, n))
    U = col.finish(True)
    ocnt, ofirst = np.zeros(U, np.uint64), np.zeros(U, np.int64)
    col.fetch(None, None, ocnt, ofirst, None)
    col.close()
    assert int(ocnt.sum()) == n
    assert (ocnt[:-1] >= ocnt[1:]).all()
    assert len(np.unique(ofirst)) == U and ofirst.min() == 0 and ofirst.max() < n
    # host recount from the generator (include/fxg_synth.h fxg_synth_read_key)
    nondup, pids = 0, []
    with np.errstate(over="ignore"):
        for a in range(0, n, 10_000_000):
            idx = np.arange(a, min(a + 10_000_000, n), dtype=np.uint64)
            g = H.sm64(np.uint64(seed) ^ np.uint64(0x5bd1e995c0ffee) ^ H.sm64(idx))
            dup = (g % np.uint64(100)) < np.uint64(40)
            u = (g >> np.uint64(8)) & np.uint64(0xFFFFFF)
            pid = (((u * u) >> np.uint64(24)) * np.uint64(n // 8 + 1)) >> np.uint64(24)
            nondup += int((~dup).sum())
            pids.append(np.unique(pid[dup]))
            # a read that is not a repeat starts its own key: its index must be one of the first indices
    distinct = len(np.unique(np.concatenate(pids)))
    assert U == nondup + distinct, (U, nondup, distinct)
    # spot check: the keys of the 3 most frequent and 3 of the singletons really are read `first`'s bytes, with that count
    top = dseq[torch.from_numpy(ofirst[:3]).cuda()].cpu().numpy()
    for k in range(3):
        same = (dseq[:, :L] == torch.from_numpy(top[k, :L]).cuda()).all(dim=1)
        assert int(same.sum().item()) == int(ocnt[k]) and int(torch.nonzero(same)[0].item()) == int(ofirst[k])


def test_collapser_grows_rows_stride_and_table(ctx):
    """a streaming caller knows neither the number of reads nor the longest one when it creates the table: the row store is
    re-strided when a longer batch arrives, and the table rebuilt (k_rehash) when it would get more than half full"""
    import fastx_toolkit_b200 as F
    rng = np.random.default_rng(11)
    parts = []
    for n, L in ((3000, 20), (50000, 45), (120000, 100), (40000, 30)):
        seq, _ = H.synth_slab(H.SEED_BASE + 4 + L, n, L, H.DUPS)
        parts.append((seq, L))
    col = F.Collapser(0, 1024, 32)                      # far too small on purpose
    base = 0
    for seq, L in parts:
        n, stride = seq.shape
        col.add(F.Batch(seq.ctypes.data, None, None, L, stride, n), None, None, base)
        base += n
    u = col.finish(True)
    stride = int(col.L.fxg_collapse_stride(col.h))
    assert stride == 112
    oseq, olen, ocnt, ofirst = np.zeros((u, stride), np.uint8), np.zeros(u, np.int32), np.zeros(u, np.uint64), np.zeros(u, np.int64)
    col.fetch(oseq, olen, ocnt, ofirst, None)
    col.close()
    # the oracle on the same reads, in the same order
    O = H.oracle()
    oc = O.fxo_collapser_new()
    rows = []
    for seq, L in parts:
        for i in range(seq.shape[0]):
            r = np.ascontiguousarray(seq[i])
            O.fxo_collapser_add(oc, H._p(r, H.u8p), L, 1)
            rows.append((seq, i, L))
    eu = O.fxo_collapser_unique(oc)
    efirst, ecnt = np.empty(eu, np.int64), np.empty(eu, np.uint64)
    O.fxo_collapser_order(oc, H._p(efirst, H.i64p), H._p(ecnt, H.u64p))
    O.fxo_collapser_free(oc)
    assert u == eu and np.array_equal(ocnt, ecnt) and np.array_equal(ofirst, efirst)
    for k in (0, 1, u // 3, u // 2, u - 1):
        seq, i, L = rows[int(efirst[k])]
        assert olen[k] == L and oseq[k, :L].tobytes() == seq[i, :L].tobytes()
