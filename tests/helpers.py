"""Shared test helpers: oracle ctypes binding, numpy synthetic generator, tiny FASTA/FASTQ I/O.

The oracle (oracle/libfastx_oracle.so) is TEST INFRASTRUCTURE; it is loaded only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_fixtures")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


class FxoAlign(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("matches", "mismatches", "neutral", "gaps", "query_start",
                                        "query_end", "target_start", "target_end")] + [("score", C.c_float)]


class FxoClipOpts(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("min_length", "keep_delta", "discard_non_clipped",
                                        "discard_clipped", "discard_unknown", "min_adapter_len")]


_oracle = None


def oracle():
    """Load (building if necessary) the C restatement."""
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(ROOT, "oracle", "libfastx_oracle.so")
    src = os.path.join(ROOT, "oracle", "fastx_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.fxo_seq_first_invalid.argtypes = [u8p, C.c_int]
    L.fxo_qual_first_invalid.argtypes = [u8p, C.c_int, C.c_int]
    L.fxo_trim_record.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fxo_filter_record.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fxo_revcomp_record.argtypes = [u8p, u8p, C.c_int, u8p, u8p]
    L.fxo_trim_batch.argtypes = [u8p, u8p, i32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, i32p, i64p]
    L.fxo_filter_batch.argtypes = [u8p, u8p, i32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, u8p, i64p]
    L.fxo_revcomp_batch.argtypes = [u8p, u8p, i32p, C.c_int, C.c_int, C.c_int64, u8p, u8p]
    L.fxo_stats_new.restype = C.c_void_p
    L.fxo_stats_new.argtypes = [C.c_int]
    L.fxo_stats_free.argtypes = [C.c_void_p]
    L.fxo_stats_add.argtypes = [C.c_void_p, u8p, u8p, C.c_int, C.c_int, C.c_int]
    L.fxo_stats_add_batch.argtypes = [C.c_void_p, u8p, u8p, i32p, C.c_int, C.c_int, C.c_int64, C.c_int]
    L.fxo_stats_export_hist.argtypes = [C.c_void_p, u64p, C.c_int]
    L.fxo_stats_print_path.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.fxo_align.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.POINTER(FxoAlign)]
    L.fxo_adapter_cutoff_index.argtypes = [C.POINTER(FxoAlign), C.c_int, C.c_int]
    L.fxo_clip_record.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.POINTER(FxoClipOpts), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fxo_clip_batch.argtypes = [u8p, i32p, i32p, C.c_int, C.c_int, C.c_int64, u8p, C.c_int, C.POINTER(FxoClipOpts), i32p, u8p, i32p]
    L.fxo_mask_batch.argtypes = [u8p, u8p, i32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, u8p, u8p, i64p, i64p]
    L.fxo_artifacts_batch.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int64, u8p]
    L.fxo_has_n_batch.argtypes = [u8p, i32p, C.c_int, C.c_int, C.c_int64, u8p]
    L.fxo_fastx_trimmer_record.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.fxo_hash_bytes.restype = C.c_uint64
    L.fxo_hash_bytes.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
    L.fxo_collapser_new.restype = C.c_void_p
    L.fxo_collapser_free.argtypes = [C.c_void_p]
    L.fxo_collapser_add.argtypes = [C.c_void_p, u8p, C.c_int, C.c_uint64]
    L.fxo_collapser_add_batch.argtypes = [C.c_void_p, u8p, i32p, C.c_int, C.c_int, C.c_int64]
    L.fxo_collapser_unique.restype = C.c_int64
    L.fxo_collapser_unique.argtypes = [C.c_void_p]
    L.fxo_collapser_order.argtypes = [C.c_void_p, i64p, u64p]
    L.fxo_collapser_print_path.argtypes = [C.c_void_p, C.c_char_p]
    _oracle = L
    return L


# ----------------------------------------------------------------------------- oracle wrappers

def o_trim(seq, qual, lens, L, stride, Q, t, min_len):
    n = qual.shape[0]
    out = np.empty(n, np.int32)
    bad = C.c_int64(-1)
    oracle().fxo_trim_batch(_p(seq, u8p), _p(qual, u8p), _p(lens, i32p), L, stride, n, Q, t, min_len,
                            _p(out, i32p), C.byref(bad))
    return out, bad.value


def o_filter(seq, qual, lens, L, stride, Q, q, p):
    n = qual.shape[0]
    out = np.empty(n, np.uint8)
    bad = C.c_int64(-1)
    oracle().fxo_filter_batch(_p(seq, u8p), _p(qual, u8p), _p(lens, i32p), L, stride, n, Q, q, p,
                              _p(out, u8p), C.byref(bad))
    return out, bad.value


def o_revcomp(seq, qual, lens, L, stride):
    n = seq.shape[0]
    oseq = np.zeros_like(seq)
    oqual = None if qual is None else np.zeros_like(qual)
    oracle().fxo_revcomp_batch(_p(seq, u8p), _p(qual, u8p), _p(lens, i32p), L, stride, n, _p(oseq, u8p), _p(oqual, u8p))
    return oseq, oqual


def o_stats_hist(seq, qual, lens, L, stride, Q, max_cycles):
    n = seq.shape[0]
    s = oracle().fxo_stats_new(max_cycles)
    oracle().fxo_stats_add_batch(s, _p(seq, u8p), _p(qual, u8p), _p(lens, i32p), L, stride, n, Q)
    hist = np.zeros((max_cycles, 5, 109), np.uint64)
    cyc = oracle().fxo_stats_export_hist(s, _p(hist, u64p), max_cycles)
    oracle().fxo_stats_free(s)
    return hist, cyc


def o_clip(seq, lens, widths, L, stride, adapter, opts):
    n = seq.shape[0]
    out_len = np.empty(n, np.int32)
    out_cls = np.empty(n, np.uint8)
    out_cut = np.empty(n, np.int32)
    ad = np.frombuffer(adapter, np.uint8).copy()
    oracle().fxo_clip_batch(_p(seq, u8p), _p(lens, i32p), _p(widths, i32p), L, stride, n, _p(ad, u8p), len(adapter),
                            C.byref(opts), _p(out_len, i32p), _p(out_cls, u8p), _p(out_cut, i32p))
    return out_len, out_cls, out_cut


def o_mask(seq, qual, lens, L, stride, Q, q, ch):
    n = seq.shape[0]
    out = np.zeros_like(seq)
    flag = np.zeros(n, np.uint8)
    mr, mb = C.c_int64(0), C.c_int64(0)
    oracle().fxo_mask_batch(_p(seq, u8p), _p(qual, u8p), _p(lens, i32p), L, stride, n, Q, q, ch, _p(out, u8p), _p(flag, u8p),
                            C.byref(mr), C.byref(mb))
    return out, flag, mr.value, mb.value


def o_has_n(seq, lens, L, stride):
    n = seq.shape[0]
    f = np.empty(n, np.uint8)
    oracle().fxo_has_n_batch(_p(seq, u8p), _p(lens, i32p), L, stride, n, _p(f, u8p))
    return f


def o_artifacts(seq, lens, L, stride):
    n = seq.shape[0]
    keep = np.zeros(n, np.uint8)
    oracle().fxo_artifacts_batch(_p(seq, u8p), _p(lens, i32p), L, stride, n, _p(keep, u8p))
    return keep


def o_fastx_trimmer(length, first, last, trim_last, min_len):
    st = C.c_int(0)
    nl = oracle().fxo_fastx_trimmer_record(length, first, last, trim_last, min_len, C.byref(st))
    return nl, st.value


def o_collapse(seq, lens, L, stride):
    n = seq.shape[0]
    c = oracle().fxo_collapser_new()
    oracle().fxo_collapser_add_batch(c, _p(seq, u8p), _p(lens, i32p), L, stride, n)
    u = oracle().fxo_collapser_unique(c)
    first = np.empty(u, np.int64)
    cnt = np.empty(u, np.uint64)
    oracle().fxo_collapser_order(c, _p(first, i64p), _p(cnt, u64p))
    oracle().fxo_collapser_free(c)
    return first, cnt


# ----------------------------------------------------------------------------- synthetic data

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def sm64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


PLAIN, WITH_N, ADAPTER, DUPS = 0, 1, 2, 3
SEED_BASE = 20260925


def synth_slab(seed, n, L, kind=PLAIN, stride=None, q_offset=33, first=0, n_total=None):
    """numpy twin of include/fxg_synth.h. Returns (seq[n,stride], qual[n,stride]) uint8, zero padded."""
    stride = stride or ((L + 15) // 16) * 16
    n_total = n_total or (first + n)
    with np.errstate(over="ignore"):
        seed = np.uint64(seed)
        idx = np.arange(first, first + n, dtype=np.uint64)
        r = sm64(seed ^ sm64(idx))
        if kind == DUPS:
            g = sm64(seed ^ np.uint64(0x5bd1e995c0ffee) ^ sm64(idx))
            dup = (g % np.uint64(100)) < np.uint64(40)
            pool = np.uint64(n_total // 8 + 1)
            u = (g >> np.uint64(8)) & np.uint64(0xFFFFFF)
            pid = (((u * u) >> np.uint64(24)) * pool) >> np.uint64(24)
            rd = sm64(seed ^ sm64(np.uint64(0x4000000000000000) | pid))
            r = np.where(dup, rd, r)
        pos = np.arange(L, dtype=np.uint64)
        h = sm64(r[:, None] + pos[None, :])
        base = np.frombuffer(b"ACGT", np.uint8)[(h & np.uint64(3)).astype(np.int64)]
        if kind == WITH_N:
            base = np.where(((h >> np.uint64(2)) & np.uint64(1023)) == 0, np.uint8(ord("N")), base)
        if kind == ADAPTER and L > 20:
            has = ((r >> np.uint64(32)) % np.uint64(100)) < np.uint64(30)
            start = 20 + ((r >> np.uint64(40)) % np.uint64(L - 20)).astype(np.int64)
            off = np.arange(L, dtype=np.int64)[None, :] - start[:, None]
            ad = np.frombuffer(b"AGATCGGAAGAGC", np.uint8)
            m = has[:, None] & (off >= 0) & (off < 13)
            base = np.where(m, ad[np.clip(off, 0, 12)], base)
        noise = (((h >> np.uint64(12)) & np.uint64(15)) + ((h >> np.uint64(16)) & np.uint64(15))
                 + ((h >> np.uint64(20)) & np.uint64(15))).astype(np.int64) - 22
        p = np.arange(L, dtype=np.int64)
        q = 38 - (18 * p * p) // (L * L) + noise
        q = np.clip(q, 2, 40)
    seq = np.zeros((n, stride), np.uint8)
    qual = np.zeros((n, stride), np.uint8)
    seq[:, :L] = base
    qual[:, :L] = (q + q_offset).astype(np.uint8)
    return seq, qual


def ragged(seq, qual, rng, min_len=1):
    """Give every read a random length in [min_len, L]; zero the tail. Returns lens."""
    n, stride = seq.shape
    L = int((seq != 0).sum(axis=1).max())
    lens = rng.integers(min_len, L + 1, size=n).astype(np.int32)
    mask = np.arange(stride)[None, :] >= lens[:, None]
    seq[mask] = 0
    if qual is not None:
        qual[mask] = 0
    return lens


# ----------------------------------------------------------------------------- tiny text I/O

def write_fastq(path, seq, qual, lens=None, L=None, prefix="r", first=0):
    n = seq.shape[0]
    with open(path, "wb") as f:
        for i in range(n):
            l = int(lens[i]) if lens is not None else L
            f.write(b"@%s%d\n" % (prefix.encode(), first + i))
            f.write(seq[i, :l].tobytes() + b"\n+\n" + qual[i, :l].tobytes() + b"\n")


def write_fasta(path, seq, lens=None, L=None, prefix="r"):
    n = seq.shape[0]
    with open(path, "wb") as f:
        for i in range(n):
            l = int(lens[i]) if lens is not None else L
            f.write(b">%s%d\n" % (prefix.encode(), i) + seq[i, :l].tobytes() + b"\n")


def read_fastx(path):
    """Returns list of (name, seq, name2, qual_line) as bytes; qual_line None for FASTA."""
    recs = []
    with open(path, "rb") as f:
        lines = [l.rstrip(b"\r\n") for l in f.read().split(b"\n")]
    if lines and lines[-1] == b"":
        lines.pop()
    if not lines:
        return recs
    if lines[0][:1] == b">":
        for i in range(0, len(lines), 2):
            recs.append((lines[i][1:], lines[i + 1], None, None))
    else:
        for i in range(0, len(lines), 4):
            recs.append((lines[i][1:], lines[i + 1], lines[i + 2][1:], lines[i + 3]))
    return recs


def qual_to_bytes(seq, qline, q_offset):
    """Mirror fastx.c:382-390: ASCII if same length as seq, else numeric tokens. Returns (bytes, is_ascii)."""
    if len(qline) == len(seq):
        return qline, True
    vals = [int(t) for t in qline.split()]
    return bytes((v + q_offset) & 0xFF for v in vals), False


def slab_from_records(recs, q_offset=33):
    L = max(len(r[1]) for r in recs)
    stride = ((L + 15) // 16) * 16
    n = len(recs)
    seq = np.zeros((n, stride), np.uint8)
    qual = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.int32)
    ascii_flags = []
    for i, (_, s, _, q) in enumerate(recs):
        lens[i] = len(s)
        seq[i, :len(s)] = np.frombuffer(s, np.uint8)
        if q is not None:
            qb, is_ascii = qual_to_bytes(s, q, q_offset)
            qual[i, :len(s)] = np.frombuffer(qb, np.uint8)
            ascii_flags.append(is_ascii)
    return seq, qual, lens, stride, ascii_flags


def ref_tool(name):
    p = os.path.join(REF_BIN, name)
    return p if os.path.exists(p) else None


def run(cmd, stdin_path=None, check=True):
    with open(stdin_path, "rb") if stdin_path else open(os.devnull, "rb") as fin:
        r = subprocess.run(cmd, stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if check and r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (cmd, r.returncode, r.stderr.decode(errors="replace")))
    return r
