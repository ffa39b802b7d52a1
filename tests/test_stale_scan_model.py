"""Groundwork for the clipper AFTER a trimming stage in a fused pipeline: the rows the reference's
aligner sees for mixed-length reads depend on all earlier reads (its query buffer only grows and keeps stale bytes).  The host
packer builds them sequentially (fxh.c, `shadow`); on the device they can come from ONE inclusive scan, because "overwrite a
prefix" is closed under composition and associative:

    op = (data[0..p), p)        apply(op, S)[x] = data[x] if x < p else S[x]
    (g after f) = (x < p_g ? data_g[x] : data_f[x]  for x < max(p_f, p_g),  max(p_f, p_g))

Read i contributes the operator (seq_i + NUL, len_i + 1); the inclusive scan at i, applied to the all-zero buffer, IS the
aligner's buffer after read i, and the running maximum of the lengths is its matrix width.  This test checks the algebra against
the sequential construction used by tests/test_pipeline_oracle.py (any bracketing of the scan, as a parallel scan would use)."""
import random


def compose(f, g):
    """g after f"""
    (df, pf), (dg, pg) = f, g
    p = max(pf, pg)
    return [dg[x] if x < pg else df[x] for x in range(p)], p


def sequential_rows(reads, W):
    shadow, wmax, rows = [0] * (W + 1), 0, []
    for r in reads:
        shadow[:len(r)] = r
        shadow[len(r)] = 0
        wmax = max(wmax, len(r))
        rows.append((list(shadow[:wmax]), wmax))
    return rows


def tree_scan(ops):
    """inclusive scan by recursive halving (a different bracketing than left-to-right)"""
    if len(ops) == 1:
        return ops
    mid = len(ops) // 2
    left, right = tree_scan(ops[:mid]), tree_scan(ops[mid:])
    return left + [compose(left[-1], r) for r in right]


def test_prefix_overwrite_scan_reproduces_the_aligner_buffer():
    rng = random.Random(5)
    for _ in range(50):
        W = rng.randint(4, 40)
        reads = [[rng.choice(b"ACGTN") for _ in range(rng.randint(1, W))] for _ in range(rng.randint(1, 60))]
        ops = [(r + [0], len(r) + 1) for r in reads]
        scanned = tree_scan(ops)
        wmax = 0
        for (data, p), (row, w), r in zip(scanned, sequential_rows(reads, W), reads):
            wmax = max(wmax, len(r))
            buf = data + [0] * (W + 1 - p)               # applied to the all-zero buffer
            assert w == wmax and buf[:w] == row
        # associativity on random triples
        for _ in range(20):
            a, b, c = (rng.choice(ops) for _ in range(3))
            assert compose(compose(a, b), c) == compose(a, compose(b, c))
