"""K-STATS second generation, checked without a GPU: the Python model of the kernel's schedules and addresses
(tests/stats2_model.py) must reproduce a direct per-base histogram and keep every ATOMS of the A scheme and of the
masked B scheme bank-conflict free."""
import random

import pytest

import stats2_model as SM


def make_rows(rng, n, L, S, Q, ragged, junk):
    lo, hi = Q - 15, min(Q + 93, 127)
    seqs, quals, lens = [], [], []
    for _ in range(n):
        l = rng.randint(1, L) if ragged else L
        sq = [rng.choice(b"ACGT") if rng.random() > 0.02 else ord("N") for _ in range(l)]
        ql = [rng.randint(lo, min(hi, lo + 70)) if rng.random() < 0.05 else rng.randint(Q, min(Q + 41, hi)) for _ in range(l)]
        if junk and rng.random() < 0.2 and l > 3:
            p = rng.randrange(l)
            if rng.random() < 0.5:
                sq[p] = rng.choice(b"acgtXU\x00\xc1")
            else:
                ql[p] = rng.choice([lo - 1, (hi + 1) & 0xFF, 200, 0]) & 0xFF
        pad = [rng.randrange(256) if junk else 0 for _ in range(S - l)]
        seqs.append(sq + pad)
        quals.append(ql + [rng.randrange(256) if junk else 0 for _ in range(S - l)])
        lens.append(l)
    return seqs, quals, lens


@pytest.mark.parametrize("n,L,S,Q,ragged,junk", [
    (37, 150, 160, 33, False, False), (21, 100, 112, 33, False, True), (19, 50, 64, 64, True, True), (9, 300, 304, 33, True, False),
    (11, 36, 48, 33, False, False), (13, 127, 128, 33, True, True), (10, 160, 160, 33, False, False), (6, 163, 176, 33, False, False),
    (9, 129, 144, 33, False, False), (17, 68, 80, 64, True, False), (5, 3, 16, 33, False, False)])
@pytest.mark.parametrize("bscheme,layout16", [(0, False), (1, False), (0, True)])
def test_model_equals_direct_histogram(n, L, S, Q, ragged, junk, bscheme, layout16):
    rng = random.Random(1000 * L + S + bscheme)
    seqs, quals, lens = make_rows(rng, n, L, S, Q, ragged, junk)
    hist, bad = SM.run_model(seqs, quals, lens, S, Q, bscheme=bscheme, layout16=layout16)
    exp, exp_bad = SM.direct_hist(seqs, quals, lens, Q)
    assert bad == exp_bad
    assert hist == exp
