"""The synthetic workload exists three times: include/fxg_synth.h compiled by gcc (bin/fxg_synth — feeds the reference
binaries), the same header compiled by nvcc (k_synth — feeds the GPU tests and bench.py) and tests/helpers.py synth_slab (numpy —
feeds the oracle).  This pins the header against the numpy twin on the CPU, for every workload kind."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H

EXE = os.path.join(H.ROOT, "bin", "fxg_synth")
KINDS = [("plain", H.PLAIN), ("n", H.WITH_N), ("adapter", H.ADAPTER), ("dups", H.DUPS)]


@pytest.mark.skipif(not os.path.exists(EXE), reason="bin/fxg_synth not built")
@pytest.mark.parametrize("kname,kind", KINDS)
@pytest.mark.parametrize("n,L,first,Q", [(3000, 150, 0, 33), (2000, 50, 12345, 33), (500, 21, 7, 64), (300, 250, 0, 33)])
def test_c_generator_equals_numpy_twin(kname, kind, n, L, first, Q):
    total = first + n + 1000
    seed = H.SEED_BASE + kind
    r = subprocess.run([EXE, "-n", str(n), "-l", str(L), "-k", kname, "-s", str(seed), "-f", str(first), "-T", str(total), "-Q", str(Q)],
                       stdout=subprocess.PIPE, check=True)
    seq, qual = H.synth_slab(seed, n, L, kind, q_offset=Q, first=first, n_total=total)
    exp = b"".join(b"@r%d\n" % (first + i) + seq[i, :L].tobytes() + b"\n+\n" + qual[i, :L].tobytes() + b"\n" for i in range(n))
    assert r.stdout == exp


@pytest.mark.skipif(not os.path.exists(EXE), reason="bin/fxg_synth not built")
def test_dups_kind_has_the_repeat_structure_of_config_e():
    """config (e): about 40 % of the reads repeat earlier pool sequences"""
    n, L = 200000, 50
    seq, _ = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    uniq = len({seq[i, :L].tobytes() for i in range(n)})
    assert 0.55 * n < uniq < 0.80 * n
