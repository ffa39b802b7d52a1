"""CPU check of the CRC-32 algebra behind `-z` (fxg_deflate.cu: the GPU computes a "pure" CRC — no initial / final inversion — per
64 KB block; fxg_crc32_concat() chains them with x^(8*len) mod P, fxg_crc32_finish() turns the chain into the gzip trailer value
that the reference's `gzip` child would write, fastx.c:214-248).  Both are host functions of libfxg.so: no GPU needed."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastx_toolkit_b200", "libfxg.so")


def pure(data):
    return (zlib.crc32(data, 0xFFFFFFFF) ^ 0xFFFFFFFF) & 0xFFFFFFFF


@pytest.mark.skipif(not os.path.exists(LIB), reason="libfxg.so not built")
def test_crc32_concat_and_finish_match_zlib():
    L = C.CDLL(LIB)
    L.fxg_crc32_concat.restype = C.c_uint32
    L.fxg_crc32_concat.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.fxg_crc32_finish.restype = C.c_uint32
    L.fxg_crc32_finish.argtypes = [C.c_uint32, C.c_uint64]
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 7, 64, 65535, 65536, 65537, 1 << 20, (1 << 22) + 13):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert L.fxg_crc32_finish(pure(data), n) == zlib.crc32(data)
        # chained in 64 KB blocks and in random pieces, as the writer does across chunks
        for cuts in (list(range(0, n, 65536)) + [n], sorted(set([0, n] + [int(c) for c in rng.integers(0, n + 1, 5)]))):
            crc = 0
            for a, b in zip(cuts[:-1], cuts[1:]):
                crc = L.fxg_crc32_concat(crc, pure(data[a:b]), b - a)
            assert crc == pure(data), (n, cuts)
    # a long run of zero bytes: only the length enters
    z = bytes(3_000_000)
    assert L.fxg_crc32_finish(L.fxg_crc32_concat(pure(b"abc"), pure(z), len(z)), 3 + len(z)) == zlib.crc32(b"abc" + z)
