"""CPU model of k_stats4's lane schedules and shared-memory layout (fastx_toolkit_b200/csrc/fxg_stats4.cu): every
(lane, window word, byte) is visited exactly once, the 32 counters of every RED instruction sit in 32 different banks
whatever the data, the LDS.128 of a quarter-warp hits 8 different bank groups, and the flush decodes the layout back to
(word, byte).  No GPU needed."""
import itertools

import pytest

PITCH_WORDS = 96


def lane_consts(lane, stride):
    q8, i8 = lane >> 3, lane & 7
    odd = ((stride >> 4) & 1) != 0
    r = ((2 * i8 + (q8 & 1)) & 7) if odd else i8
    kb = (2 * (q8 >> 1) + (i8 >> 2)) if odd else q8
    return r, kb


def counter_word(w, k):
    """(word index inside the bin, half) of the counter of window word w, byte k — s4_counter()"""
    if w < 32:
        c, wi = w >> 2, w & 3
        return 32 * (wi >> 1) + 8 * k + c, wi & 1
    return 64 + 4 * (w - 32) + k, None


def flush_decode(col):
    """what the flush loop makes of word `col` of a bin: list of (w, k, half)"""
    if col < 64:
        hi, k, c = col >> 5, (col & 31) >> 3, col & 7
        w0 = 4 * c + 2 * hi
        return [(w0, k, 0), (w0 + 1, k, 1)]
    j = col - 64
    return [(32 + (j >> 2), j & 3, None)]


@pytest.mark.parametrize("stride", [64, 112, 128, 160, 176, 48])
def test_a_region_schedule_is_a_conflict_free_bijection(stride):
    seen = set()
    for t, wi, i in itertools.product(range(8), range(4), range(4)):
        banks, halves = set(), set()
        for lane in range(32):
            r, kb = lane_consts(lane, stride)
            c, k = (t + r) & 7, (i + kb) & 3
            w = 4 * c + wi
            word, half = counter_word(w, k)
            banks.add(word % 32)           # the bin contributes bin * 96 words = 0 (mod 32)
            halves.add(half)
            seen.add((lane, w, k))
        assert len(banks) == 32, (t, wi, i)
        assert len(halves) == 1            # the increment is an immediate: one half per instruction
    assert len(seen) == 32 * 32 * 4        # every lane x word 0..31 x byte exactly once (8*4*4 steps x 32 lanes)


def test_b_region_schedule_is_a_conflict_free_bijection():
    seen = set()
    for a, b in itertools.product(range(8), range(4)):
        banks = set()
        for lane in range(32):
            W, k = ((lane >> 2) + a) & 7, (b + lane) & 3
            word, half = counter_word(32 + W, k)
            assert half is None
            banks.add(word % 32)
            seen.add((lane, 32 + W, k))
        assert len(banks) == 32, (a, b)
    assert len(seen) == 32 * 8 * 4


@pytest.mark.parametrize("stride", [32, 48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208])
def test_lds128_quarter_warps_hit_8_bank_groups(stride):
    """a 16-byte shared load is served per quarter-warp: the 8 lanes must sit in 8 different 16-byte bank groups"""
    for t in range(8):
        for q in range(4):
            groups = set()
            for lane in range(8 * q, 8 * q + 8):
                r, _ = lane_consts(lane, stride)
                addr16 = lane * (stride >> 4) + ((t + r) & 7)
                groups.add(addr16 % 8)
            assert len(groups) == 8, (stride, t, q)


def test_lane_constants_are_a_bijection_onto_8x4():
    for stride in (160, 112):
        assert len({lane_consts(lane, stride) for lane in range(32)}) == 32


def test_flush_inverts_the_layout():
    back = {}
    for col in range(PITCH_WORDS):
        for w, k, half in flush_decode(col):
            back[(w, k)] = (col, half)
    assert len(back) == 160
    for w in range(40):
        for k in range(4):
            assert back[(w, k)] == counter_word(w, k)


def test_u16_halves_cannot_wrap_between_flushes():
    warps, threads = 12, 12 * 32
    rounds = 65535 // threads
    assert rounds * warps * 32 <= 65535      # one increment per read and counter at most
