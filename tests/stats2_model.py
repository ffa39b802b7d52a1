"""Python model of K-STATS' second-generation kernel (fastx_toolkit_b200/csrc/fxg_stats.cu, k_stats2): the lane schedules
(A scheme, masked B scheme, static B scheme), the packed decode, the counter addresses and the flush, restated integer
for integer so that the design can be checked WITHOUT a GPU:

  * the histogram it produces must equal a direct per-base count (src/fastx_quality_stats/fastx_quality_stats.c:166-216),
    for uniform / ragged lengths, junk padding, illegal bytes, N bases, q' >= 64, several passes (reads > 160 bases);
  * every shared-memory increment instruction of the A scheme and of the masked B scheme must be bank-conflict free
    (32 lanes on 32 distinct banks — the point of the [bin][40k + w] layout), as must the tile loads at stride 160.

layout16=True models the experimental third-generation layout (fxg_stats3.cu): u16 counters paired along k (bytes k and
k+2 of a word share one 32-bit word, increments 1 / 65536), where only the A scheme stays strictly conflict free (the
masked B scheme meets the partner byte's lane on the SAME word: a 2-way same-address collision by design).

Test infrastructure only (tests/test_stats2_model.py); nothing here is on the product path.
"""

M32 = 0xFFFFFFFF
VLUT_LO, VLUT_HI, V2LUT_HI = 0x43FF41FF, 0x474EFF54, 0x47FFFF54      # fxg_device.cuh / fxg_stats.cu tables
N6_LO, N6_HI = 0x40000000, 0x800000C0
NLUT_LO, NLUT_HI = 0x01800080, 0x02048003
PITCH, MAXW = 640, 40                                                 # S2_PITCH, ST_MAXW


def prmt(a, b, sel):
    """PTX prmt.b32 (default mode): nibble bit 3 replicates the sign of the selected byte"""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    r = 0
    for i in range(4):
        n = (sel >> (4 * i)) & 0xF
        v = src[n & 7]
        if n & 8:
            v = 0xFF if v & 0x80 else 0
        r |= v << (8 * i)
    return r


def decode(sw, qw, lo4):
    """stats2_decode: (test word, comb) — test == 0 iff four plain A/C/G/T bases with 0 <= q' < 64"""
    y = sw & 0x07070707
    sel = prmt((y | (y >> 4)) & M32, 0, 0x4420)
    e = prmt(VLUT_LO, V2LUT_HI, sel)
    n6 = prmt(N6_LO, N6_HI, sel)
    comb = (n6 + qw - lo4) & M32
    return ((sw ^ e) | ((comb ^ n6) & 0xC0C0C0C0)) & M32, comb


class Pass:
    def __init__(self, q_offset, w0, nw, max_cycles, layout16=False):
        self.layout16 = layout16
        self.lo, self.hi = q_offset - 15, min(q_offset + 93, 127)
        self.lo4 = (self.lo * 0x01010101) & M32
        self.hs = [0] * (256 * 160)          # the shared histogram, u32 [bin][40k + w]  (layout16: [bin][40(k&1) + w], two u16 halves)
        self.glob = {}                       # the global u64 table (cycle, nuc, q')
        self.w0, self.nw, self.max_cycles = w0, nw, max_cycles
        self.conflicts = 0

    def counter(self, bin_, k, o):
        """(byte address, increment) of the counter of (bin, byte k of the word at byte offset o)"""
        if self.layout16:
            return bin_ * 384 + (k & 1) * 160 + o, (65536 if k & 2 else 1)      # pitch padded to 96 words: a multiple of 32 banks
        return bin_ * PITCH + k * 160 + o, 1

    def gadd(self, cyc, nuc, qp, v=1):
        if cyc < self.max_cycles:
            self.glob[(cyc, nuc, qp)] = self.glob.get((cyc, nuc, qp), 0) + v

    def byte(self, c, q, wrel, k):
        """stats2_byte: the exact per-base path; returns 1 for an illegal base / quality"""
        code = c & 7
        legal = prmt(VLUT_LO, VLUT_HI, code) & 0xFF
        nuc = prmt(NLUT_LO, NLUT_HI, code) & 0xFF
        qp = (q - self.lo) & M32
        if legal != c or qp > self.hi - self.lo:
            return 1
        if nuc < 4 and qp < 64:
            a, inc = self.counter(nuc * 64 + qp, k, 4 * wrel)
            self.hs[a // 4] += inc
        else:
            self.gadd(4 * (self.w0 + wrel) + k, nuc, qp)
        return 0

    def flush(self):
        for i, v in enumerate(self.hs):
            if not v:
                continue
            if self.layout16:
                b, r = divmod(i, 96)
                assert r < 80
                kk, wr = divmod(r, 40)
                lo, hi = v & 0xFFFF, v >> 16
                assert hi < 65536
                if lo:
                    self.gadd(4 * (self.w0 + wr) + kk, b >> 6, b & 63, lo)
                if hi:
                    self.gadd(4 * (self.w0 + wr) + kk + 2, b >> 6, b & 63, hi)
            else:
                b, pc = divmod(i, 160)
                k, wr = divmod(pc, 40)
                self.gadd(4 * (self.w0 + wr) + k, b >> 6, b & 63, v)


def run_model(seqs, quals, lens, stride, q_offset, bscheme=0, tile_reads=8, layout16=False):
    """seqs/quals: rows of `stride` byte values; returns (global histogram dict, set of bad reads)"""
    n = len(lens)
    ragged = any(l != lens[0] for l in lens)
    words = (stride + 3) // 4 if ragged else (lens[0] + 3) // 4
    max_cycles = max(lens)
    bad_reads, hist = set(), {}
    for w0 in range(0, words, MAXW):
        nw = min(words - w0, MAXW)
        P = Pass(q_offset, w0, nw, max_cycles, layout16)
        passoff, ncols = 4 * w0, 4 * nw
        nsb, nb8 = (1 if nw > 16 else 0), (nw + 7) >> 3
        for tile in range((n + tile_reads - 1) // tile_reads):
            lanes = []
            for lane in range(32):
                j, rr = lane & 3, lane >> 2
                g = tile * tile_reads + rr
                active = rr < tile_reads and g < n
                L = lens[g] if active else 0
                if active and (L <= 0 or L > stride):
                    bad_reads.add(g)
                    L = 0
                lanes.append((j, rr, g, min(L - passoff, ncols)))

            def word(o, g):
                s = int.from_bytes(bytes(seqs[g][passoff + o:passoff + o + 4]), "little")
                q = int.from_bytes(bytes(quals[g][passoff + o:passoff + o + 4]), "little")
                return s, q

            def step(offsets, dyn_k, masked, may_conflict=False):
                """one word per lane: 4 ATOMS instructions (i = 0..3) across the warp"""
                addrs = [[None] * 32 for _ in range(4)]
                load_banks = []
                for lane, (j, rr, g, Lp) in enumerate(lanes):
                    o = offsets[lane]
                    vb = Lp - o
                    if not masked:
                        if vb < 4:
                            continue
                        vb = 4
                    if vb > 0:
                        load_banks.append(((rr * stride + passoff + o) // 4) % 32)
                        sw, qw = word(o, g)
                    else:
                        sw = qw = 0
                    m = M32 if vb >= 4 else (0 if vb <= 0 else (1 << (8 * vb)) - 1)
                    t, comb = decode((sw & m) | (0x41414141 & ~m & M32), (qw & m) | (P.lo4 & ~m & M32), P.lo4)
                    if t == 0:
                        for i in range(4):
                            k = (j + i) & 3 if dyn_k else i
                            if k >= vb:
                                continue
                            a, inc = P.counter((comb >> (8 * k)) & 0xFF, k, o)
                            addrs[i][lane] = a
                            P.hs[a // 4] += inc
                    else:
                        bd = 0
                        for k in range(min(vb, 4)):
                            bd |= P.byte((sw >> (8 * k)) & 0xFF, (qw >> (8 * k)) & 0xFF, o >> 2, k)
                        if bd:
                            bad_reads.add(g)
                if stride == 160 and not may_conflict:
                    assert len(load_banks) == len(set(load_banks)), "tile load bank conflict"
                for i in range(4):
                    banks = [(a // 4) % 32 for a in addrs[i] if a is not None]
                    if len(banks) != len(set(banks)) and not may_conflict:
                        P.conflicts += 1

            if nsb:                                   # A scheme: lane (rr, j) takes word 4*((t+rr)&7) + j
                for t in range(8):
                    step([(16 * rr + 4 * j + 16 * t) & 0x7F for (j, rr, g, Lp) in lanes], dyn_k=False, masked=True)
            if bscheme == 0:                          # masked B scheme: words (2j+s+g(rr))&7 of each 8-word block, bytes k = (j+i)&3
                for b8 in range(nsb * 4, nb8):
                    for s in range(2):
                        step([32 * b8 + 4 * ((2 * j + s + (((rr & 3) << 1) | (rr >> 2))) & 7) for (j, rr, g, Lp) in lanes], dyn_k=True, masked=True,
                             may_conflict=layout16)
            else:                                     # static B scheme: A-scheme code on the blocks, tails byte by byte
                for b8 in range(nsb * 4, nb8):
                    for t in range(2):
                        step([32 * b8 + 16 * ((t + rr) & 1) + 4 * j for (j, rr, g, Lp) in lanes], dyn_k=False, masked=False, may_conflict=True)
                for (j, rr, g, Lp) in lanes:
                    if Lp > 0 and j < (Lp & 3) and (Lp >> 2) >= 32 * nsb:
                        o = Lp & ~3
                        if P.byte(seqs[g][passoff + o + j], quals[g][passoff + o + j], o >> 2, j):
                            bad_reads.add(g)
        P.flush()
        assert P.conflicts == 0, "%d conflicting ATOMS instructions" % P.conflicts
        for key, v in P.glob.items():
            hist[key] = hist.get(key, 0) + v
    return hist, bad_reads


def direct_hist(seqs, quals, lens, q_offset):
    """the reference's accumulation, per base; a read with an illegal byte is 'bad' (its legal bases still count here)"""
    lo, hi = q_offset - 15, min(q_offset + 93, 127)
    hist, bad = {}, set()
    for g, L in enumerate(lens):
        for c in range(L):
            b, q = seqs[g][c], quals[g][c]
            if chr(b) not in "ACGTN" or q < lo or q > hi:
                bad.add(g)
                continue
            key = (c, "ACGTN".index(chr(b)), q - lo)
            hist[key] = hist.get(key, 0) + 1
    return hist, bad
