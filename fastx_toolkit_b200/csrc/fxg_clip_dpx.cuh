// fxg_clip_dpx.cuh — K-CLIP-ALIGN, integer fast path: two reads per thread in packed s16x2, built on the
// DPX instructions of sm_90+/sm_100a (VIADD.16x2, VIMNMX.S16x2, VIADDMNMX.S16x2).
//
// Restates HalfLocalSequenceAlignment::populate_matrix + find_optimal_alignment_from_point
// (src/libfastx/sequence_alignment.cpp:340-428, 496-604) for reads AND adapters without 'N': then every
// score of the reference is a sum of {+1, -1, -5, 0} — an integer that fp32 represents exactly — so integer
// arithmetic takes the same branches as the reference's float compares, bit for bit.  Reads with an 'N' (score
// 0.1f, whose fp32 rounding decides ties) or an illegal character are NOT handled here: they are reported
// back (`exact`) and go through the fp32 kernel (fxg_clip.cu, k_clip_bits).
//
// Representation.  A cell holds F(x,y) = 32*score + 16*(x+y) + BIAS in one s16 half (score = the reference's
// integer score; match +32, mismatch -32, gap -160 after scaling).  The 16*(x+y) term makes the three candidates
//     diagonal: F(x-1,y-1) + {64 match, 0 mismatch}        up / left: F(.,.) - 144
// so that the diagonal needs one add of a profile byte (looked up for both reads by ONE PRMT from a 4-byte
// profile word per adapter row) and the other two share one packed max and one fused add-max (VIADDMNMX).
// BIAS keeps every half in (0, 0x8000): plain 32-bit adds/subtracts of packed words cannot carry across the
// halves, so the compiler may issue them on either integer pipe — the packed min/max run on the ALU pipe only,
// which is what bounds this kernel.  F - 16y + (15 - y) orders the cells of a column by (score desc, row asc), so
// one more VIADDMNMX per cell maintains the column's first maximum with its row in the low 4 bits.
// Origin flags (2 bits per cell and read: "the diagonal lost", "left beat up") are accumulated per column in
// two packed words and stored; the reference's backtrace then runs over them.
//
// The same source compiles for the host (portable emulation of the packed ops): tests/ runs it on the CPU
// against the oracle, so the algorithm is checked without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FXG_DPX_HD __host__ __device__ __forceinline__
#else
#define FXG_DPX_HD static inline
#endif

namespace fxg {
namespace dpx {

constexpr int BIAS = 0x4000;
constexpr uint32_t GAPADD2 = 0xFF70FF70u;   // -144 = 16 - 160 in both halves: one gap, one step of x+y
constexpr uint32_t SENT2 = 0u;              // the banned "left" candidate (reference: -100000.0f): below every cell
constexpr uint32_t ONE2 = 0x00010001u;
constexpr uint32_t CM_INIT2 = 0u;
constexpr int BEST_INIT = 0;
constexpr int MSP_MATCH = 64, MSP_MISMATCH = 0;    // 32 * (+-1) + 32

FXG_DPX_HD uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

#if defined(__CUDA_ARCH__)
FXG_DPX_HD uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
FXG_DPX_HD uint32_t vsub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
FXG_DPX_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
FXG_DPX_HD uint32_t vminu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
FXG_DPX_HD uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
// max per half; ge_lo / ge_hi = (a >= b) per half (one VIMNMX.S16x2 with two predicate outputs)
FXG_DPX_HD uint32_t vbmax2(uint32_t a, uint32_t b, bool &ge_hi, bool &ge_lo) { return __vibmax_s16x2(a, b, &ge_hi, &ge_lo); }
FXG_DPX_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
FXG_DPX_HD uint32_t ld32(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
#else
FXG_DPX_HD int16_t h_lo(uint32_t a) { return (int16_t)(a & 0xFFFFu); }
FXG_DPX_HD int16_t h_hi(uint32_t a) { return (int16_t)(a >> 16); }
FXG_DPX_HD uint32_t vadd2(uint32_t a, uint32_t b) { return pack2(h_lo(a) + h_lo(b), h_hi(a) + h_hi(b)); }
FXG_DPX_HD uint32_t vsub2(uint32_t a, uint32_t b) { return pack2(h_lo(a) - h_lo(b), h_hi(a) - h_hi(b)); }
FXG_DPX_HD uint32_t vmax2(uint32_t a, uint32_t b) { return pack2(h_lo(a) > h_lo(b) ? h_lo(a) : h_lo(b), h_hi(a) > h_hi(b) ? h_hi(a) : h_hi(b)); }
FXG_DPX_HD uint32_t vminu2(uint32_t a, uint32_t b)
{
    const uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
FXG_DPX_HD uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vadd2(a, b), c); }
FXG_DPX_HD uint32_t vbmax2(uint32_t a, uint32_t b, bool &ge_hi, bool &ge_lo)
{
    ge_lo = h_lo(a) >= h_lo(b); ge_hi = h_hi(a) >= h_hi(b);
    return vmax2(a, b);
}
FXG_DPX_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t nib = (s >> (4 * i)) & 0xFu;
        const uint32_t src = (nib & 4u) ? b : a;
        uint32_t v = (src >> (8 * (nib & 3u))) & 0xFFu;
        if (nib & 8u) v = (v & 0x80u) ? 0xFFu : 0u;
        r |= v << (8 * i);
    }
    return r;
}
FXG_DPX_HD uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
#endif

// profile word of adapter character tc: byte c = (base(c) == tc ? match : mismatch) - gap, scaled; the base
// code of a read character ch is (ch >> 1) & 3: A -> 0, C -> 1, T -> 2, G -> 3
FXG_DPX_HD uint32_t profile_word(uint32_t tc)
{
    const uint32_t a = tc == 'A' ? MSP_MATCH : MSP_MISMATCH, c = tc == 'C' ? MSP_MATCH : MSP_MISMATCH;
    const uint32_t t = tc == 'T' ? MSP_MATCH : MSP_MISMATCH, g = tc == 'G' ? MSP_MATCH : MSP_MISMATCH;
    return a | (c << 8) | (t << 16) | (g << 24);
}
// F(-1, y): target_border[y] (sequence_alignment.cpp:340-363: 0 for y <= 3, -5*(y-3) below) in the virtual column x = -1
FXG_DPX_HD int border_f(int y) { return BIAS + 32 * (y <= 3 ? 0 : -5 * (y - 3)) + 16 * (y - 1); }

// One DP column for both reads.  `sel` = 0x0404 | code(read0) | code(read1) << 8 (PRMT selector: the profile
// byte of each read — bytes 4..7 of the operand pair — zero-extended to 16 bits by byte 0 of a zero word).
// fp holds F(x-1, .) on entry and F(x, .) on exit; `qb` = F(x-1, -1) in both halves (query_border = 0).
template <int HMAX, bool BAN>
FXG_DPX_HD void column(int H, int x, uint32_t sel, uint32_t qb, const uint32_t (&prof)[HMAX], uint32_t (&fp)[HMAX], uint32_t &a1,
                       uint32_t &a2, uint32_t &cm)
{
    a1 = 0; a2 = 0; cm = CM_INIT2;
    uint32_t diag = qb;                    // F(x-1, -1)   (x == 0: target_border[-1], 0.0 in practice)
    uint32_t up = qb + 0x00100010u;        // F(x, -1)
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        const uint32_t msp = prmt(0u, prof[y], sel);                // profile word as the SECOND operand: it lives in a uniform register
        const uint32_t ul = diag + msp;                             // FROM_UPPER_LEFT candidate
        uint32_t left = fp[y];
        if (BAN && y > 3 && y - 3 > x) left = SENT2;                // sequence_alignment.cpp:388-390
        diag = fp[y];
        const uint32_t m = vmax2(up, left);
        const uint32_t sc = vaddmax2(m, GAPADD2, ul);               // max(diagonal, max(up, left) + gap)
        // strict '>' in the reference's candidate order (diagonal, up, left): the diagonal keeps ties, up keeps
        // ties against left; both differences are >= 0 per half
        const uint32_t e1 = vminu2(sc - ul, ONE2);                  // 1: up or left beat the diagonal
        const uint32_t e2 = vminu2(m - up, ONE2);                   // 1: left > up
        a1 += e1 << y;
        a2 += e2 << y;
        const bool live = (y < HMAX - 3) || (y < H);                // only the last 3 rows can lie beyond the adapter
        if (live) cm = vaddmax2(sc, pack2(15 - 17 * y, 15 - 17 * y), cm);   // column's first maximum, row in the low 4 bits
        up = sc;
        fp[y] = sc;
    }
}

struct PairOut {
    int bx[2], by[2];          // best cell (query_end, target_end)
    uint32_t lo[2], hi[2];     // payload words as in fxg_clip.cu: matches | mism<<7 | neutral<<14 | tstart<<21 ; gaps | qstart<<15
    uint32_t exact;            // bit r set: read r has an 'N' or an illegal character -> fp32 kernel
};

// V2 table of fxg_stats.cu: legal characters by low-3-bit code with 'N' poisoned
constexpr uint32_t DV_LO = 0x43FF41FFu, DV_HI = 0x47FFFF54u;

// Both reads have length L (1 <= L <= MAXW), the adapter has H <= HMAX <= 16 characters, none of them 'N'.
template <int HMAX, int MAXW>
FXG_DPX_HD void align_pair(const uint8_t *row0, const uint8_t *row1, int L, const uint8_t *adapter, int H, PairOut &out)
{
    uint32_t prof[HMAX], fp[HMAX];
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        prof[y] = profile_word(y < H ? (uint32_t)adapter[y] : 0u);
        fp[y] = pack2(border_f(y), border_f(y));
    }
    uint32_t org1[MAXW], org2[MAXW];
    int best0 = BEST_INIT, best1 = BEST_INIT, bx0 = 0, bx1 = 0;
    uint32_t exact = 0;
    // anything but A/C/G/T (SWAR table lookup by the low 3 bits, as fxg_device.cuh seq_bad_bits) -> exact path
    for (int x0 = 0; x0 < L; x0 += 4) {
        const uint32_t w0 = ld32(row0 + x0), w1 = ld32(row1 + x0);
        const int nb = L - x0;
        const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
        const uint32_t y0 = w0 & 0x07070707u, y1 = w1 & 0x07070707u;
        const uint32_t s0 = prmt(y0 | (y0 >> 4), 0u, 0x4420u), s1 = prmt(y1 | (y1 >> 4), 0u, 0x4420u);
        if ((w0 ^ prmt(DV_LO, DV_HI, s0)) & m) exact |= 1u;
        if ((w1 ^ prmt(DV_LO, DV_HI, s1)) & m) exact |= 2u;
    }
    uint32_t qb = pack2(BIAS - 32, BIAS - 32);          // F(-1, -1)
    // first maximum in (x outer, y inner) order: a later column wins only with a strictly larger score
#define FXG_DPX_COLUMN_END(x)                                                                   \
    do {                                                                                        \
        org1[x] = a1; org2[x] = a2;                                                             \
        const int cm0 = (int)(cm & 0xFFFFu) - 16 * (x), cm1 = (int)(cm >> 16) - 16 * (x);       \
        if (cm0 > (best0 | 15)) { best0 = cm0; bx0 = (x); }                                     \
        if (cm1 > (best1 | 15)) { best1 = cm1; bx1 = (x); }                                     \
        qb += 0x00100010u;                                                                      \
    } while (0)
    // the first HMAX-4 columns carry the "left" ban (y - 3 > x)
    const int xban = L < HMAX - 4 ? L : HMAX - 4;
    for (int x = 0; x < xban; x++) {
        const uint32_t sel = (((uint32_t)row0[x] >> 1) & 3u) | ((((uint32_t)row1[x] >> 1) & 3u) << 8) | 0x0404u;
        uint32_t a1, a2, cm;
        column<HMAX, true>(H, x, sel, qb, prof, fp, a1, a2, cm);
        FXG_DPX_COLUMN_END(x);
    }
    // the rest, one 4-byte word of each read at a time (the next words are requested before the current ones are used)
    uint32_t w0n = 0, w1n = 0;
    if (xban < L) { w0n = ld32(row0 + xban); w1n = ld32(row1 + xban); }
    for (int x0 = xban; x0 < L; x0 += 4) {
        const uint32_t w0 = w0n, w1 = w1n;
        if (x0 + 4 < L) { w0n = ld32(row0 + x0 + 4); w1n = ld32(row1 + x0 + 4); }
        const int nb = (L - x0 < 4) ? (L - x0) : 4;
        uint32_t c0 = (w0 >> 1) & 0x03030303u, c1 = (w1 >> 1) & 0x03030303u;
#pragma unroll 1
        for (int k = 0; k < nb; k++) {
            const int x = x0 + k;
            const uint32_t sel = (c0 & 0x03u) | ((c1 & 0x03u) << 8) | 0x0404u;
            c0 >>= 8; c1 >>= 8;
            uint32_t a1, a2, cm;
            column<HMAX, false>(H, x, sel, qb, prof, fp, a1, a2, cm);
            FXG_DPX_COLUMN_END(x);
        }
    }
#undef FXG_DPX_COLUMN_END
    out.exact = exact;
    out.bx[0] = bx0; out.bx[1] = bx1;
    out.by[0] = 15 - (best0 & 15); out.by[1] = 15 - (best1 & 15);
    // backtrace (find_optimal_alignment_from_point, sequence_alignment.cpp:496-604)
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint8_t *row = r ? row1 : row0;
        int qi = out.bx[r], ti = out.by[r];
        int matches = 0, mism = 0, gaps = 0, qstart = qi, tstart = ti;
        while (qi >= 0 && ti >= 0) {
            qstart = qi; tstart = ti;
            const int pos = ti + 16 * r;
            if (!((org1[qi] >> pos) & 1u)) {                              // FROM_UPPER_LEFT: the diagonal kept the cell
                if (row[qi] == adapter[ti]) matches++; else mism++;
                qi--; ti--;
            }
            else if ((org2[qi] >> pos) & 1u) { gaps++; qi--; }            // FROM_LEFT: left > up (and > diagonal)
            else { gaps++; ti--; }                                        // FROM_UPPER
        }
        out.lo[r] = (uint32_t)matches | ((uint32_t)mism << 7) | ((uint32_t)tstart << 21);
        out.hi[r] = (uint32_t)gaps | ((uint32_t)qstart << 15);
    }
}

}  // namespace dpx
}  // namespace fxg
