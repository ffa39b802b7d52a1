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
// Scores are kept scaled by 16 (match +16, mismatch -16, gap -80): the low 4 bits of a score are free, so
// "score*16 + (15 - y)" orders the cells of a column by (score desc, row asc) and ONE VIADDMNMX per cell
// maintains the column's first maximum.  The state per adapter row is g = score + gap, which is both the
// "left" candidate of the next column and the "up" candidate of the next row; the diagonal candidate is
// g(x-1,y-1) + (match score - gap), looked up for both reads by one PRMT from a 4-byte profile word per row.
// Origin flags (2 bits per cell and read: "up beat the diagonal", "left beat both") are accumulated per
// column in two packed words and stored; the reference's backtrace then runs over them.
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

// All packed values carry a bias of 0x4000 per half: they stay in (0, 0x8000), so a plain 32-bit add or subtract
// of two packed words never carries or borrows across the halves and the compiler is free to issue it on either
// integer pipe (IADD3 on the ALU pipe or IMAD on the FMA pipe) — the packed min/max (ALU pipe only) are the
// scarce resource of this kernel.
constexpr int BIAS = 0x4000;
constexpr uint32_t G0_2 = 0x3FB03FB0u;      // bias + gap * 16 (= -80) in both halves
constexpr uint32_t GAPSUB2 = 0x00500050u;   // subtracting it adds one gap to both halves
constexpr uint32_t SENT2 = 0u;              // the banned "left" candidate (reference: -100000.0f): below every score
constexpr uint32_t ONE2 = 0x00010001u;
constexpr uint32_t CM_INIT2 = 0u;
constexpr int BEST_INIT = 0;
constexpr int MSP_MATCH = 96, MSP_MISMATCH = 64;   // (+-1 - gap) * 16

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
// target_border[y] + gap, scaled (sequence_alignment.cpp:340-363): 0 for y <= 3, -5*(y-3) below
FXG_DPX_HD int border_g16(int y) { return BIAS + (y <= 3 ? 0 : -80 * (y - 3)) - 80; }

// One DP column for both reads.  `sel` = 0x0404 | code(read0) | code(read1) << 8 (PRMT selector: the profile
// byte of each read — bytes 4..7 of the operand pair — zero-extended to 16 bits by byte 0 of a zero word).  gp holds g(x-1, .) on entry and g(x, .) on exit.
template <int HMAX, bool BAN>
FXG_DPX_HD void column(int H, int x, uint32_t sel, const uint32_t (&prof)[HMAX], uint32_t (&gp)[HMAX], uint32_t &a1, uint32_t &a2,
                       uint32_t &cm)
{
    a1 = 0; a2 = 0; cm = CM_INIT2;
    uint32_t diag = G0_2;      // g(x-1, -1): query_border[x-1] (= target_border[-1] = 0 for x == 0) + gap
    uint32_t up = G0_2;        // g(x, -1):   query_border[x] + gap
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        const uint32_t msp = prmt(0u, prof[y], sel);                // profile word as the SECOND operand: it lives in a uniform register
        const uint32_t ul = diag + msp;                             // FROM_UPPER_LEFT candidate (plain add: see BIAS)
        uint32_t left = gp[y];                                      // FROM_LEFT candidate
        if (BAN && y > 3 && y - 3 > x) left = SENT2;                // sequence_alignment.cpp:388-390
        diag = gp[y];
        const uint32_t m2 = vmax2(ul, up);
        const uint32_t sc = vmax2(m2, left);
        // strict '>' in the reference's candidate order (diagonal, up, left); the differences are >= 0 per half
        const uint32_t t1 = vminu2(m2 - ul, ONE2);                  // 1: up > diagonal
        const uint32_t t2 = vminu2(sc - m2, ONE2);                  // 1: left > max(diagonal, up)
        a1 += t1 << y;
        a2 += t2 << y;
        const bool live = (y < HMAX - 3) || (y < H);                // only the last 3 rows can lie beyond the adapter
        if (live) cm = vaddmax2(sc, pack2(15 - y, 15 - y), cm);     // first maximum of the column, row in the low 4 bits
        up = sc - GAPSUB2;
        gp[y] = up;
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
    uint32_t prof[HMAX], gp[HMAX];
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        prof[y] = profile_word(y < H ? (uint32_t)adapter[y] : 0u);
        gp[y] = pack2(border_g16(y), border_g16(y));
    }
    uint32_t org1[MAXW], org2[MAXW];
    int best0 = BEST_INIT, best1 = BEST_INIT, bx0 = 0, bx1 = 0;
    uint32_t exact = 0;
    for (int x0 = 0; x0 < L; x0 += 4) {
        const uint32_t w0 = ld32(row0 + x0), w1 = ld32(row1 + x0);
        const int nb = (L - x0 < 4) ? (L - x0) : 4;
        const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
        {   // anything but A/C/G/T (SWAR table lookup by the low 3 bits, as fxg_device.cuh seq_bad_bits)
            const uint32_t y0 = w0 & 0x07070707u, y1 = w1 & 0x07070707u;
            const uint32_t s0 = prmt(y0 | (y0 >> 4), 0u, 0x4420u), s1 = prmt(y1 | (y1 >> 4), 0u, 0x4420u);
            if ((w0 ^ prmt(DV_LO, DV_HI, s0)) & m) exact |= 1u;
            if ((w1 ^ prmt(DV_LO, DV_HI, s1)) & m) exact |= 2u;
        }
        const uint32_t c0 = (w0 >> 1) & 0x03030303u, c1 = (w1 >> 1) & 0x03030303u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < nb) {
                const int x = x0 + k;
                const uint32_t sel = (prmt(c0, c1, (uint32_t)(k | ((4 + k) << 4))) & 0x0303u) | 0x0404u;
                uint32_t a1, a2, cm;
                if (x < HMAX - 4) column<HMAX, true>(H, x, sel, prof, gp, a1, a2, cm);
                else column<HMAX, false>(H, x, sel, prof, gp, a1, a2, cm);
                org1[x] = a1; org2[x] = a2;
                // first maximum in (x outer, y inner) order: a later column wins only with a strictly larger score
                const int cm0 = (int)(cm & 0xFFFFu), cm1 = (int)(cm >> 16);
                if (cm0 > (best0 | 15)) { best0 = cm0; bx0 = x; }
                if (cm1 > (best1 | 15)) { best1 = cm1; bx1 = x; }
            }
        }
    }
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
            if ((org2[qi] >> pos) & 1u) { gaps++; qi--; }                 // FROM_LEFT
            else if ((org1[qi] >> pos) & 1u) { gaps++; ti--; }            // FROM_UPPER
            else {                                                        // FROM_UPPER_LEFT
                if (row[qi] == adapter[ti]) matches++; else mism++;
                qi--; ti--;
            }
        }
        out.lo[r] = (uint32_t)matches | ((uint32_t)mism << 7) | ((uint32_t)tstart << 21);
        out.hi[r] = (uint32_t)gaps | ((uint32_t)qstart << 15);
    }
}

}  // namespace dpx
}  // namespace fxg
