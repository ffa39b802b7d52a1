// fxg_stats.cu — K-STATS: per-cycle x nucleotide x quality histograms for fastx_quality_stats
// (src/fastx_quality_stats/fastx_quality_stats.c:166-216 `read_file`).  Everything the tool prints
// (count/min/max/sum/mean/quartiles/whiskers, per nucleotide) derives from hist[cycle][nuc][q+15]
// (SURVEY.md Appendix A.5), so the device only counts; the host derives and prints.
//
// Fast kernel: one persistent CTA per SM owns a shared-memory histogram (u32, A/C/G/T x q' < 64 for up
// to 160 cycles); each warp streams its own tiles of 32 reads through a private TMA ring; lane = read,
// and at step t lane l works on 4-byte word (t + l) mod nwords of its read, so the 32 lanes of one
// ATOMS instruction touch 32 different cycles (no same-address serialisation) and, thanks to the padded
// word-block pitch, 32 different banks when their quality values agree.  Per sample the cost is
// PRMT (assemble nuc<<8 | q'<<2) + IADD + ATOMS; validation and the nucleotide lookup are SWAR per word.
// 'N' bases and q' >= 64 (rare) go straight to the global u64 histogram.
#include "fxg_kernels.cuh"

namespace fxg {

// nucleotide index table for PRMT, by the low 3 bits of the base: A(1)->0 C(3)->1 G(7)->2 T(4)->3 N(6)->4
constexpr uint32_t NLUT_LO = 0x01800080u;   // codes 0..3: -,A,-,C   (0x80 marks "not a base")
constexpr uint32_t NLUT_HI = 0x02048003u;   // codes 4..7: T,-,N,G

__device__ __forceinline__ void hist_global_add(unsigned long long *hist, int max_cycles, int cycle, int nuc, int qp,
                                                unsigned long long w)
{
    if (cycle < max_cycles) atomicAdd(&hist[((size_t)cycle * 5 + nuc) * 109 + qp], w);
}

// one 4-byte word (4 consecutive cycles) of one read
template <bool TAIL>
__device__ __forceinline__ void stats_word(const StatsParams &P, const QualK &qk, uint32_t sw, uint32_t qw, int wi, int remb,
                                           uint32_t hs_addr, uint32_t &bads, uint32_t &badq)
{
    const uint32_t m = TAIL ? head_mask(remb) : 0xFFFFFFFFu;
    const uint32_t sel = base_selector(sw);
    const uint32_t wbad_s = (sw ^ __byte_perm(VLUT_LO, VLUT_HI, sel)) & m;
    const uint32_t nuc4 = __byte_perm(NLUT_LO, NLUT_HI, sel);
    const uint32_t wbad_q = qual_bad_bits(qw, qw | HI, qk) & HI & m;
    bads |= wbad_s;
    badq |= wbad_q;
    const int rel = wi - P.w0;
    if ((wbad_s | wbad_q) != 0 || rel < 0 || rel >= P.nw) return;
    const uint32_t qp4 = qw - qk.lo4;                     // q+15 per byte (legal bytes: no borrow)
    const uint32_t blk = hs_addr + (uint32_t)rel * ST_WBLK;
    if (!TAIL && ((qp4 & 0xC0C0C0C0u) | (nuc4 & 0xFCFCFCFCu)) == 0) {
        const uint32_t qs4 = qp4 << 2;
        // offset = q'*4 + nuc*256: byte0 <- qs4.k, byte1 <- nuc4.k, bytes 2,3 <- sign(nuc4.k) = 0
        const uint32_t o0 = prmt_raw(qs4, nuc4, 0xCC40u), o1 = prmt_raw(qs4, nuc4, 0xDD51u);
        const uint32_t o2 = prmt_raw(qs4, nuc4, 0xEE62u), o3 = prmt_raw(qs4, nuc4, 0xFF73u);
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(blk + o0) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(blk + o1 + ST_KBLK) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(blk + o2 + 2 * ST_KBLK) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(blk + o3 + 3 * ST_KBLK) : "memory");
    } else {
        const int nk = TAIL ? remb : 4;
        for (int k = 0; k < nk; k++) {                    // tail bytes, 'N', or q' >= 64
            const uint32_t nuc = (nuc4 >> (8 * k)) & 0xFFu, qp = (qp4 >> (8 * k)) & 0xFFu;
            if (nuc < 4u && qp < (uint32_t)ST_QWIN)
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(blk + k * ST_KBLK + nuc * 256u + qp * 4u) : "memory");
            else
                hist_global_add(P.hist, P.max_cycles, 4 * wi + k, (int)nuc, (int)qp, 1ull);
        }
    }
}

// G lanes per read (lane j takes words j, j+G, ...), 6*G warps per CTA: the tiles of all warps together always
// hold 192 reads, so more lanes per read means more resident warps (better latency hiding) for the same smem.
template <int G>
__global__ void __launch_bounds__(ST_WARPS * G * 32, 1) k_stats(const __grid_constant__ StatsParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int WARPS = ST_WARPS * G;
    constexpr int NTHREADS = WARPS * 32;
    __shared__ __align__(8) uint64_t full_bar[WARPS][2];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int S = P.stride, stages = P.stages, R = P.tile_reads;
    const uint32_t hist_bytes = (uint32_t)P.nw * ST_WBLK;
    uint8_t *hs = smem;                                              // shared histogram
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    const uint32_t stage_bytes = slab_bytes * 2u;
    uint8_t *wbase = smem + ((hist_bytes + 127u) & ~127u) + (size_t)w * stages * stage_bytes;
    uint64_t *bars = full_bar[w];
    const int64_t ntiles = (P.n + R - 1) / R;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + w, GW = (int64_t)gridDim.x * WARPS;
    const QualK qk = P.qk;

    for (uint32_t i = tid * 4; i < hist_bytes; i += NTHREADS * 4) *reinterpret_cast<uint32_t *>(hs + i) = 0u;
    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int s) {
        const int64_t r0 = tile * R;
        const int64_t left = P.n - r0;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * (uint32_t)S;
        uint8_t *dst = wbase + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&bars[s], bytes * 2u);
        bulk_g2s(dst, P.seq + r0 * S, bytes, &bars[s]);
        bulk_g2s(dst + slab_bytes, P.qual + r0 * S, bytes, &bars[s]);
    };
    if (lane == 0) {
        for (int i = 0; i < stages; i++) {
            const int64_t t = gw + (int64_t)i * GW;
            if (t < ntiles) issue(t, i);
        }
    }

    const uint32_t hs_addr = smem_u32(hs);
    const int j = lane & (G - 1), rr = lane / G;
    int s = 0;
    uint32_t parity = 0;

    for (int64_t tile = gw; tile < ntiles; tile += GW) {
        mbar_wait(&bars[s], parity);
        const int64_t g = tile * R + rr;
        const bool active = rr < R && g < P.n;
        int L = 0;
        if (active) L = P.len ? __ldg(P.len + g) : P.uniform_len;
        const bool lenbad = active && (L <= 0 || L > S);
        if (lenbad) L = 0;
        const uint8_t *srow = wbase + (size_t)s * stage_bytes + (size_t)(active ? rr : 0) * S;
        const uint8_t *qrow = srow + slab_bytes;
        const int nwf = L >> 2;                                  // full words of this read
        const int nk = nwf > j ? (nwf - j + G - 1) / G : 0;      // words of this lane: j, j+G, ...
        uint32_t bads = 0, badq = 0;

        // skewed start so that the lanes of one instruction work on different cycles; next word prefetched
        int k = nk > 0 ? rr % nk : 0;
        uint32_t sw = 0, qw = 0;
        if (nk > 0) {
            sw = *reinterpret_cast<const uint32_t *>(srow + 4 * (G * k + j));
            qw = *reinterpret_cast<const uint32_t *>(qrow + 4 * (G * k + j));
        }
        for (int t = 0; t < nk; t++) {
            const int wi = G * k + j;
            if (++k == nk) k = 0;
            const int wn = G * k + j;
            const uint32_t sw_n = *reinterpret_cast<const uint32_t *>(srow + 4 * wn);
            const uint32_t qw_n = *reinterpret_cast<const uint32_t *>(qrow + 4 * wn);
            stats_word<false>(P, qk, sw, qw, wi, 0, hs_addr, bads, badq);
            sw = sw_n; qw = qw_n;
        }
        // trailing 1..3 bases (owned by the lane whose turn it would be)
        const int remb = L & 3;
        if (remb && (nwf & (G - 1)) == j) {
            const uint32_t tsw = *reinterpret_cast<const uint32_t *>(srow + 4 * nwf);
            const uint32_t tqw = *reinterpret_cast<const uint32_t *>(qrow + 4 * nwf);
            stats_word<true>(P, qk, tsw, tqw, nwf, remb, hs_addr, bads, badq);
        }
        if (((bads | badq) != 0 || lenbad) && active)
            atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));

        __syncwarp();
        if (lane == 0) {
            const int64_t nt = tile + (int64_t)stages * GW;
            if (nt < ntiles) issue(nt, s);
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }

    // flush the CTA's shared histogram into the global u64 table
    __syncthreads();
    const int bins = P.nw * 4 * 4 * ST_QWIN;
    for (int i = tid; i < bins; i += NTHREADS) {
        const int qp = i & (ST_QWIN - 1), nuc = (i >> 6) & 3, k = (i >> 8) & 3, rel = i >> 10;
        const uint32_t v = *reinterpret_cast<const uint32_t *>(hs + (size_t)rel * ST_WBLK + k * ST_KBLK + nuc * 256 + qp * 4);
        if (v) hist_global_add(P.hist, P.max_cycles, 4 * (P.w0 + rel) + k, nuc, qp, (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------------
// K-STATS, second generation: bank-conflict-free shared histogram.
//
// Layout: hist[bin = nuc*64 + q'][pc] u32, 160 physical columns per bin (pitch 640 B), where the column of
// cycle 4w+k (w = 4-byte word of the read, k = byte in the word) is pc = 40k + w.  The bank of a counter is
// therefore (8k + w) mod 32 — a function of the CYCLE only, never of the data — so an ATOMS whose 32 lanes
// sit on 32 columns with distinct (8k + w) mod 32 is one wavefront whatever the qualities are.
// A warp owns a tile of 8 reads, 4 lanes per read (lane = 4*rr + j):
//   * A scheme, 32-word superblocks: at step t lane (rr,j) takes word 4*((t+rr)&7) + j — the 32 lanes cover
//     32 different words, k is the same for all lanes of one ATOMS (static, address immediate);
//   * B scheme, 8-word blocks (the words past the last full superblock): at step s lane (rr,j) takes word
//     (2j+s+g(rr))&7 of the block and issues its four bytes in the rotated order k = (j+i)&3, so the four lanes
//     that share a word use four different k.
// Per word: 3 ops for the PRMT selector, 2 PRMT (legal-character table with 'N' poisoned, nuc<<6 table), one
// IADD3 (nuc<<6 | q'), 2 LOP3 for "all four samples are plain A/C/G/T with 0 <= q' < 64" (this subsumes the
// quality range check when Q-15 <= 64); per sample: byte extract, IMAD (bin*640 + column), ATOMS.
// Words with 'N', q' >= 64 or illegal bytes take a per-byte path; the 1..3 bytes past the last full word are
// done once per tile with one lane per byte.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t V2LUT_HI = 0x47FFFF54u;      // VLUT_HI with 'N' (code 6) poisoned: N goes to the per-byte path
constexpr uint32_t N6_LO = 0x40000000u;         // nuc<<6 by code: 1:A->00, 3:C->40
constexpr uint32_t N6_HI = 0x800000C0u;         //                 4:T->C0, 7:G->80

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void reds_inc(uint32_t addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// one base (exact path): validation as the reader does it, 'N' and q' >= 64 go to the global table.
// Returns 1 when the base or its quality is illegal.
__device__ __forceinline__ uint32_t stats2_byte(const StatsParams &P, uint32_t c, uint32_t q, int wrel, int k, uint32_t hs_addr)
{
    const uint32_t lo = P.qk.lo4 & 0xFFu, hmax = 127u - (P.qk.hik4 & 0xFFu) - lo;
    const uint32_t code = c & 7u;
    const uint32_t legal = __byte_perm(VLUT_LO, VLUT_HI, code) & 0xFFu;
    const uint32_t nuc = __byte_perm(NLUT_LO, NLUT_HI, code) & 0xFFu;
    const uint32_t qp = q - lo;
    if (legal != c || qp > hmax) return 1u;
    if (nuc < 4u && qp < (uint32_t)ST_QWIN)
        reds_inc(hs_addr + (nuc * 64u + qp) * S2_PITCH + (uint32_t)k * (4u * ST_MAXW) + 4u * (uint32_t)wrel);
    else
        hist_global_add(P.hist, P.max_cycles, 4 * (P.w0 + wrel) + k, (int)nuc, (int)qp, 1ull);
    return 0u;
}
// per-byte path of one full word (rare: 'N', q' >= 64 or an illegal byte somewhere in the word)
__device__ __noinline__ uint32_t stats2_slow_word(const StatsParams &P, uint32_t sw, uint32_t qw, int wrel, int nbytes, uint32_t hs_addr)
{
    uint32_t bad = 0;
    for (int k = 0; k < nbytes; k++) bad |= stats2_byte(P, (sw >> (8 * k)) & 0xFFu, (qw >> (8 * k)) & 0xFFu, wrel, k, hs_addr);
    return bad;
}

// loop-invariant operands kept in registers (opaque to the compiler, which would otherwise rebuild the
// 32-bit immediates in front of every PRMT)
struct Stats2K {
    uint32_t vlut_lo, n6_lo, neg_lo4;
};

// decode one full word: comb = (nuc<<6 | q') per byte, valid when the returned test word is 0 (four plain
// A/C/G/T bases with 0 <= q' < 64)
__device__ __forceinline__ uint32_t stats2_decode(const Stats2K &K, uint32_t sw, uint32_t qw, uint32_t &comb)
{
    const uint32_t y = sw & 0x07070707u;
    const uint32_t sel = prmt_raw(y | (y >> 4), 0u, 0x4420u);
    const uint32_t e = prmt_raw(K.vlut_lo, V2LUT_HI, sel);
    const uint32_t n6 = prmt_raw(K.n6_lo, N6_HI, sel);
    comb = n6 + qw + K.neg_lo4;
    return (sw ^ e) | ((comb ^ n6) & 0xC0C0C0C0u);
}
// four counter increments of one decoded word at byte offset o = 4*wrel of the pass window
template <bool DYNK>
__device__ __forceinline__ void stats2_emit(uint32_t comb, uint32_t o, uint32_t hs_addr, const uint32_t (&ksel)[4], const uint32_t (&koff)[4])
{
    if (!DYNK) {
        const uint32_t b0 = comb & 0xFFu, b1 = prmt_raw(comb, 0u, 0x4441u), b2 = prmt_raw(comb, 0u, 0x4442u), b3 = comb >> 24;
        const uint32_t col = hs_addr + o;
        reds_inc(b0 * S2_PITCH + col);
        reds_inc(b1 * S2_PITCH + col + 1u * (4u * ST_MAXW));
        reds_inc(b2 * S2_PITCH + col + 2u * (4u * ST_MAXW));
        reds_inc(b3 * S2_PITCH + col + 3u * (4u * ST_MAXW));
    } else {
        const uint32_t col = hs_addr + o;
#pragma unroll
        for (int i = 0; i < 4; i++) reds_inc(prmt_raw(comb, 0u, ksel[i]) * S2_PITCH + (col + koff[i]));
    }
}
template <bool DYNK>
__device__ __forceinline__ void stats2_word(const StatsParams &P, const Stats2K &K, uint32_t sw, uint32_t qw, uint32_t o, uint32_t hs_addr,
                                            const uint32_t (&ksel)[4], const uint32_t (&koff)[4], uint32_t &bad)
{
    uint32_t comb;
    if (stats2_decode(K, sw, qw, comb) == 0u) stats2_emit<DYNK>(comb, o, hs_addr, ksel, koff);
    else bad |= stats2_slow_word(P, sw, qw, (int)(o >> 2), 4, hs_addr);
}
// two full words with one branch between them
template <bool DYNK>
__device__ __forceinline__ void stats2_pair(const StatsParams &P, const Stats2K &K, uint32_t sw0, uint32_t qw0, uint32_t o0, uint32_t sw1,
                                            uint32_t qw1, uint32_t o1, uint32_t hs_addr, const uint32_t (&ksel)[4],
                                            const uint32_t (&koff)[4], uint32_t &bad)
{
    uint32_t c0, c1;
    const uint32_t t0 = stats2_decode(K, sw0, qw0, c0), t1 = stats2_decode(K, sw1, qw1, c1);
    if ((t0 | t1) == 0u) {
        stats2_emit<DYNK>(c0, o0, hs_addr, ksel, koff);
        stats2_emit<DYNK>(c1, o1, hs_addr, ksel, koff);
    } else {
        if (t0 == 0u) stats2_emit<DYNK>(c0, o0, hs_addr, ksel, koff);
        else bad |= stats2_slow_word(P, sw0, qw0, (int)(o0 >> 2), 4, hs_addr);
        if (t1 == 0u) stats2_emit<DYNK>(c1, o1, hs_addr, ksel, koff);
        else bad |= stats2_slow_word(P, sw1, qw1, (int)(o1 >> 2), 4, hs_addr);
    }
}

// B scheme: two words per lane, each with vb = 0..4 (or more) valid bytes.  Bytes past the end of the read are
// replaced by a plain 'A' with q' = 0 so that the packed test still decides, and are simply not counted (their
// ATOMS is predicated off); vb <= 0 switches the whole word off.  One code path for full words, the last 1..3
// bases of a read and the words beyond it.
__device__ __forceinline__ void stats2_emit_masked(uint32_t comb, uint32_t o, int vb, uint32_t hs_addr, const uint32_t (&ksel)[4],
                                                   const uint32_t (&koff)[4])
{
    const uint32_t col = hs_addr + o;
    const uint32_t dummy = hs_addr + (uint32_t)S2_HIST_BYTES + ((o >> 2) & 31u) * 4u;   // a counter nobody reads (branch-free masking)
#pragma unroll
    for (int i = 0; i < 4; i++) {        // byte k = ksel & 3 counts iff k < vb
        const uint32_t addr = prmt_raw(comb, 0u, ksel[i]) * S2_PITCH + (col + koff[i]);
        reds_inc((int)(ksel[i] & 3u) < vb ? addr : dummy);
    }
}
__device__ __forceinline__ void stats2_pair_masked(const StatsParams &P, const Stats2K &K, uint32_t sw0, uint32_t qw0, uint32_t o0, int vb0,
                                                   uint32_t sw1, uint32_t qw1, uint32_t o1, int vb1, uint32_t hs_addr,
                                                   const uint32_t (&ksel)[4], const uint32_t (&koff)[4], uint32_t &bad)
{
    const uint32_t m0 = head_mask(vb0), m1 = head_mask(vb1), lo4 = 0u - K.neg_lo4;
    uint32_t c0, c1;
    const uint32_t t0 = stats2_decode(K, (sw0 & m0) | (0x41414141u & ~m0), (qw0 & m0) | (lo4 & ~m0), c0);
    const uint32_t t1 = stats2_decode(K, (sw1 & m1) | (0x41414141u & ~m1), (qw1 & m1) | (lo4 & ~m1), c1);
    if ((t0 | t1) == 0u) {
        stats2_emit_masked(c0, o0, vb0, hs_addr, ksel, koff);
        stats2_emit_masked(c1, o1, vb1, hs_addr, ksel, koff);
    } else {
        if (t0 == 0u) stats2_emit_masked(c0, o0, vb0, hs_addr, ksel, koff);
        else bad |= stats2_slow_word(P, sw0, qw0, (int)(o0 >> 2), vb0 < 4 ? vb0 : 4, hs_addr);
        if (t1 == 0u) stats2_emit_masked(c1, o1, vb1, hs_addr, ksel, koff);
        else bad |= stats2_slow_word(P, sw1, qw1, (int)(o1 >> 2), vb1 < 4 ? vb1 : 4, hs_addr);
    }
}

// A scheme, ragged tile: the word that holds the last 1..3 bases of a read (byte order k = 0..3 as in the A scheme)
__device__ __forceinline__ void stats2_word_masked(const StatsParams &P, const Stats2K &K, uint32_t sw, uint32_t qw, uint32_t o, int vb,
                                                   uint32_t hs_addr, uint32_t &bad)
{
    const uint32_t ksel_s[4] = { 0x4440u, 0x4441u, 0x4442u, 0x4443u };
    const uint32_t koff_s[4] = { 0u, 4u * ST_MAXW, 8u * ST_MAXW, 12u * ST_MAXW };
    const uint32_t m = head_mask(vb), lo4 = 0u - K.neg_lo4;
    uint32_t c;
    if (stats2_decode(K, (sw & m) | (0x41414141u & ~m), (qw & m) | (lo4 & ~m), c) == 0u) stats2_emit_masked(c, o, vb, hs_addr, ksel_s, koff_s);
    else bad |= stats2_slow_word(P, sw, qw, (int)(o >> 2), vb, hs_addr);
}

__device__ __noinline__ uint32_t stats2_byte_exact(const StatsParams &P, uint32_t c, uint32_t q, int wrel, int k, uint32_t hs_addr)
{
    return stats2_byte(P, c, q, wrel, k, hs_addr);
}
// one of the last 1..3 bases of a read (static B scheme): a plain base with 0 <= q' < 64 is one increment, anything else
// takes the exact path.  (q - lo) mod 256 < 64 implies lo <= q <= hi because this kernel runs with lo <= 64 only.
__device__ __forceinline__ uint32_t stats2_tail_byte(const StatsParams &P, const Stats2K &K, uint32_t c, uint32_t q, int wrel, int k, uint32_t hs_addr)
{
    const uint32_t code = c & 7u;
    const uint32_t legal = prmt_raw(K.vlut_lo, V2LUT_HI, code) & 0xFFu;       // 'N' poisoned: goes to the exact path
    const uint32_t n6 = prmt_raw(K.n6_lo, N6_HI, code) & 0xFFu;
    const uint32_t qp = (q + K.neg_lo4) & 0xFFu;
    if (legal == c && qp < (uint32_t)ST_QWIN) {
        reds_inc(hs_addr + (n6 + qp) * S2_PITCH + (uint32_t)k * (4u * ST_MAXW) + 4u * (uint32_t)wrel);
        return 0u;
    }
    return stats2_byte_exact(P, c, q, wrel, k, hs_addr);
}

template <int WARPS, int BS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_stats2(const __grid_constant__ StatsParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NTHREADS = WARPS * 32;
    __shared__ __align__(8) uint64_t full_bar[WARPS];

    // one stage per warp: a warp's wait for HBM is covered by the other warps of the SM
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int S = P.stride, R = P.tile_reads;                          // R <= 8
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    uint8_t *wbase = smem + S2_HIST_BYTES + S2_DUMMY_BYTES + (size_t)w * (2u * slab_bytes);
    uint64_t *bar = &full_bar[w];
    const uint32_t ntiles = (uint32_t)((P.n + R - 1) / R);            // host guarantees n / R < 2^31
    const uint32_t gw = blockIdx.x * WARPS + w, GW = gridDim.x * WARPS;

    for (uint32_t i = tid * 16; i < (uint32_t)S2_HIST_BYTES; i += NTHREADS * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // lane 0 keeps the global addresses of its warp's next tile and advances them by one grid stride per tile
    const int64_t gstep = (int64_t)GW * R * S;
    const uint8_t *gs = P.seq + (int64_t)gw * R * S, *gq = P.qual + (int64_t)gw * R * S;
    auto issue = [&](uint32_t tile) {
        const uint32_t bytes = (tile + 1u == ntiles) ? (uint32_t)(P.n - (int64_t)tile * R) * (uint32_t)S : slab_bytes;
        mbar_arrive_expect_tx(bar, bytes * 2u);
        bulk_g2s(wbase, gs, bytes, bar);
        bulk_g2s(wbase + slab_bytes, gq, bytes, bar);
        gs += gstep; gq += gstep;
    };
    if (lane == 0 && gw < ntiles) issue(gw);

    const uint32_t hs_addr = smem_u32(smem);
    const int j = lane & 3, rr = lane >> 2;
    Stats2K K;
    const uint32_t zero = (uint32_t)((unsigned long long)P.n >> 62);       // 0, but only known at run time
    K.vlut_lo = VLUT_LO + zero; K.n6_lo = N6_LO + zero; K.neg_lo4 = zero - P.qk.lo4;
    const int passoff = 4 * P.w0, ncols = 4 * P.nw;
    const int nsb = P.nw > 16 ? 1 : 0;                           // A scheme over words 0..31 (nw <= ST_MAXW = 40: one superblock;
                                                                 // short windows use the 8-word blocks of the B scheme only)
    const int nb8 = (P.nw + 7) >> 3;                             // 8-word blocks in all
    uint32_t ksel[4], koff[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t k = (uint32_t)((j + i) & 3);
        ksel[i] = 0x4440u + k;
        koff[i] = k * (4u * ST_MAXW);
    }
    const uint32_t uoj0 = (uint32_t)(16 * rr + 4 * j);           // A scheme, t = 0: word 4*rr + j
    const int grr = ((rr & 3) << 1) | (rr >> 2);                 // B scheme skew: a bijection of 0..7 with g(rr+4) - g(rr) odd,
                                                                 // so rows rr and rr+4 (same bank octet at S = 160) never collide
    const uint32_t srow = smem_u32(wbase) + (uint32_t)(rr < R ? rr : 0) * (uint32_t)S + (uint32_t)passoff;
    const uint32_t qrow = srow + slab_bytes;
    const bool lane_on = rr < R;
    const int ulen = P.uniform_len;
    const bool ragged = P.len != nullptr;
    uint32_t parity = 0;

    for (uint32_t tile = gw; tile < ntiles; tile += GW) {
        const int64_t g = (int64_t)tile * R + rr;
        const bool active = lane_on && g < P.n;
        int L = 0;
        if (active) L = ragged ? __ldg(P.len + g) : ulen;
        const bool lenbad = active && (L <= 0 || L > S);
        if (lenbad) L = 0;
        int Lp = L - passoff;                                    // bases of this read inside the pass window
        if (Lp > ncols) Lp = ncols;
        const int lim = Lp - 4;                                  // a word at byte offset o is full iff o <= lim
        uint32_t bad = 0;
        mbar_wait(bar, parity);
        parity ^= 1u;

        if (nsb) {
            constexpr int sb = 0;
            if (__all_sync(0xFFFFFFFFu, lim >= 128 * sb + 124)) {   // every read of the tile fills the superblock
#pragma unroll
                for (int t = 0; t < 8; t += 2) {
                    const uint32_t o0 = 128u * (uint32_t)sb + ((uoj0 + 16u * t) & 0x7Fu);
                    const uint32_t o1 = 128u * (uint32_t)sb + ((uoj0 + 16u * t + 16u) & 0x7Fu);
                    stats2_pair<false>(P, K, lds32(srow + o0), lds32(qrow + o0), o0, lds32(srow + o1), lds32(qrow + o1), o1, hs_addr, ksel,
                                       koff, bad);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t o = 128u * (uint32_t)sb + ((uoj0 + 16u * t) & 0x7Fu);
                    const int vb = Lp - (int)o;
                    if (vb >= 4) stats2_word<false>(P, K, lds32(srow + o), lds32(qrow + o), o, hs_addr, ksel, koff, bad);
                    else if (vb > 0) stats2_word_masked(P, K, lds32(srow + o), lds32(qrow + o), o, vb, hs_addr, bad);   // last 1..3 bases
                }
            }
        }
        if (BS == 0) {
            for (int b8 = nsb * 4; b8 < nb8; b8++) {             // the words past the superblock, the read's last 1..3 bases included
                const uint32_t o0 = 32u * (uint32_t)b8 + 4u * (uint32_t)((2 * j + grr) & 7);
                const uint32_t o1 = 32u * (uint32_t)b8 + 4u * (uint32_t)((2 * j + 1 + grr) & 7);
                const int vb0 = Lp - (int)o0, vb1 = Lp - (int)o1;
                uint32_t sw0 = 0, qw0 = 0, sw1 = 0, qw1 = 0;
                if (vb0 > 0) { sw0 = lds32(srow + o0); qw0 = lds32(qrow + o0); }
                if (vb1 > 0) { sw1 = lds32(srow + o1); qw1 = lds32(qrow + o1); }
                stats2_pair_masked(P, K, sw0, qw0, o0, vb0, sw1, qw1, o1, vb1, hs_addr, ksel, koff, bad);
            }
        } else {
            // static B scheme (FXG_STATS_B=1, measured and NOT the default): the 8-word blocks past the superblock with the
            // A scheme's code (k static, full words only).  Fewer instructions (125 instead of 165 per tile at 150 bp),
            // but the four reads of a warp that meet in one column cost three extra wavefronts per ATOMS: 13.08 vs 13.02
            // Greads/s at 150 bp, 14.9 vs 17.8 at 50 bp, 3.05 vs 3.80 at 250 bp (B200, round 1)
            for (int b8 = nsb * 4; b8 < nb8; b8++) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const uint32_t o = 32u * (uint32_t)b8 + 16u * (uint32_t)((t + rr) & 1) + 4u * (uint32_t)j;
                    if ((int)o <= lim) stats2_word<false>(P, K, lds32(srow + o), lds32(qrow + o), o, hs_addr, ksel, koff, bad);
                }
            }
            // the last 1..3 bases, unless the ragged A path already took them: lane j of the read takes byte j
            if (Lp > 0 && j < (Lp & 3) && (Lp >> 2) >= 32 * nsb) {
                const uint32_t ob = (uint32_t)(Lp & ~3) + (uint32_t)j;
                bad |= stats2_tail_byte(P, K, lds8(srow + ob), lds8(qrow + ob), Lp >> 2, j, hs_addr);
            }
        }
        if ((bad != 0 || lenbad) && active)
            atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));

        __syncwarp();
        if (lane == 0 && tile + GW < ntiles) issue(tile + GW);
    }

    // flush the CTA's shared histogram into the global u64 table
    __syncthreads();
    for (int i = tid; i < S2_HIST_BYTES / 4; i += NTHREADS) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(smem + 4 * (size_t)i);
        if (v) {
            const int bin = i / (4 * ST_MAXW), pc = i - bin * (4 * ST_MAXW);
            const int k = pc / ST_MAXW, wr = pc - k * ST_MAXW;
            hist_global_add(P.hist, P.max_cycles, 4 * (P.w0 + wr) + k, bin >> 6, bin & 63, (unsigned long long)v);
        }
    }
}

// General fallback (any stride, FASTA input, per-read weights): one thread per 16-byte chunk, global atomics.
__global__ void __launch_bounds__(256) k_stats_simple(const StatsParams P)
{
    const int chunks = P.stride >> 4;
    const int64_t total = P.n * chunks;
    const QualK qk = P.qk;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        int L = P.len ? __ldg(P.len + i) : P.uniform_len;
        const bool lenbad = (L <= 0 || L > P.stride);
        if (lenbad) { if (c == 0) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + i)); continue; }
        const int nb = L - 16 * c;
        if (nb <= 0) continue;
        const unsigned long long wgt = P.weight ? (unsigned long long)__ldg(P.weight + i) : 1ull;
        const size_t off = (size_t)i * P.stride + (size_t)c * 16;
        const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(P.seq + off));
        uint4 q4 = make_uint4(0, 0, 0, 0);
        if (P.qual) q4 = __ldg(reinterpret_cast<const uint4 *>(P.qual + off));
        const uint32_t sw[4] = { s4.x, s4.y, s4.z, s4.w }, qw[4] = { q4.x, q4.y, q4.z, q4.w };
        bool bad = false;
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            const uint32_t m = head_mask(nb - 4 * wd);
            if (seq_bad_bits(sw[wd]) & m) bad = true;
            if (P.qual && (qual_bad_bits(qw[wd], qw[wd] | HI, qk) & HI & m)) bad = true;
        }
        if (bad) { atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + i)); continue; }
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            const uint32_t nuc4 = __byte_perm(NLUT_LO, NLUT_HI, base_selector(sw[wd]));
            const uint32_t qp4 = P.qual ? (qw[wd] - qk.lo4) : 0x0F0F0F0Fu;   // FASTA: all counts land in the q = 0 bin
            for (int k = 0; k < 4; k++) {
                if (4 * wd + k < nb)
                    hist_global_add(P.hist, P.max_cycles, 16 * c + 4 * wd + k, (nuc4 >> (8 * k)) & 0xFF, (qp4 >> (8 * k)) & 0xFF, wgt);
            }
        }
    }
}

cudaError_t launch_stats(const StatsParams &p, int g, int grid, uint32_t smem_bytes, cudaStream_t st)
{
#define FXG_STATS_LAUNCH(GV)                                                                        \
    do {                                                                                           \
        cudaFuncSetAttribute(k_stats<GV>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM); \
        k_stats<GV><<<grid, ST_WARPS * 32 * GV, smem_bytes, st>>>(p);                               \
    } while (0)
    if (g == 1) FXG_STATS_LAUNCH(1);
    else if (g == 2) FXG_STATS_LAUNCH(2);
    else if (g == 4) FXG_STATS_LAUNCH(4);
    else return cudaErrorInvalidValue;
#undef FXG_STATS_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_stats2(const StatsParams &p, int warps, int grid, uint32_t smem_bytes, cudaStream_t st)
{
#define FXG_STATS2_LAUNCH(WV, BV)                                                                        \
    do {                                                                                                \
        cudaFuncSetAttribute(k_stats2<WV, BV>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM); \
        k_stats2<WV, BV><<<grid, WV * 32, smem_bytes, st>>>(p);                                          \
    } while (0)
    // p.stages doubles as the B-scheme selector: 0 = masked blocks with per-lane byte order, 1 = static blocks
    if (warps == 24 && p.stages == 0) FXG_STATS2_LAUNCH(24, 0);
    else if (warps == 24) FXG_STATS2_LAUNCH(24, 1);
    else if (warps == 20) FXG_STATS2_LAUNCH(20, 0);
    else if (warps == 16) FXG_STATS2_LAUNCH(16, 0);
    else if (warps == 12) FXG_STATS2_LAUNCH(12, 0);
    else return cudaErrorInvalidValue;
#undef FXG_STATS2_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_stats_simple(const StatsParams &p, int sm_count, cudaStream_t st)
{
    const int64_t total = p.n * (p.stride >> 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    if (blocks < 1) blocks = 1;
    k_stats_simple<<<(unsigned)blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t stats_set_smem_attrs() { return cudaSuccess; }

}  // namespace fxg
