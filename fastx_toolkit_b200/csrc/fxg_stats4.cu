// fxg_stats4.cu — K-STATS (`k_stats4`): lane = read, 32-read tiles, two warps per tile.
// Accumulates hist[cycle][nuc][q+15] of src/fastx_quality_stats/fastx_quality_stats.c:166-216 (`read_file`).
//
// Round 1's kernel (4 lanes per read, 8-read tiles; profiles/r01_ncu_stats_full.txt) was bound by instruction issue — per
// tile it paid ~160 warp-instructions of prologue and TMA issue for only 1 200 bases, and the six words past its 32-word
// superblock cost 3.8x more per word than the rest.  Here a lane owns a whole read, so
//   * the per-tile cost (barrier wait, lengths, TMA issue by lane 0) is spread over 32 reads instead of 8,
//   * every lane is busy in every step of the A region (words 0..31 = eight 16-byte chunks, read with LDS.128) and of the
//     B region (words 32..39), with no cross-lane reduction anywhere,
//   * per base: one PRMT (bin byte), one IMAD (bin * pitch + column), one RED.
//
// Conflict freedom by construction (checked bank by bank in tests/test_stats4_model.py): a bin (nuc*64 + q') owns 96
// consecutive 32-bit words of shared memory, so the bank of a counter never depends on the data.
//   A region: at chunk step t lane l works on chunk c = (t + r_l) & 7 and, in byte step i, on byte k = (i + kb_l) & 3 of the
//             word wi (static in the instruction).  (r_l, kb_l) is a bijection of the 32 lanes onto 8 x 4, so the 32
//             counters of one RED have 32 different (c, k): word 32*(wi>>1) + 8k + c of the bin = 32 different banks.
//             Words wi and wi^1 share a 32-bit word as two u16 halves (the increment is the immediate 1 or 65 536);
//             a half is flushed into the global u64 table before it can wrap (every S4_FLUSH_ROUNDS tiles per warp).
//   B region: lane l visits word W = ((l>>2) + a) & 7 in word step a and byte k = (b + l) & 3 in byte step b: 32 different
//             (W, k) per RED, full u32 counters at word 64 + 4W + k of the bin = 32 different banks.
//   LDS.128 of the A region: for an even row pitch (in 16-byte units) r_l = l & 7 puts the 8 lanes of a quarter-warp on 8
//             different bank groups; for an odd pitch r_l = (2(l&7) + ((l>>3)&1)) & 7 does.
// Bytes the packed test cannot take (an 'N', q' >= 64, an illegal character) go one by one through the exact path;
// 'N' and q' >= 64 count straight into the global table.
#include "fxg_kernels.cuh"

namespace fxg {

// shared with fxg_stats.cu (same tables)
constexpr uint32_t S4_NLUT_LO = 0x01800080u, S4_NLUT_HI = 0x02048003u;   // nucleotide index by base code (0x80 = not a base)
constexpr uint32_t S4_V2LUT_HI = 0x47FFFF54u;                            // VLUT_HI with 'N' poisoned
constexpr uint32_t S4_N6_LO = 0x40000000u, S4_N6_HI = 0x800000C0u;       // nuc << 6 by base code

__device__ __forceinline__ void s4_global_add(unsigned long long *hist, int max_cycles, int cycle, int nuc, int qp, unsigned long long w)
{
    if (cycle < max_cycles) atomicAdd(&hist[((size_t)cycle * 5 + nuc) * 109 + qp], w);
}
__device__ __forceinline__ uint32_t s4_lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 s4_lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void s4_red(uint32_t addr, uint32_t inc)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
}

// increment only when `on` (a predicated RED: lanes that are off touch no bank at all)
__device__ __forceinline__ void s4_red_if(uint32_t addr, uint32_t inc, bool on)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p red.shared.add.u32 [%0], %1;\n}" ::"r"(addr), "r"(inc), "r"((uint32_t)on) : "memory");
}
// increment only when k < vb (the byte lies inside the read)
__device__ __forceinline__ void s4_red_lt(uint32_t addr, uint32_t inc, int k, int vb)
{
    asm volatile("{\n .reg .pred p;\n setp.lt.s32 p, %2, %3;\n @p red.shared.add.u32 [%0], %1;\n}" ::"r"(addr), "r"(inc), "r"(k), "r"(vb) : "memory");
}
__device__ __forceinline__ void s4_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte address (relative to the histogram) and increment of the shared counter of (bin, window word w, byte k)
__device__ __forceinline__ uint32_t s4_counter(uint32_t bin, int w, int k, uint32_t &inc)
{
    if (w < 32) {
        const int c = w >> 2, wi = w & 3;
        inc = (wi & 1) ? 0x10000u : 1u;
        return bin * (uint32_t)S4_PITCH + 4u * (uint32_t)(32 * (wi >> 1) + 8 * k + c);
    }
    inc = 1u;
    return bin * (uint32_t)S4_PITCH + 4u * (uint32_t)(64 + 4 * (w - 32) + k);
}

// loop-invariant operands kept in registers (opaque to the compiler, which would otherwise rebuild the immediates)
struct Stats4K {
    uint32_t vlut_lo, n6_lo, neg_lo4;
};

// comb = (nuc<<6 | q') per byte; the returned word is 0 iff the four bytes are plain A/C/G/T with 0 <= q' < 64, which
// (this kernel runs with Q - 15 <= 64 only) implies the reader's quality range check
__device__ __forceinline__ uint32_t s4_decode(const Stats4K &K, uint32_t sw, uint32_t qw, uint32_t &comb)
{
    const uint32_t y = sw & 0x07070707u;
    const uint32_t sel = prmt_raw(y | (y >> 4), 0u, 0x4420u);
    const uint32_t e = prmt_raw(K.vlut_lo, S4_V2LUT_HI, sel);
    const uint32_t n6 = prmt_raw(K.n6_lo, S4_N6_HI, sel);
    comb = n6 + qw + K.neg_lo4;
    return (sw ^ e) | ((comb ^ n6) & 0xC0C0C0C0u);
}

// one byte, exactly as the reader and read_file treat it.  Returns 1 when the base or its quality is illegal.
__device__ __forceinline__ uint32_t s4_byte(const StatsParams &P, uint32_t c, uint32_t q, int w, int k, uint32_t hs_addr)
{
    const uint32_t lo = P.qk.lo4 & 0xFFu, hmax = 127u - (P.qk.hik4 & 0xFFu) - lo;
    const uint32_t code = c & 7u;
    const uint32_t legal = __byte_perm(VLUT_LO, VLUT_HI, code) & 0xFFu;
    const uint32_t nuc = __byte_perm(S4_NLUT_LO, S4_NLUT_HI, code) & 0xFFu;
    const uint32_t qp = q - lo;
    if (legal != c || qp > hmax) return 1u;
    if (nuc < 4u && qp < 64u) {
        uint32_t inc;
        const uint32_t a = s4_counter(nuc * 64u + qp, w, k, inc);
        s4_red(hs_addr + a, inc);
    } else {
        s4_global_add(P.hist, P.max_cycles, 4 * (P.w0 + w) + k, (int)nuc, (int)qp, 1ull);
    }
    return 0u;
}
__device__ __noinline__ uint32_t s4_slow_word(const StatsParams &P, uint32_t sw, uint32_t qw, int w, int nbytes, uint32_t hs_addr)
{
    uint32_t bad = 0;
    for (int k = 0; k < nbytes; k++) bad |= s4_byte(P, (sw >> (8 * k)) & 0xFFu, (qw >> (8 * k)) & 0xFFu, w, k, hs_addr);
    return bad;
}

// A region, full word: four REDs.  cb[i] = hs_addr + 32*k_i + 4c (per lane, per chunk), IMM = byte offset of the word pair,
// INC = 1 or 65 536
template <int WI>
__device__ __forceinline__ void s4_emit_a(uint32_t comb, const uint32_t (&ksel)[4], const uint32_t (&cb)[4])
{
    constexpr uint32_t IMM = 128u * (uint32_t)(WI >> 1), INC = (WI & 1) ? 0x10000u : 1u;
#pragma unroll
    for (int i = 0; i < 4; i++) s4_red(prmt_raw(comb, 0u, ksel[i]) * (uint32_t)S4_PITCH + cb[i] + IMM, INC);
}
// the same with bytes past the end of the read steered to a scratch counter (vb = valid bytes of the word, may be <= 0)
template <int WI>
__device__ __forceinline__ void s4_emit_a_masked(uint32_t comb, int vb, const uint32_t (&ksel)[4], const uint32_t (&cb)[4])
{
    constexpr uint32_t IMM = 128u * (uint32_t)(WI >> 1), INC = (WI & 1) ? 0x10000u : 1u;
#pragma unroll
    for (int i = 0; i < 4; i++)
        s4_red_lt(prmt_raw(comb, 0u, ksel[i]) * (uint32_t)S4_PITCH + cb[i] + IMM, INC, (int)(ksel[i] & 3u), vb);
}

// TILES tile buffers per CTA, PAIR warps per buffer (PAIR = 2: the two warps of a pair read the same 32 reads and split the
// chunk steps between them — twice the warps to hide latency behind, for the same shared memory)
template <int TILES, int PAIR>
__global__ void __launch_bounds__(TILES * PAIR * 32, 1) k_stats4(const __grid_constant__ StatsParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NTHREADS = TILES * PAIR * 32;
    __shared__ __align__(8) uint64_t full_bar[TILES], empty_bar[TILES];

    const int tid = threadIdx.x, lane = tid & 31, w = (tid >> 5) / PAIR, half = (tid >> 5) % PAIR;
    const int S = P.stride, R = P.tile_reads;                          // R <= 32 reads per tile, one per lane
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    uint8_t *wbase = smem + S4_HIST_BYTES + S4_DUMMY_BYTES + (size_t)w * (2u * slab_bytes);
    uint64_t *bar = &full_bar[w], *ebar = &empty_bar[w];
    const uint32_t ntiles = (uint32_t)((P.n + R - 1) / R);            // host guarantees n / R < 2^31
    const uint32_t gw = blockIdx.x * TILES + w, GW = gridDim.x * TILES;
    const uint32_t gw0 = blockIdx.x * TILES;                           // the CTA's first tile slot has the most tiles
    const uint32_t rounds = gw0 < ntiles ? (ntiles - gw0 + GW - 1) / GW : 0u;
    constexpr int T0 = 8 / PAIR;                                       // chunk / word steps per warp

    for (uint32_t i = tid * 16; i < (uint32_t)(S4_HIST_BYTES + S4_DUMMY_BYTES); i += NTHREADS * 16)
        *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    if (lane == 0 && half == 0) {
        mbar_init(bar, 1);
        mbar_init(ebar, PAIR);
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t gstep = (int64_t)GW * R * S;
    const uint8_t *gs = P.seq + (int64_t)gw * R * S, *gq = P.qual + (int64_t)gw * R * S;
    auto issue = [&](uint32_t tile) {
        const uint32_t bytes = (tile + 1u == ntiles) ? (uint32_t)(P.n - (int64_t)tile * R) * (uint32_t)S : slab_bytes;
        mbar_arrive_expect_tx(bar, bytes * 2u);
        bulk_g2s(wbase, gs, bytes, bar);
        bulk_g2s(wbase + slab_bytes, gq, bytes, bar);
        gs += gstep; gq += gstep;
    };
    if (lane == 0 && half == 0 && gw < ntiles) issue(gw);

    const uint32_t hs_addr = smem_u32(smem);
    Stats4K K;
    // 0, but read from (zeroed) shared memory: per thread and opaque, so the three constants live in ordinary registers — as
    // uniform values or immediates they are copied into one in front of every PRMT that uses them
    const uint32_t zero = s4_lds32(hs_addr + (uint32_t)S4_HIST_BYTES + 4u * (uint32_t)lane);
    K.vlut_lo = VLUT_LO + zero; K.n6_lo = S4_N6_LO + zero; K.neg_lo4 = zero - P.qk.lo4;
    const int passoff = 4 * P.w0, ncols = 4 * P.nw;              // this pass covers cycles [passoff, passoff + ncols), ncols <= 160

    // per-lane schedule constants
    const int q8 = lane >> 3, i8 = lane & 7;
    const bool odd_pitch = ((S >> 4) & 1) != 0;
    const int r_l = odd_pitch ? ((2 * i8 + (q8 & 1)) & 7) : i8;
    const int kb_l = odd_pitch ? (2 * (q8 >> 1) + (i8 >> 2)) : q8;
    uint32_t ksel[4], kcol[4], kselb[4], kcolb[4];
    int kb4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t k = (uint32_t)((i + kb_l) & 3);
        ksel[i] = 0x4440u + k;
        kcol[i] = hs_addr + 32u * k;                             // + 4c per chunk
        const uint32_t kb = (uint32_t)((i + lane) & 3);
        kselb[i] = 0x4440u + kb;
        kcolb[i] = hs_addr + 4u * (64u + kb);                    // + 16W per word
        kb4[i] = (int)kb;
    }
    const int wb0 = lane >> 2;
    const uint32_t srow = smem_u32(wbase) + (uint32_t)(lane < R ? lane : 0) * (uint32_t)S + (uint32_t)passoff;
    const uint32_t qrow = srow + slab_bytes;
    const int ulen = P.uniform_len;
    const bool ragged = P.len != nullptr;
    const uint32_t lo4 = P.qk.lo4;
    uint32_t parity = 0, eparity = 0;
    const int tbeg = half * T0;

    for (uint32_t round = 0; round < rounds; round++) {
        const uint32_t tile = gw + round * GW;
        if (tile < ntiles) {
            const int64_t g = (int64_t)tile * R + lane;
            const bool active = lane < R && g < P.n;
            int L = 0;
            if (active) L = ragged ? __ldg(P.len + g) : ulen;
            const bool lenbad = active && (L <= 0 || L > S);
            if (lenbad) L = 0;
            int Lp = L - passoff;                                    // bases of this read inside the pass window
            if (Lp > ncols) Lp = ncols;
            uint32_t bad = 0;
            mbar_wait(bar, parity);
            parity ^= 1u;

            // ---- A region: words 0..31 ----
            if (__all_sync(0xFFFFFFFFu, Lp >= 128)) {
#pragma unroll 2
                for (int t = tbeg; t < tbeg + T0; t++) {
                    const uint32_t c = (uint32_t)((t + r_l) & 7);
                    const uint4 s4 = s4_lds128(srow + 16u * c), q4 = s4_lds128(qrow + 16u * c);
                    uint32_t cb[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) cb[i] = kcol[i] + 4u * c;
                    uint32_t c0, c1, c2, c3;
                    const uint32_t t0 = s4_decode(K, s4.x, q4.x, c0), t1 = s4_decode(K, s4.y, q4.y, c1);
                    const uint32_t t2 = s4_decode(K, s4.z, q4.z, c2), t3 = s4_decode(K, s4.w, q4.w, c3);
                    if ((t0 | t1 | t2 | t3) == 0u) {
                        s4_emit_a<0>(c0, ksel, cb); s4_emit_a<1>(c1, ksel, cb); s4_emit_a<2>(c2, ksel, cb); s4_emit_a<3>(c3, ksel, cb);
                    } else {
                        const int w4 = 4 * (int)c;
                        if (t0 == 0u) s4_emit_a<0>(c0, ksel, cb); else bad |= s4_slow_word(P, s4.x, q4.x, w4 + 0, 4, hs_addr);
                        if (t1 == 0u) s4_emit_a<1>(c1, ksel, cb); else bad |= s4_slow_word(P, s4.y, q4.y, w4 + 1, 4, hs_addr);
                        if (t2 == 0u) s4_emit_a<2>(c2, ksel, cb); else bad |= s4_slow_word(P, s4.z, q4.z, w4 + 2, 4, hs_addr);
                        if (t3 == 0u) s4_emit_a<3>(c3, ksel, cb); else bad |= s4_slow_word(P, s4.w, q4.w, w4 + 3, 4, hs_addr);
                    }
                }
            } else {
                // some read of the tile ends inside the A region: every word carries its count of valid bytes
                // short reads: the rotation runs over the smallest power of two of chunks that covers the longest read of the
                // tile (several lanes then share a counter: a same-address RED costs wavefronts, not instructions)
                const int tmax = (__reduce_max_sync(0xFFFFFFFFu, Lp) + 15) >> 4;     // chunks any lane still needs (warp uniform)
                if (tmax > 0) {
                    const int nc = tmax > 4 ? 8 : tmax > 2 ? 4 : tmax > 1 ? 2 : 1;
                    const int per = nc >= PAIR ? nc / PAIR : 1;                      // chunk steps per warp of the pair
                    const int t_lo = nc >= PAIR ? half * per : 0, t_hi = nc >= PAIR ? t_lo + per : (half == 0 ? 1 : 0);
                    for (int t = t_lo; t < t_hi; t++) {
                        const uint32_t c = (uint32_t)((t + r_l) & (nc - 1));
                        const int vbc = Lp - 16 * (int)c;                            // valid bytes from this chunk on
                        if (!__any_sync(0xFFFFFFFFu, vbc > 0)) continue;
                        uint4 s4 = make_uint4(0, 0, 0, 0), q4 = make_uint4(0, 0, 0, 0);
                        if (vbc > 0) { s4 = s4_lds128(srow + 16u * c); q4 = s4_lds128(qrow + 16u * c); }
                        uint32_t cb[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) cb[i] = kcol[i] + 4u * c;
                        const uint32_t sws[4] = { s4.x, s4.y, s4.z, s4.w }, qws[4] = { q4.x, q4.y, q4.z, q4.w };
                        uint32_t comb[4], tst[4];
#pragma unroll
                        for (int wi = 0; wi < 4; wi++) {
                            const uint32_t m = head_mask(vbc - 4 * wi);
                            tst[wi] = s4_decode(K, (sws[wi] & m) | (0x41414141u & ~m), (qws[wi] & m) | (lo4 & ~m), comb[wi]);
                        }
                        const int w4 = 4 * (int)c;
#define S4_MASKED_WORD(WI)                                                                                           \
    do {                                                                                                             \
        const int vb_ = vbc - 4 * WI;                                                                                \
        if (tst[WI] == 0u) s4_emit_a_masked<WI>(comb[WI], vb_, ksel, cb);                                     \
        else if (vb_ > 0) bad |= s4_slow_word(P, sws[WI], qws[WI], w4 + WI, vb_ < 4 ? vb_ : 4, hs_addr);             \
    } while (0)
                        S4_MASKED_WORD(0); S4_MASKED_WORD(1); S4_MASKED_WORD(2); S4_MASKED_WORD(3);
#undef S4_MASKED_WORD
                    }
                }
            }
            // ---- B region: words 32..39 (the read's last bases included) ----
            {
                const int LpB = Lp - 128;
                const bool b_full4 = __all_sync(0xFFFFFFFFu, LpB >= 16);      // every read holds words 32..35 complete
                if (b_full4) {
                    // words 32..35 without masks: lane l visits word (l>>2 + a) & 3 — two lanes per counter (same address:
                    // two wavefronts per RED), a fraction of the masked steps' instructions
#pragma unroll
                    for (int a = half * (4 / PAIR); a < (half + 1) * (4 / PAIR); a++) {
                        const uint32_t W = (uint32_t)((wb0 + a) & 3);
                        const uint32_t sw = s4_lds32(srow + 128u + 4u * W), qw = s4_lds32(qrow + 128u + 4u * W);
                        uint32_t comb;
                        if (s4_decode(K, sw, qw, comb) == 0u) {
#pragma unroll
                            for (int b = 0; b < 4; b++) s4_red(prmt_raw(comb, 0u, kselb[b]) * (uint32_t)S4_PITCH + kcolb[b] + 16u * W, 1u);
                        } else {
                            bad |= s4_slow_word(P, sw, qw, 32 + (int)W, 4, hs_addr);
                        }
                    }
                }
                // the rest (words 36.. after the full block, else 32..): every byte carries its validity.  The rotation runs over
                // the smallest power of two of words that covers the tile's longest read (150 bp: words 36 and 37 only — lanes
                // then share counters, which costs wavefronts of the half-idle LSU, not issue slots)
                const int wlo = b_full4 ? 4 : 0;
                const int need = (__reduce_max_sync(0xFFFFFFFFu, LpB) - 4 * wlo + 3) >> 2;        // words any lane still needs
                if (need > 0) {
                    const int wn = need > 4 ? 8 : need > 2 ? 4 : need > 1 ? 2 : 1;
                    const int per = wn >= PAIR ? wn / PAIR : 1;
                    const int a_lo = wn >= PAIR ? half * per : 0, a_hi = wn >= PAIR ? a_lo + per : (half == 0 ? 1 : 0);
#pragma unroll 2
                    for (int a = a_lo; a < a_hi; a++) {
                        const uint32_t W = (uint32_t)(wlo + ((wb0 + a) & (wn - 1)));
                        const int vb = LpB - 4 * (int)W;
                        uint32_t sw = 0, qw = 0;
                        if (vb > 0) { sw = s4_lds32(srow + 128u + 4u * W); qw = s4_lds32(qrow + 128u + 4u * W); }
                        const uint32_t m = head_mask(vb);
                        uint32_t comb;
                        const uint32_t tst = s4_decode(K, (sw & m) | (0x41414141u & ~m), (qw & m) | (lo4 & ~m), comb);
                        if (tst == 0u) {
#pragma unroll
                            for (int b = 0; b < 4; b++)
                                s4_red_lt(prmt_raw(comb, 0u, kselb[b]) * (uint32_t)S4_PITCH + kcolb[b] + 16u * W, 1u, kb4[b], vb);
                        } else if (vb > 0) {
                            bad |= s4_slow_word(P, sw, qw, 32 + (int)W, vb < 4 ? vb : 4, hs_addr);
                        }
                    }
                }
            }
            if ((bad != 0 || lenbad) && active) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));

            __syncwarp();
            if (PAIR == 1) {
                if (lane == 0 && tile + GW < ntiles) issue(tile + GW);
            } else if (lane == 0) {
                s4_mbar_arrive(ebar);                                  // this warp is done with the buffer
                if (half == 0 && tile + GW < ntiles) {                 // the pair's first warp refills it when both are
                    mbar_wait(ebar, eparity);
                    issue(tile + GW);
                }
                eparity ^= 1u;
            }
        }
        // a u16 half of the A region holds at most one increment per read: flush before 65 535 reads went through this CTA
        if ((round + 1u) % (uint32_t)(65535 / (TILES * 32)) == 0u || round + 1u == rounds) {
            __syncthreads();
            for (int i = tid; i < S4_HIST_BYTES / 4; i += NTHREADS) {
                uint32_t *cell = reinterpret_cast<uint32_t *>(smem) + i;
                const uint32_t v = *cell;
                if (v == 0u) continue;
                *cell = 0u;
                const int bin = i / (S4_PITCH / 4), col = i - bin * (S4_PITCH / 4);
                if (col < 64) {
                    const int hi = col >> 5, k = (col & 31) >> 3, c = col & 7;
                    const int w0 = 4 * c + 2 * hi;                     // low half: word w0, high half: word w0 + 1
                    if (v & 0xFFFFu) s4_global_add(P.hist, P.max_cycles, 4 * (P.w0 + w0) + k, bin >> 6, bin & 63, (unsigned long long)(v & 0xFFFFu));
                    if (v >> 16) s4_global_add(P.hist, P.max_cycles, 4 * (P.w0 + w0 + 1) + k, bin >> 6, bin & 63, (unsigned long long)(v >> 16));
                } else {
                    const int j = col - 64;
                    s4_global_add(P.hist, P.max_cycles, 4 * (P.w0 + 32 + (j >> 2)) + (j & 3), bin >> 6, bin & 63, (unsigned long long)v);
                }
            }
            __syncthreads();
        }
    }
}

cudaError_t launch_stats4(const StatsParams &p, int grid, uint32_t smem_bytes, cudaStream_t st)
{
    // two warps per tile buffer: 24 warps hide the latencies that 12 do not (13.7 -> 13.9-14.5 G reads/s at 150 bp, round 2)
    cudaFuncSetAttribute(k_stats4<S4_WARPS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
    k_stats4<S4_WARPS, 2><<<grid, S4_WARPS * 64, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fxg
