// fxg_stats3.cu — K-STATS, third generation (EXPERIMENTAL, not the default: FXG_STATS_V=3; not yet measured on a GPU).
//
// Same reference loop as fxg_stats.cu (src/fastx_quality_stats/fastx_quality_stats.c:166-216) and the same
// conflict-free idea as k_stats2 (bank of a counter = function of the cycle only), with the two changes the round-1
// profiles ask for:
//   * u16 counters, paired along k: the counters of bytes k and k+2 of a word share one 32-bit word, so an increment
//     is "add 1" (k = 0,1) or "add 65536" (k = 2,3) — an immediate in the A scheme.  Word (bin, k&1, w) sits at
//     bin*384 + (k&1)*160 + 4w: the pitch of a bin is padded from 80 to 96 words so that it stays a multiple of 32
//     banks and the bank, (8*(k&1) + w) mod 32, stays independent of the data (the Python model caught the unpadded
//     pitch).  The histogram shrinks from 160 KB to 96 KB; a CTA flushes it to the global u64 table every 341
//     iterations (24 warps x 8 reads x 341 <= 65 535, so no half can overflow).
//   * the 64 KB freed hold a SECOND tile buffer per warp: the bulk copy of tile i+1 runs under the work on tile i
//     (k_stats2: one buffer, a warp idles for the latency of its own load).
// Lane schedules (A scheme on words 0..31, masked B scheme on the 8-word blocks behind them), the packed decode and the
// exact per-byte path are those of k_stats2; tests/stats2_model.py models this layout too (layout16=True).
#include "fxg_kernels.cuh"

namespace fxg {
namespace s3 {

constexpr int WARPS = 24;
constexpr int NTHREADS = WARPS * 32;
constexpr uint32_t PITCH = 384;                 // bytes per bin: 2 x 40 words, padded to 96 (a multiple of 32 banks)
constexpr int HIST_BYTES = 256 * 384;           // 98 304
constexpr int DUMMY_BYTES = 128;
constexpr int FLUSH_EVERY = 65535 / (WARPS * 8);   // iterations between flushes (341)

constexpr uint32_t NLUT_LO = 0x01800080u, NLUT_HI = 0x02048003u;     // as in fxg_stats.cu
constexpr uint32_t V2LUT_HI = 0x47FFFF54u, N6_LO = 0x40000000u, N6_HI = 0x800000C0u;

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void reds_add1(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void reds_add64k(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 65536;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void gadd(unsigned long long *hist, int max_cycles, int cycle, int nuc, int qp, unsigned long long w)
{
    if (cycle < max_cycles) atomicAdd(&hist[((size_t)cycle * 5 + nuc) * 109 + qp], w);
}
// counter of (bin, byte k of word wrel): address and increment
__device__ __forceinline__ uint32_t caddr(uint32_t hs_addr, uint32_t bin, uint32_t k, uint32_t wrel) { return hs_addr + bin * PITCH + (k & 1u) * 160u + 4u * wrel; }
__device__ __forceinline__ uint32_t cinc(uint32_t k) { return (k & 2u) ? 65536u : 1u; }

__device__ __forceinline__ uint32_t byte_exact(const StatsParams &P, uint32_t c, uint32_t q, int wrel, int k, uint32_t hs_addr)
{
    const uint32_t lo = P.qk.lo4 & 0xFFu, hmax = 127u - (P.qk.hik4 & 0xFFu) - lo;
    const uint32_t code = c & 7u;
    const uint32_t legal = __byte_perm(VLUT_LO, VLUT_HI, code) & 0xFFu;
    const uint32_t nuc = __byte_perm(NLUT_LO, NLUT_HI, code) & 0xFFu;
    const uint32_t qp = q - lo;
    if (legal != c || qp > hmax) return 1u;
    if (nuc < 4u && qp < 64u) reds_add(caddr(hs_addr, nuc * 64u + qp, (uint32_t)k, (uint32_t)wrel), cinc((uint32_t)k));
    else gadd(P.hist, P.max_cycles, 4 * (P.w0 + wrel) + k, (int)nuc, (int)qp, 1ull);
    return 0u;
}
__device__ __noinline__ uint32_t slow_word(const StatsParams &P, uint32_t sw, uint32_t qw, int wrel, int nbytes, uint32_t hs_addr)
{
    uint32_t bad = 0;
    for (int k = 0; k < nbytes; k++) bad |= byte_exact(P, (sw >> (8 * k)) & 0xFFu, (qw >> (8 * k)) & 0xFFu, wrel, k, hs_addr);
    return bad;
}

struct K3 { uint32_t vlut_lo, n6_lo, neg_lo4; };

__device__ __forceinline__ uint32_t decode(const K3 &K, uint32_t sw, uint32_t qw, uint32_t &comb)
{
    const uint32_t y = sw & 0x07070707u;
    const uint32_t sel = prmt_raw(y | (y >> 4), 0u, 0x4420u);
    const uint32_t e = prmt_raw(K.vlut_lo, V2LUT_HI, sel);
    const uint32_t n6 = prmt_raw(K.n6_lo, N6_HI, sel);
    comb = n6 + qw + K.neg_lo4;
    return (sw ^ e) | ((comb ^ n6) & 0xC0C0C0C0u);
}
// A scheme: byte order k = 0..3, increments are immediates
__device__ __forceinline__ void emit_static(uint32_t comb, uint32_t o, uint32_t hs_addr)
{
    const uint32_t b0 = comb & 0xFFu, b1 = prmt_raw(comb, 0u, 0x4441u), b2 = prmt_raw(comb, 0u, 0x4442u), b3 = comb >> 24;
    const uint32_t col = hs_addr + o;
    reds_add1(b0 * PITCH + col);
    reds_add1(b1 * PITCH + col + 160u);
    reds_add64k(b2 * PITCH + col);
    reds_add64k(b3 * PITCH + col + 160u);
}
// masked emit: byte k counts iff k < vb; kk[i] = byte index of the i-th increment (static {0,1,2,3} or rotated per lane)
__device__ __forceinline__ void emit_masked(uint32_t comb, uint32_t o, int vb, uint32_t hs_addr, const uint32_t (&kk)[4])
{
    const uint32_t col = hs_addr + o;
    const uint32_t dummy = hs_addr + (uint32_t)HIST_BYTES + ((o >> 2) & 31u) * 4u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t k = kk[i];
        const uint32_t addr = prmt_raw(comb, 0u, 0x4440u + k) * PITCH + col + (k & 1u) * 160u;
        reds_add((int)k < vb ? addr : dummy, cinc(k));
    }
}
__device__ __forceinline__ void pair_full(const StatsParams &P, const K3 &K, uint32_t sw0, uint32_t qw0, uint32_t o0, uint32_t sw1, uint32_t qw1,
                                          uint32_t o1, uint32_t hs_addr, uint32_t &bad)
{
    uint32_t c0, c1;
    const uint32_t t0 = decode(K, sw0, qw0, c0), t1 = decode(K, sw1, qw1, c1);
    if ((t0 | t1) == 0u) {
        emit_static(c0, o0, hs_addr);
        emit_static(c1, o1, hs_addr);
    } else {
        if (t0 == 0u) emit_static(c0, o0, hs_addr); else bad |= slow_word(P, sw0, qw0, (int)(o0 >> 2), 4, hs_addr);
        if (t1 == 0u) emit_static(c1, o1, hs_addr); else bad |= slow_word(P, sw1, qw1, (int)(o1 >> 2), 4, hs_addr);
    }
}
__device__ __forceinline__ void word_masked(const StatsParams &P, const K3 &K, uint32_t sw, uint32_t qw, uint32_t o, int vb, uint32_t hs_addr,
                                            const uint32_t (&kk)[4], uint32_t &bad)
{
    const uint32_t m = head_mask(vb), lo4 = 0u - K.neg_lo4;
    uint32_t c;
    if (decode(K, (sw & m) | (0x41414141u & ~m), (qw & m) | (lo4 & ~m), c) == 0u) emit_masked(c, o, vb, hs_addr, kk);
    else bad |= slow_word(P, sw, qw, (int)(o >> 2), vb < 4 ? vb : 4, hs_addr);
}
__device__ __forceinline__ void pair_masked(const StatsParams &P, const K3 &K, uint32_t sw0, uint32_t qw0, uint32_t o0, int vb0, uint32_t sw1,
                                            uint32_t qw1, uint32_t o1, int vb1, uint32_t hs_addr, const uint32_t (&kk)[4], uint32_t &bad)
{
    const uint32_t m0 = head_mask(vb0), m1 = head_mask(vb1), lo4 = 0u - K.neg_lo4;
    uint32_t c0, c1;
    const uint32_t t0 = decode(K, (sw0 & m0) | (0x41414141u & ~m0), (qw0 & m0) | (lo4 & ~m0), c0);
    const uint32_t t1 = decode(K, (sw1 & m1) | (0x41414141u & ~m1), (qw1 & m1) | (lo4 & ~m1), c1);
    if ((t0 | t1) == 0u) {
        emit_masked(c0, o0, vb0, hs_addr, kk);
        emit_masked(c1, o1, vb1, hs_addr, kk);
    } else {
        if (t0 == 0u) emit_masked(c0, o0, vb0, hs_addr, kk); else bad |= slow_word(P, sw0, qw0, (int)(o0 >> 2), vb0 < 4 ? vb0 : 4, hs_addr);
        if (t1 == 0u) emit_masked(c1, o1, vb1, hs_addr, kk); else bad |= slow_word(P, sw1, qw1, (int)(o1 >> 2), vb1 < 4 ? vb1 : 4, hs_addr);
    }
}

// CTA-wide: add the shared histogram to the global u64 table and clear it
__device__ __forceinline__ void flush(const StatsParams &P, uint8_t *smem, int tid)
{
    __syncthreads();
    for (int i = tid; i < HIST_BYTES / 4; i += NTHREADS) {
        uint32_t *p = reinterpret_cast<uint32_t *>(smem + 4 * (size_t)i);
        const uint32_t v = *p;
        if (v) {
            *p = 0u;
            const int bin = i / 96, r = i - bin * 96;                    // r >= 80: padding, never written
            const int kk = r / 40, wr = r - kk * 40;
            const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
            if (lo) gadd(P.hist, P.max_cycles, 4 * (P.w0 + wr) + kk, bin >> 6, bin & 63, lo);
            if (hi) gadd(P.hist, P.max_cycles, 4 * (P.w0 + wr) + kk + 2, bin >> 6, bin & 63, hi);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1) k_stats3(const __grid_constant__ StatsParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[WARPS][2];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int S = P.stride, R = P.tile_reads;                          // R <= 8
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    const uint32_t stage_bytes = 2u * slab_bytes;
    uint8_t *wbase = smem + HIST_BYTES + DUMMY_BYTES + (size_t)w * (2u * stage_bytes);
    uint64_t *bars = full_bar[w];
    const uint32_t ntiles = (uint32_t)((P.n + R - 1) / R);
    const uint32_t gw = blockIdx.x * WARPS + w, GW = gridDim.x * WARPS;
    const uint32_t iters = (ntiles + GW - 1) / GW;                     // the same for every warp of the grid (flush barriers)

    for (uint32_t i = tid * 16; i < (uint32_t)(HIST_BYTES + DUMMY_BYTES); i += NTHREADS * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](uint32_t tile, int s) {
        const int64_t r0 = (int64_t)tile * R;
        const int64_t left = P.n - r0;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * (uint32_t)S;
        uint8_t *dst = wbase + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&bars[s], bytes * 2u);
        bulk_g2s(dst, P.seq + r0 * S, bytes, &bars[s]);
        bulk_g2s(dst + slab_bytes, P.qual + r0 * S, bytes, &bars[s]);
    };
    if (lane == 0) {
        if (gw < ntiles) issue(gw, 0);
        if (gw + GW < ntiles) issue(gw + GW, 1);
    }

    const uint32_t hs_addr = smem_u32(smem);
    const int j = lane & 3, rr = lane >> 2;
    K3 K;
    const uint32_t zero = (uint32_t)((unsigned long long)P.n >> 62);       // 0, but only known at run time
    K.vlut_lo = VLUT_LO + zero; K.n6_lo = N6_LO + zero; K.neg_lo4 = zero - P.qk.lo4;
    const int passoff = 4 * P.w0, ncols = 4 * P.nw;
    const int nsb = P.nw > 16 ? 1 : 0;
    const int nb8 = (P.nw + 7) >> 3;
    const uint32_t kstat[4] = { 0u, 1u, 2u, 3u };
    uint32_t kdyn[4];
#pragma unroll
    for (int i = 0; i < 4; i++) kdyn[i] = (uint32_t)((j + i) & 3);
    const uint32_t uoj0 = (uint32_t)(16 * rr + 4 * j);
    const int grr = ((rr & 3) << 1) | (rr >> 2);
    const uint32_t rowoff = (uint32_t)(rr < R ? rr : 0) * (uint32_t)S + (uint32_t)passoff;
    const uint32_t wbase_addr = smem_u32(wbase);
    const bool lane_on = rr < R;

    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t tile = gw + it * GW;
        if (tile < ntiles) {
            const int s = (int)(it & 1u);
            const int64_t g = (int64_t)tile * R + rr;
            const bool active = lane_on && g < P.n;
            int L = 0;
            if (active) L = P.len ? __ldg(P.len + g) : P.uniform_len;
            const bool lenbad = active && (L <= 0 || L > S);
            if (lenbad) L = 0;
            int Lp = L - passoff;
            if (Lp > ncols) Lp = ncols;
            const int lim = Lp - 4;
            const uint32_t srow = wbase_addr + (uint32_t)s * stage_bytes + rowoff;
            const uint32_t qrow = srow + slab_bytes;
            uint32_t bad = 0;
            mbar_wait(&bars[s], (it >> 1) & 1u);

            if (nsb) {
                if (__all_sync(0xFFFFFFFFu, lim >= 124)) {
#pragma unroll
                    for (int t = 0; t < 8; t += 2) {
                        const uint32_t o0 = (uoj0 + 16u * t) & 0x7Fu, o1 = (uoj0 + 16u * t + 16u) & 0x7Fu;
                        pair_full(P, K, lds32(srow + o0), lds32(qrow + o0), o0, lds32(srow + o1), lds32(qrow + o1), o1, hs_addr, bad);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const uint32_t o = (uoj0 + 16u * t) & 0x7Fu;
                        const int vb = Lp - (int)o;
                        if (vb > 0) word_masked(P, K, lds32(srow + o), lds32(qrow + o), o, vb, hs_addr, kstat, bad);
                    }
                }
            }
            for (int b8 = nsb * 4; b8 < nb8; b8++) {
                const uint32_t o0 = 32u * (uint32_t)b8 + 4u * (uint32_t)((2 * j + grr) & 7);
                const uint32_t o1 = 32u * (uint32_t)b8 + 4u * (uint32_t)((2 * j + 1 + grr) & 7);
                const int vb0 = Lp - (int)o0, vb1 = Lp - (int)o1;
                uint32_t sw0 = 0, qw0 = 0, sw1 = 0, qw1 = 0;
                if (vb0 > 0) { sw0 = lds32(srow + o0); qw0 = lds32(qrow + o0); }
                if (vb1 > 0) { sw1 = lds32(srow + o1); qw1 = lds32(qrow + o1); }
                pair_masked(P, K, sw0, qw0, o0, vb0, sw1, qw1, o1, vb1, hs_addr, kdyn, bad);
            }
            if ((bad != 0 || lenbad) && active)
                atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));

            __syncwarp();
            if (lane == 0 && tile + 2u * GW < ntiles) issue(tile + 2u * GW, s);
        }
        if ((it + 1u) % (uint32_t)FLUSH_EVERY == 0u && it + 1u < iters) flush(P, smem, tid);     // before any u16 half can overflow
    }
    flush(P, smem, tid);
}

}  // namespace s3

cudaError_t launch_stats3(const StatsParams &p, int grid, uint32_t smem_bytes, cudaStream_t st)
{
    cudaFuncSetAttribute(s3::k_stats3, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
    s3::k_stats3<<<grid, s3::NTHREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fxg
