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
