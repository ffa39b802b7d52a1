// fxg_stats.cu — K-STATS fallback: per-cycle x nucleotide x quality histograms for fastx_quality_stats
// (src/fastx_quality_stats/fastx_quality_stats.c:166-216 `read_file`) with global atomics.  Everything the tool prints
// (count/min/max/sum/mean/quartiles/whiskers, per nucleotide) derives from hist[cycle][nuc][q+15],
// so the device only counts; the host derives and prints.
//
// The fast kernel is k_stats4 (fxg_stats4.cu: shared-memory histogram, lane = read).  This one takes what that layout
// does not: FASTA input (no qualities), per-read weights (collapsed "N-COUNT" identifiers), quality offsets above 79 and
// strides whose tiles do not fit beside the histogram.
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

cudaError_t launch_stats_simple(const StatsParams &p, int sm_count, cudaStream_t st)
{
    const int64_t total = p.n * (p.stride >> 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    if (blocks < 1) blocks = 1;
    k_stats_simple<<<(unsigned)blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fxg
