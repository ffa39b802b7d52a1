// fxg_extra.cu — the "next" rows of SURVEY.md §8(f-2): the loop bodies of three more tools that share the slab
// plumbing, plus stand-alone record validation (what fastx_trimmer needs: its body is pure pointer arithmetic).
//
//   K-VALIDATE  reader checks only                  src/libfastx/fastx.c:45-54,118-135,361-362
//   K-MASK      fastq_masker body                   src/fastq_masker/fastq_masker.c:92-107
//   K-ARTIFACT  fastx_artifacts_filter decision     src/fastx_artifacts_filter/fastx_artifacts_filter.c:56-114
//   K-HASN      fastq_to_fasta's test  src/fastq_to_fasta/fastq_to_fasta.c:79-82  strchr(nucleotides,'N')
//
// Streaming byte kernels on the same SoA slabs; thread-per-16-byte-chunk (mask, validate) or G lanes per read
// (artifact counts) with plain coalesced global loads — HBM-bound, no shared-memory staging needed at 1-2 passes/byte.
#include "fxg_kernels.cuh"

namespace fxg {

struct ExtraParams {
    const uint8_t *seq;
    const uint8_t *qual;        // NULL for FASTA (validate / artifacts)
    const int32_t *len;
    int32_t uniform_len, stride;
    int64_t n;
    QualK qk;                   // thr4 = mask threshold in the byte domain
    uint32_t mask4;             // mask character replicated 4x
    uint8_t *out_seq;           // K-MASK
    uint8_t *keep;              // K-ARTIFACT
    int64_t index_base;
    unsigned long long *counters;   // CNT_OUT = kept / masked reads, CNT_AUX0 = masked nucleotides
    unsigned long long chunk_magic; // ceil(2^64 / chunks): t / chunks == __umul64hi(t, chunk_magic) for t * chunks < 2^64
};

__device__ __forceinline__ int read_len(const ExtraParams &P, int64_t i) { return P.len ? __ldg(P.len + i) : P.uniform_len; }

// one thread per 16-byte chunk: validation (+ masking when MODE == 1, + "read holds an N" flag when MODE == 3)
template <int MODE>
__global__ void __launch_bounds__(256) k_chunks(const ExtraParams P)
{
    constexpr bool MASK = MODE == 1;
    constexpr bool HASN = MODE == 3;
    const int chunks = P.stride >> 4;
    const int64_t total = P.n * chunks;
    unsigned long long masked_nuc = 0;
    const bool one = chunks == 1;          // a 16-byte stride: chunk index == read index
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = one ? t : (int64_t)__umul64hi((unsigned long long)t, P.chunk_magic);   // t / chunks without the 64-bit divide
        const int c = (int)(t - i * chunks);
        const int L = read_len(P, i);
        if (L <= 0 || L > P.stride) { if (c == 0) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + i)); continue; }
        const int nb = L - 16 * c;
        const size_t off = (size_t)i * P.stride + (size_t)c * 16;
        if (nb <= 0) { if (MASK) *reinterpret_cast<uint4 *>(P.out_seq + off) = make_uint4(0, 0, 0, 0); continue; }
        const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(P.seq + off));
        uint4 q4 = make_uint4(0, 0, 0, 0);
        if (P.qual) q4 = __ldg(reinterpret_cast<const uint4 *>(P.qual + off));
        uint32_t sw[4] = { s4.x, s4.y, s4.z, s4.w };
        const uint32_t qw[4] = { q4.x, q4.y, q4.z, q4.w };
        uint32_t bad = 0, any_masked = 0, has_n = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t m = head_mask(nb - 4 * w);
            bad |= seq_bad_bits(sw[w]) & m;
            if (HASN) {                                   // zero byte of (x ^ "NNNN"), exact for any byte value
                const uint32_t z = sw[w] ^ 0x4E4E4E4Eu;
                has_n |= ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & HI & m;
            }
            if (P.qual) {
                const uint32_t xh = qw[w] | HI;
                bad |= qual_bad_bits(qw[w], xh, P.qk) & HI & m;
                if (MASK) {
                    // bytes with q < threshold take the mask character (fastq_masker.c:96-102)
                    const uint32_t low = (~qual_ge_bits(xh, P.qk)) & HI & m;       // bit7 per low byte
                    const uint32_t sel = (low >> 7) * 0xFFu;                        // 0xFF per low byte
                    sw[w] = (sw[w] & ~sel) | (P.mask4 & sel);
                    masked_nuc += __popc(low);
                    any_masked |= low;
                }
            }
            if (MASK) sw[w] &= m;
        }
        if (bad) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + i));
        if (MASK) {
            *reinterpret_cast<uint4 *>(P.out_seq + off) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
            if (any_masked) P.keep[i] = 1;          // "this read had a masked base" flag (benign race: all writers store 1)
        }
        if (HASN && has_n) P.keep[i] = 1;           // same idiom: flags start at 0, every writer stores 1
    }
    if (MASK) {
        masked_nuc = __reduce_add_sync(0xffffffffu, (unsigned)masked_nuc) ;
        if ((threadIdx.x & 31) == 0 && masked_nuc) atomicAdd(&P.counters[CNT_AUX0], masked_nuc);
    }
}

// K-ARTIFACT: 4 lanes per read, SWAR equality counts of A/C/G/T, quad reduction, keep = none reaches total-3
__global__ void __launch_bounds__(256) k_artifact(const ExtraParams P)
{
    const int j = threadIdx.x & 3;
    unsigned kept = 0;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; base < ((P.n + 63) & ~63ll); base += ((int64_t)gridDim.x * blockDim.x) >> 2) {
        const int64_t i = base;
        const bool active = i < P.n;
        int L = active ? read_len(P, i) : 0;
        const bool lenbad = active && (L <= 0 || L > P.stride);
        if (lenbad) L = 0;
        const uint8_t *srow = P.seq + (size_t)(active ? i : 0) * P.stride;
        const uint8_t *qrow = P.qual ? P.qual + (size_t)(active ? i : 0) * P.stride : NULL;
        uint32_t ca = 0, cc = 0, cg = 0, ct = 0, bad = 0;
        uint32_t ba = 0, bc = 0, bg = 0, bt = 0;      // byte-lane partial counts (POPC is a quarter-rate pipe: avoid it)
        int since = 0;
        for (int c = j; c * 16 < L; c += 4) {
            const int nb = L - 16 * c;
            const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(srow) + c);
            const uint32_t sw[4] = { s4.x, s4.y, s4.z, s4.w };
            uint4 q4 = make_uint4(0, 0, 0, 0);
            if (qrow) q4 = __ldg(reinterpret_cast<const uint4 *>(qrow) + c);
            const uint32_t qw[4] = { q4.x, q4.y, q4.z, q4.w };
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t m = head_mask(nb - 4 * w);
                const uint32_t x = sw[w];
                bad |= seq_bad_bits(x) & m;
                if (qrow) bad |= qual_bad_bits(qw[w], qw[w] | HI, P.qk) & HI & m;
                // for legal bases (all < 128): (x ^ pat) + 0x7F.. has bit7 set iff the byte differs from pat
                const uint32_t mh = m & HI;
                ba += (~((x ^ 0x41414141u) + 0x7F7F7F7Fu) & mh) >> 7;
                bc += (~((x ^ 0x43434343u) + 0x7F7F7F7Fu) & mh) >> 7;
                bg += (~((x ^ 0x47474747u) + 0x7F7F7F7Fu) & mh) >> 7;
                bt += (~((x ^ 0x54545454u) + 0x7F7F7F7Fu) & mh) >> 7;
            }
            if (++since == 48) {                       // 4 words x 48 chunks = 192 < 256 per byte lane
                ca += __dp4a(ba, ONES, 0u); cc += __dp4a(bc, ONES, 0u); cg += __dp4a(bg, ONES, 0u); ct += __dp4a(bt, ONES, 0u);
                ba = bc = bg = bt = 0; since = 0;
            }
        }
        ca += __dp4a(ba, ONES, 0u); cc += __dp4a(bc, ONES, 0u); cg += __dp4a(bg, ONES, 0u); ct += __dp4a(bt, ONES, 0u);
#pragma unroll
        for (int o = 2; o > 0; o >>= 1) {
            ca += __shfl_xor_sync(0xffffffffu, ca, o); cc += __shfl_xor_sync(0xffffffffu, cc, o);
            cg += __shfl_xor_sync(0xffffffffu, cg, o); ct += __shfl_xor_sync(0xffffffffu, ct, o);
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (active && j == 0) {
            if (bad || lenbad) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + i));
            const int lim = L - 3;      // max_allowed_different_bases = 3 (fastx_artifacts_filter.c:66,99-107)
            const bool artifact = (int)ca >= lim || (int)cc >= lim || (int)cg >= lim || (int)ct >= lim;
            P.keep[i] = artifact ? 0 : 1;
            kept += artifact ? 0u : 1u;
        }
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&P.counters[CNT_OUT], (unsigned long long)kept);
}

__global__ void k_count_flags(const uint8_t *flags, int64_t n, unsigned long long *out)
{
    unsigned c = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) c += flags[t] ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

static unsigned egrid(int64_t n, int sm) { int64_t b = (n + 255) / 256; if (b > (int64_t)sm * 32) b = (int64_t)sm * 32; if (b < 1) b = 1; return (unsigned)b; }

// op: 0 validate, 1 mask, 2 artifacts, 3 has-N flags
cudaError_t launch_extra(int op, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len, int stride, int64_t n,
                         int q_offset, int thr_q, int mask_char, uint8_t *out_seq, uint8_t *flags, int64_t index_base,
                         unsigned long long *counters, int sm_count, cudaStream_t st)
{
    ExtraParams p;
    p.seq = seq; p.qual = qual; p.len = len; p.uniform_len = uniform_len; p.stride = stride; p.n = n;
    p.qk = make_qualk(q_offset, thr_q);
    p.mask4 = (uint32_t)(uint8_t)mask_char * ONES;
    p.out_seq = out_seq; p.keep = flags; p.index_base = index_base; p.counters = counters;
    p.chunk_magic = ~0ull / (unsigned long long)(stride >> 4) + 1ull;      // stride >> 4 == 1: wraps to 0, handled below
    if ((stride >> 4) == 1) p.chunk_magic = 0;
    if (op == 0) k_chunks<0><<<egrid(n * (stride >> 4), sm_count), 256, 0, st>>>(p);
    else if (op == 1 || op == 3) {
        cudaMemsetAsync(flags, 0, (size_t)n, st);
        if (op == 1) k_chunks<1><<<egrid(n * (stride >> 4), sm_count), 256, 0, st>>>(p);
        else k_chunks<3><<<egrid(n * (stride >> 4), sm_count), 256, 0, st>>>(p);
        k_count_flags<<<egrid(n, sm_count), 256, 0, st>>>(flags, n, &counters[CNT_OUT]);
    } else k_artifact<<<egrid(n * 4, sm_count), 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fxg
