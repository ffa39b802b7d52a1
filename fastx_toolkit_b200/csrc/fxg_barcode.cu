// fxg_barcode.cu — K-BARCODE: the matching loop of the reference's barcode splitter,
//   scripts/fastx_barcode_splitter.pl:208-290 (match_sequences) and :296 (mismatch_count),
// one thread per read.  The reference is Perl: for every read it compares a fragment of the sequence (its
// first or last B characters, B = barcode length) with every entry of the barcode list (each barcode followed
// by its `--partial` forms, :170-176), counts mismatches by XOR-ing the two strings and keeps the FIRST entry
// with the lowest count ('<', :262), provided the count is <= --mismatches; otherwise the read is 'unmatched'.
//
// What exactly the Perl computes (kept bit for bit, quirks included):
//   mismatch_count(f, b) = length(f) - #{ p < min(|f|,|b|) : f[p] == b[p] }      (a string XOR pads the shorter
//                          operand with NULs, and only NUL bytes of the result count as equal)
//   mm = mismatch_count(fragment, entry) + (B - |entry|)                          (:258-260)
// so a shortened (partial) entry is aligned at the START of the fragment for --bol and --eol alike, a read
// shorter than B is compared over its own length only, and an empty sequence line matches the first barcode
// with 0 mismatches.  best starts at B, so an entry can only win with mm < B.
//
// Layout: fragments are the rows of an ordinary batch (seq = fragment bytes, len = fragment length, stride a
// multiple of 16, <= 64 here); the entry table lives in shared memory (rows of the same stride, zero padded).
#include <string.h>

#include "fxg_kernels.cuh"

namespace fxg {

template <int WORDS>
__global__ void __launch_bounds__(256) k_barcode(const BarcodeParams P)
{
    __shared__ __align__(16) uint32_t s_ent[BC_TILE_ENTRIES * WORDS];
    __shared__ int32_t s_len[BC_TILE_ENTRIES];
    const int ne = P.e1 - P.e0;
    for (int i = threadIdx.x; i < ne * WORDS; i += blockDim.x)
        s_ent[i] = reinterpret_cast<const uint32_t *>(P.entries + (size_t)P.e0 * P.stride)[i];
    for (int i = threadIdx.x; i < ne; i += blockDim.x) s_len[i] = P.elen[P.e0 + i];
    __syncthreads();

    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n; r += (int64_t)gridDim.x * blockDim.x) {
        uint32_t f[WORDS];
        const uint4 *src = reinterpret_cast<const uint4 *>(P.frag + (size_t)r * P.stride);
#pragma unroll
        for (int q = 0; q < WORDS / 4; q++) {
            const uint4 v = __ldg(src + q);
            f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
        }
        const int fl = __ldg(P.flen + r);
        int bmm = P.e0 == 0 ? P.barcode_len : P.best_mm[r];
        int bidx = P.e0 == 0 ? -1 : P.best[r];
        for (int e = 0; e < ne; e++) {
            const int el = s_len[e];
            const int ml = fl < el ? fl : el;                  // positions both strings have
            int differ = 0;
#pragma unroll
            for (int w = 0; w < WORDS; w++) {
                const uint32_t d = f[w] ^ s_ent[e * WORDS + w];
                // bit 7 of a byte of nz is set iff that byte of d is non-zero (exact for any byte value)
                const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & HI;
                differ += __popc(nz & head_mask(ml - 4 * w));
            }
            const int mm = (fl - ml) + differ + (P.barcode_len - el);
            if (mm < bmm) { bmm = mm; bidx = P.e0 + e; }       // strict '<': the first best entry keeps the read
        }
        if (P.last) P.best[r] = (bidx >= 0 && bmm <= P.allowed) ? bidx : -1;
        else { P.best_mm[r] = bmm; P.best[r] = bidx; }
    }
}

cudaError_t launch_barcode(const BarcodeParams &p, int sm_count, cudaStream_t st)
{
    int64_t blocks = (p.n + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    if (blocks < 1) blocks = 1;
    switch (p.stride) {
    case 16: k_barcode<4><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 32: k_barcode<8><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 48: k_barcode<12><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    case 64: k_barcode<16><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace fxg
