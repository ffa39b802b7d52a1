// fxg_deflate.cu — SURVEY.md §8(f-4): `-z` output compressed on the GPU instead of piping the text through a forked
// gzip (src/libfastx/fastx.c:214-248).  The emitted text of a chunk is cut into 64 KB blocks; every block becomes one
// DEFLATE block with its own dynamic Huffman code over the literals (no LZ77 matches: FASTQ text is dominated by the 2-bit
// bases and the narrow quality alphabet, which an order-0 code already packs 2-3x; the byte stream any gunzip produces is
// the reference's), followed by an empty stored block so that it ends on a byte boundary and the blocks of all chunks can
// simply be concatenated (the pigz trick).  The host adds the 10-byte gzip header, the final empty block and the CRC-32 /
// ISIZE trailer; CRCs are computed here per block (table driven per 256-byte strip, then combined with carry-less
// multiplications by x^(8*256*2^j) mod P) and combined across blocks on the host.
//
//   K-DEFLATE-PLAN   per block: byte histogram -> length-limited Huffman code (one thread; 257 symbols) -> canonical codes,
//                    compressed size, block CRC
//   (scan of the block sizes)
//   K-DEFLATE-EMIT   per block: header (HLIT = 257 literal/length codes, one unused distance code, code lengths written
//                    with a flat 4-bit code-length code), the literals, end-of-block, sync marker
#include <cub/cub.cuh>
#include <stdio.h>

#include "fxg.h"
#include "fxg_kernels.cuh"

namespace fxg {

constexpr int DF_BLOCK = 65536;            // input bytes per DEFLATE block
constexpr int DF_THREADS = 256;
constexpr int DF_STRIP = DF_BLOCK / DF_THREADS;     // 256 bytes per thread
constexpr int DF_NSYM = 257;               // literals + end-of-block
constexpr uint32_t DF_POLY = 0xEDB88320u;
constexpr int DF_HEADER_BITS = 3 + 5 + 5 + 4 + 19 * 3 + (DF_NSYM + 1) * 4;     // block header + code lengths (4 bits each)

struct DeflateParams {
    const uint8_t *text;
    uint64_t bytes;
    uint32_t nblocks;
    uint32_t *codes;          // [nblocks][DF_NSYM]: (length << 16) | bit-reversed code
    uint64_t *sizes;          // [nblocks + 1] compressed bytes per block (plan), exclusive scan -> offsets (emit)
    uint32_t *crc;            // [nblocks] pure CRC-32 register of the block's bytes (init 0, no final xor)
    uint32_t xp[8];           // x^(8*256*2^j) mod P, j = 0..7
    uint8_t *out;
};

__device__ __forceinline__ uint32_t df_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0u) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ DF_POLY : b >> 1;
    }
    return p;
}

// One thread: Huffman code lengths of `freq` (DF_NSYM symbols, at least one non-zero), at most 15 bits.
// Two-queue merge over the symbols sorted by frequency; when the tree gets deeper than 15 the frequencies are halved and
// the tree rebuilt (what matters is a valid prefix code, the reference only defines the DEcompressed bytes).
__device__ void df_build_lengths(const uint32_t *freq_in, uint8_t *len_out, uint32_t *scratch)
{
    uint32_t *f = scratch;                   // [257] working frequencies
    uint16_t *order = (uint16_t *)(scratch + 260);        // [257] symbols sorted by frequency
    uint32_t *nodew = scratch + 400;         // [514] node weights: leaves (sorted order) then internal nodes
    uint16_t *parent = (uint16_t *)(scratch + 920);       // [514]
    for (int i = 0; i < DF_NSYM; i++) f[i] = freq_in[i];
    for (;;) {
        int n = 0;
        for (int i = 0; i < DF_NSYM; i++) if (f[i]) order[n++] = (uint16_t)i;
        for (int i = 0; i < DF_NSYM; i++) len_out[i] = 0;
        if (n == 1) { len_out[order[0]] = 1; return; }
        // insertion sort by frequency (n <= 257; the text alphabet is far smaller)
        for (int i = 1; i < n; i++) {
            const uint16_t s = order[i]; const uint32_t w = f[s];
            int j = i - 1;
            while (j >= 0 && f[order[j]] > w) { order[j + 1] = order[j]; j--; }
            order[j + 1] = s;
        }
        for (int i = 0; i < n; i++) nodew[i] = f[order[i]];
        int leaf = 0, inode = n, nnodes = n;          // queue heads: leaves [leaf, n), internal [inode, nnodes)
        while ((n - leaf) + (nnodes - inode) > 1) {
            int pick[2];
            for (int k = 0; k < 2; k++) {
                if (leaf < n && (inode >= nnodes || nodew[leaf] <= nodew[inode])) pick[k] = leaf++;
                else pick[k] = inode++;
            }
            nodew[nnodes] = nodew[pick[0]] + nodew[pick[1]];
            parent[pick[0]] = (uint16_t)nnodes; parent[pick[1]] = (uint16_t)nnodes;
            nnodes++;
        }
        const int root = nnodes - 1;
        int maxl = 0;
        for (int i = 0; i < n; i++) {
            int d = 0, x = i;
            while (x != root) { x = parent[x]; d++; }
            len_out[order[i]] = (uint8_t)d;
            if (d > maxl) maxl = d;
        }
        if (maxl <= 15) return;
        for (int i = 0; i < DF_NSYM; i++) if (f[i]) f[i] = (f[i] + 1u) >> 1;
    }
}

__device__ __forceinline__ uint32_t df_bitrev(uint32_t v, int n) { return __brev(v) >> (32 - n); }

__global__ void __launch_bounds__(DF_THREADS) k_deflate_plan(const DeflateParams P)
{
    __shared__ uint32_t s_freq[DF_NSYM + 3];
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_crc[DF_THREADS];
    __shared__ uint8_t s_len[DF_NSYM + 7];
    __shared__ uint32_t s_scratch[1200];
    const int tid = threadIdx.x;
    {   // CRC table (reflected CRC-32)
        uint32_t c = (uint32_t)tid;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ DF_POLY : c >> 1;
        s_tab[tid] = c;
    }
    for (uint32_t blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x) {
        for (int i = tid; i < DF_NSYM + 3; i += DF_THREADS) s_freq[i] = 0;
        __syncthreads();
        const uint64_t b0 = (uint64_t)blk * DF_BLOCK;
        const uint32_t blen = (uint32_t)((P.bytes - b0 < (uint64_t)DF_BLOCK) ? (P.bytes - b0) : DF_BLOCK);
        // strips are aligned to the END of the block (a pure CRC ignores leading zero bytes, so the front strip may be short)
        const int nstrips = (int)((blen + DF_STRIP - 1) / DF_STRIP);
        const int t_rev = tid;                                     // strip t_rev counts from the end
        uint32_t reg = 0;
        if (t_rev < nstrips) {
            const int64_t hi = (int64_t)blen - (int64_t)DF_STRIP * t_rev, lo = hi - DF_STRIP > 0 ? hi - DF_STRIP : 0;
            for (int64_t i = lo; i < hi; i++) {
                const uint32_t c = P.text[b0 + (uint64_t)i];
                atomicAdd(&s_freq[c], 1u);
                reg = s_tab[(reg ^ c) & 0xFFu] ^ (reg >> 8);
            }
        }
        s_crc[tid] = reg;                      // s_crc[t] = pure CRC of the strip t from the end
        __syncthreads();
        // tree combine: crc(A || B) = A * x^(8|B|) + B; at level j the right part is 256 * 2^j bytes long
        for (int j = 0; j < 8; j++) {
            const int step = 1 << j;
            if ((tid & (2 * step - 1)) == 0 && tid + step < DF_THREADS) {
                const uint32_t left = s_crc[tid + step], right = s_crc[tid];     // larger index = further from the end = left
                s_crc[tid] = (left ? df_multmodp(P.xp[j], left) : 0u) ^ right;
            }
            __syncthreads();
        }
        if (tid == 0) {
            P.crc[blk] = s_crc[0];
            s_freq[256] = 1;                   // end-of-block
            df_build_lengths(s_freq, s_len, s_scratch);
        }
        __syncthreads();
        if (tid == 0) {
            // canonical codes (RFC 1951 3.2.2), stored bit-reversed: DEFLATE packs Huffman codes starting from their MSB
            uint32_t bl_count[17], next_code[17];
            for (int i = 0; i < 17; i++) bl_count[i] = 0;
            for (int i = 0; i < DF_NSYM; i++) bl_count[s_len[i]]++;
            bl_count[0] = 0;
            uint32_t code = 0;
            for (int b = 1; b <= 15; b++) { code = (code + bl_count[b - 1]) << 1; next_code[b] = code; }
            uint64_t bits = DF_HEADER_BITS;
            for (int i = 0; i < DF_NSYM; i++) {
                const int l = s_len[i];
                uint32_t v = 0;
                if (l) { v = ((uint32_t)l << 16) | df_bitrev(next_code[l]++, l); bits += (uint64_t)s_freq[i] * (uint64_t)l; }
                P.codes[(size_t)blk * DF_NSYM + i] = v;
            }
            bits += 3;                                             // header of the empty stored block
            P.sizes[blk] = (bits + 7) / 8 + 4;                     // ... padded to a byte, then LEN = 0, NLEN = 0xFFFF
        }
        __syncthreads();
    }
}

// bit writer into zero-initialised global memory: `pos` = absolute bit position in P.out (which is 4-byte aligned)
__device__ __forceinline__ void df_put(uint32_t *out32, uint64_t &pos, uint64_t v, int n, uint64_t own_lo, uint64_t own_hi)
{
    // [own_lo, own_hi): words this thread writes alone (plain OR into its accumulator would need state; atomics are cheap here
    // because every word is touched by at most two threads)
    (void)own_lo; (void)own_hi;
    while (n > 0) {
        const uint32_t w = (uint32_t)(pos >> 5), sh = (uint32_t)(pos & 31u);
        const int take = n < (int)(32u - sh) ? n : (int)(32u - sh);
        const uint32_t part = (uint32_t)(v & ((take == 32) ? 0xFFFFFFFFull : ((1ull << take) - 1ull)));
        atomicOr(&out32[w], part << sh);
        v >>= take; n -= take; pos += (uint64_t)take;
    }
}

__global__ void __launch_bounds__(DF_THREADS) k_deflate_emit(const DeflateParams P)
{
    __shared__ uint32_t s_code[DF_NSYM + 3];
    __shared__ uint32_t s_bits[DF_THREADS];
    typedef cub::BlockScan<uint32_t, DF_THREADS> Scan;
    __shared__ typename Scan::TempStorage s_scan;
    const int tid = threadIdx.x;
    uint32_t *out32 = reinterpret_cast<uint32_t *>(P.out);
    for (uint32_t blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < DF_NSYM; i += DF_THREADS) s_code[i] = P.codes[(size_t)blk * DF_NSYM + i];
        __syncthreads();
        const uint64_t b0 = (uint64_t)blk * DF_BLOCK;
        const uint32_t blen = (uint32_t)((P.bytes - b0 < (uint64_t)DF_BLOCK) ? (P.bytes - b0) : DF_BLOCK);
        const uint32_t lo = (uint32_t)tid * DF_STRIP, hi = lo + DF_STRIP < blen ? lo + DF_STRIP : blen;
        uint32_t nbits = 0;
        for (uint32_t i = lo; i < hi; i++) nbits += s_code[P.text[b0 + i]] >> 16;
        uint32_t before = 0;
        Scan(s_scan).ExclusiveSum(nbits, before);
        s_bits[tid] = nbits;
        const uint64_t base = P.sizes[blk] * 8ull;                 // the block starts on a byte boundary
        if (tid == 0) {
            uint64_t pos = base;
            df_put(out32, pos, 0u, 1, 0, 0);                       // BFINAL = 0
            df_put(out32, pos, 2u, 2, 0, 0);                       // BTYPE = 10: dynamic Huffman codes
            df_put(out32, pos, 0u, 5, 0, 0);                       // HLIT: 257 literal/length codes
            df_put(out32, pos, 0u, 5, 0, 0);                       // HDIST: 1 distance code
            df_put(out32, pos, 15u, 4, 0, 0);                      // HCLEN: 19 code-length codes
            // code-length code: symbols 0..15 get 4 bits each (a complete code, canonical code = the symbol), 16..18 unused;
            // they are transmitted in the order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15
            for (int k = 0; k < 19; k++) df_put(out32, pos, k < 3 ? 0u : 4u, 3, 0, 0);
            for (int i = 0; i < DF_NSYM; i++) df_put(out32, pos, df_bitrev(s_code[i] >> 16, 4), 4, 0, 0);
            df_put(out32, pos, df_bitrev(0u, 4), 4, 0, 0);         // the one distance code: length 0 = no distance codes at all
        }
        uint64_t pos = base + (uint64_t)DF_HEADER_BITS + before;
        // pack the strip into 64-bit pieces, one atomicOr per 32-bit word
        uint64_t acc = 0; int accn = 0;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t c = s_code[P.text[b0 + i]];
            acc |= (uint64_t)(c & 0xFFFFu) << accn;
            accn += (int)(c >> 16);
            if (accn >= 32) { df_put(out32, pos, acc & 0xFFFFFFFFull, 32, 0, 0); acc >>= 32; accn -= 32; }
        }
        if (accn) df_put(out32, pos, acc, accn, 0, 0);
        __syncthreads();
        if (tid == DF_THREADS - 1) {
            // end-of-block, then the empty stored block: 3 header bits, pad to a byte, LEN = 0x0000, NLEN = 0xFFFF
            uint64_t p2 = base + (uint64_t)DF_HEADER_BITS + before + nbits;
            const uint32_t eob = s_code[256];
            df_put(out32, p2, eob & 0xFFFFu, (int)(eob >> 16), 0, 0);
            df_put(out32, p2, 0u, 3, 0, 0);
            p2 = (p2 + 7) & ~7ull;
            df_put(out32, p2, 0xFFFF0000ull, 32, 0, 0);
        }
    }
}

}  // namespace fxg

using namespace fxg;

// host side of the CRC algebra (same polynomial arithmetic as df_multmodp)
static uint32_t h_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0u) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ DF_POLY : b >> 1;
    }
    return p;
}
static uint32_t h_xpow8n(uint64_t nbytes)      // x^(8 n) mod P
{
    static uint32_t x2n[64];
    static int init = 0;
    if (!init) { x2n[0] = 1u << 30; for (int k = 1; k < 64; k++) x2n[k] = h_multmodp(x2n[k - 1], x2n[k - 1]); init = 1; }
    uint32_t p = 1u << 31;
    int k = 3;
    while (nbytes) { if (nbytes & 1) p = h_multmodp(x2n[k & 63], p); nbytes >>= 1; k++; }
    return p;
}
// pure CRC of A || B from the pure CRCs of A and B
extern "C" uint32_t fxg_crc32_concat(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    return (len_b ? h_multmodp(h_xpow8n(len_b), crc_a) : crc_a) ^ crc_b;
}
// the CRC-32 gzip stores, from the pure CRC of the whole stream and its length
extern "C" uint32_t fxg_crc32_finish(uint32_t crc_pure, uint64_t total_len)
{
    return (fxg_crc32_concat(0xFFFFFFFFu, crc_pure, total_len)) ^ 0xFFFFFFFFu;
}

namespace fxg {

size_t deflate_scratch_bytes(size_t max_text_bytes)
{
    const size_t nb = (max_text_bytes + DF_BLOCK - 1) / DF_BLOCK + 1;
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scan, (uint64_t *)NULL, (uint64_t *)NULL, (int)nb + 1);
    return nb * DF_NSYM * 4 + (nb + 1) * 8 * 2 + nb * 4 + scan + 1024;
}
size_t deflate_out_bound(size_t text_bytes)
{
    const size_t nb = (text_bytes + DF_BLOCK - 1) / DF_BLOCK;
    return text_bytes + text_bytes / 8 + nb * 160 + 64;           // 9 bits per literal at the very worst, header, sync marker
}

// text (device) -> concatenated byte-aligned DEFLATE blocks in out (device, 4-byte aligned, capacity deflate_out_bound()).
// Returns the compressed size and the pure CRC of the text through host pointers after a stream synchronisation.
cudaError_t deflate_run(const uint8_t *d_text, size_t bytes, uint8_t *d_out, void *d_scratch, size_t scratch_bytes, uint64_t *h_pinned /* >= 2 words */,
                        uint32_t *h_crc_blocks /* pinned, >= nblocks */, size_t *out_bytes, uint32_t *crc_pure, int sm_count, cudaStream_t st,
                        int64_t *launches)
{
    *out_bytes = 0; *crc_pure = 0;
    if (bytes == 0) return cudaSuccess;
    const uint32_t nb = (uint32_t)((bytes + DF_BLOCK - 1) / DF_BLOCK);
    DeflateParams p;
    p.text = d_text; p.bytes = bytes; p.nblocks = nb; p.out = d_out;
    uint8_t *s = (uint8_t *)d_scratch;
    p.codes = (uint32_t *)s; s += (size_t)nb * DF_NSYM * 4;
    s = (uint8_t *)(((uintptr_t)s + 15) & ~(uintptr_t)15);
    uint64_t *sizes = (uint64_t *)s; s += (size_t)(nb + 1) * 8;
    uint64_t *offs = (uint64_t *)s; s += (size_t)(nb + 1) * 8;
    p.crc = (uint32_t *)s; s += (size_t)nb * 4;
    s = (uint8_t *)(((uintptr_t)s + 15) & ~(uintptr_t)15);
    void *scan_tmp = s;
    size_t scan_bytes = scratch_bytes - (size_t)(s - (uint8_t *)d_scratch);
    for (int j = 0; j < 8; j++) p.xp[j] = h_xpow8n((uint64_t)DF_STRIP << j);
    p.sizes = sizes;
    const unsigned grid = nb < (unsigned)sm_count * 4u ? nb : (unsigned)sm_count * 4u;
    cudaError_t e;
    if ((e = cudaMemsetAsync(sizes + nb, 0, 8, st)) != cudaSuccess) return e;
    k_deflate_plan<<<grid, DF_THREADS, 0, st>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, sizes, offs, (int)nb + 1, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(h_pinned, offs + nb, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(h_crc_blocks, p.crc, (size_t)nb * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    const size_t total = (size_t)h_pinned[0];
    if ((e = cudaMemsetAsync(d_out, 0, (total + 7) & ~(size_t)3, st)) != cudaSuccess) return e;
    p.sizes = offs;
    k_deflate_emit<<<grid, DF_THREADS, 0, st>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 4;
    uint32_t crc = 0;
    for (uint32_t b = 0; b < nb; b++) {
        const uint64_t bl = (b + 1 == nb) ? bytes - (uint64_t)b * DF_BLOCK : DF_BLOCK;
        crc = fxg_crc32_concat(crc, h_crc_blocks[b], bl);
    }
    *out_bytes = total; *crc_pure = crc;
    return cudaSuccess;
}

}  // namespace fxg
