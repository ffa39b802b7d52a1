// fxg_text.cu — the steps on either side of the per-read loop, moved to the GPU so the host does no per-byte work
// (SURVEY.md §8f-1): raw FASTQ text in, FASTQ text out.
//
//   K-LINES  newline index of a text chunk                              (reader: fgets/chomp, fastx.c:324-378)
//   K-RECS   4 lines -> record table + structural checks               (fastx.c:331-347,361-362,382-390)
//   K-PACK   sequence / quality lines -> the SoA slabs the op kernels use
//   <op>     K-TRIM / K-FILTER (fused validation) on the slabs — the kernels of fxg_kernels.cu, unchanged
//   K-EMIT   surviving records -> output text "@name\nSEQ[:len]\n+name2\nQUAL[:len]\n"   (fastx.c:440-473)
//
// Record formats: 4-line FASTQ with ASCII qualities, 4-line FASTQ with NUMERIC qualities (a chunk whose every record has
// a quality line of another length than its sequence: fastx.c:382-390 decides per record, K-NUMQ parses the numbers as
// convert_numeric_quality_score_line does, fastx.c:137-167, and K-EMIT prints them back as write_numeric_qual_string does,
// fastx.c:421-438), and 2-line FASTA (fastx.c:348-352; collapsed "N-COUNT" identifiers give the read weights,
// fastx.c:475-497).
// Anything the fast path does not handle bit-exactly by construction — a structural problem, a chunk that mixes ASCII and
// numeric records, a malformed number, an illegal base/quality, an over-long line — is reported as an *anomaly* with the
// index of the first affected record; nothing is emitted for that chunk and the host re-reads it with the (slower)
// host parser, which reproduces the reference's output prefix and error message exactly.
// Scans are CUB (library primitive); everything else is hand-written.
#include <cub/cub.cuh>
#include <stdio.h>
#include <string.h>

#include "fxg.h"
#include "fxg_kernels.cuh"

namespace fxg {

// ---- K-LINES ---------------------------------------------------------------------------------------------
// each thread owns 64 bytes: count newlines, then (after a scan) write their positions
__global__ void __launch_bounds__(256) k_nl_count(const uint8_t *text, uint64_t bytes, uint32_t *cnt, uint64_t nthreads_total)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads_total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b0 = t * 64;
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t off = b0 + 16 * k;
            if (off + 16 <= bytes) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + off));
                const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t x = w[i] ^ 0x0A0A0A0Au;                         // zero byte <=> '\n'
                    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
                    c += __popc(z);
                }
            } else {
                for (uint64_t p = off; p < bytes && p < off + 16; p++) c += (text[p] == '\n');
            }
        }
        cnt[t] = c;
    }
}

__global__ void __launch_bounds__(256) k_nl_scatter(const uint8_t *text, uint64_t bytes, const uint32_t *scan, uint32_t *line_end,
                                                    uint64_t nthreads_total)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads_total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b0 = t * 64;
        uint32_t o = scan[t];
        const uint64_t e = (b0 + 64 < bytes) ? b0 + 64 : bytes;
        for (uint64_t p = b0; p < e; p += 4) {
            if (p + 4 <= e) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(text + p));
                const uint32_t x = w ^ 0x0A0A0A0Au;
                uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // bit7 of each '\n' byte
                while (z) {
                    const int b = (__ffs(z) - 1) >> 3;
                    line_end[o++] = (uint32_t)(p + b);
                    z &= z - 1;
                }
            } else {
                for (uint64_t q = p; q < e; q++) if (text[q] == '\n') line_end[o++] = (uint32_t)q;
            }
        }
    }
}

// ---- K-RECS ------------------------------------------------------------------------------------------------
struct RecTable {
    uint32_t *start;     // [n_rec*lpr] start offset of each line
    uint32_t *llen;      // [n_rec*lpr] length of each line (CR removed)
    int lpr;             // lines per record: 4 (FASTQ) or 2 (FASTA)
};

enum { AN_NONE = 0, AN_PREFIX = 1, AN_EMPTY_SEQ = 2, AN_QUAL_LEN = 3, AN_LONG_LINE = 4, AN_BAD_RECORD = 5, AN_LINE_COUNT = 6 };
// scalars block (u64 words): [0] first anomaly (record << 8 | class, min), [1] records kept, [2] max_len (int), [3] min_len (int),
//                            [4] records whose quality line has another length than the sequence (numeric candidates),
//                            [5] first such record (min), [6] sum of the read weights in, [7] sum of the read weights kept,
//                            [8..13] clipper on FASTA: read weights per FXG_CLIP_* class
enum { SC_ANOM = 0, SC_KEPT = 1, SC_MAXLEN = 2, SC_MINLEN = 3, SC_NNUM = 4, SC_FIRSTNUM = 5, SC_WIN = 6, SC_WKEPT = 7, SC_CLASS = 8 /* 6 words */, SC_WORDS = 16 };

// get_reads_count() (fastx.c:475-497) of a FASTA identifier: the number after the first '-', if positive, else 1
__device__ __forceinline__ int32_t reads_count_dev(const uint8_t *name, uint32_t n)
{
    uint32_t i = 0;
    while (i < n && name[i] != '-') i++;
    if (i >= n) return 1;
    i++;
    while (i < n && (name[i] == ' ' || (name[i] >= 9 && name[i] <= 13))) i++;      // atoi skips white space
    bool neg = false;
    if (i < n && (name[i] == '+' || name[i] == '-')) { neg = name[i] == '-'; i++; }
    long long v = 0;
    while (i < n && name[i] >= '0' && name[i] <= '9') { v = v * 10 + (name[i] - '0'); if (v > 0x7FFFFFFFll) v = 0x7FFFFFFFll; i++; }
    if (neg) v = -v;
    return v > 0 ? (int32_t)v : 1;
}

template <int LPR>
__global__ void __launch_bounds__(256) k_recs(const uint8_t *text, const uint32_t *line_end, uint32_t n_rec, RecTable rt,
                                              int32_t *seq_len, int32_t *weight, unsigned long long *sc)
{
    int local_max = 0, local_min = 0x7FFFFFFF;
    unsigned long long wsum = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        uint32_t st[LPR], ln[LPR];
#pragma unroll
        for (int k = 0; k < LPR; k++) {
            const uint32_t li = LPR * r + k;
            const uint32_t s = li == 0 ? 0u : line_end[li - 1] + 1u;
            uint32_t e = line_end[li];
            uint32_t l = e - s;
            if (l > 0 && text[e - 1] == '\r') l--;            // chomp(): CRLF
            st[k] = s; ln[k] = l;
            rt.start[li] = s; rt.llen[li] = l;
        }
        int an = AN_NONE;
        if (ln[0] == 0 || text[st[0]] != (LPR == 4 ? '@' : '>')) an = AN_PREFIX;
        else if (ln[1] == 0) an = AN_EMPTY_SEQ;
        else if (ln[0] >= 24998u || ln[1] >= 24998u || (LPR == 4 && (ln[2] >= 24998u || ln[3] >= 24998u))) an = AN_LONG_LINE;
        else {
            // chomp() cuts a line at its FIRST CR (src/libfastx/chomp.c:34-44): a CR inside a name line changes what the
            // reference writes back, so such a record goes to the host parser (sequence / quality lines: a CR is an illegal
            // character there and the validation of the op kernels reports it)
            for (uint32_t i = 0; i < ln[0]; i++) if (text[st[0] + i] == '\r') an = AN_BAD_RECORD;
            if (LPR == 4) for (uint32_t i = 0; i < ln[2]; i++) if (text[st[2] + i] == '\r') an = AN_BAD_RECORD;
        }
        if (an != AN_NONE) atomicMin(&sc[SC_ANOM], ((unsigned long long)r << 8) | (unsigned long long)an);
        else if (LPR == 4 && ln[3] != ln[1]) {                // numeric quality line, or a broken record: the host decides per chunk
            atomicAdd(&sc[SC_NNUM], 1ull);
            atomicMin(&sc[SC_FIRSTNUM], (unsigned long long)r);
        }
        seq_len[r] = (int32_t)ln[1];
        if (LPR == 2) {
            const int32_t w = reads_count_dev(text + st[0] + 1, ln[0] ? ln[0] - 1 : 0);
            weight[r] = w;
            wsum += (unsigned long long)w;
        }
        if ((int)ln[1] > local_max && an == AN_NONE) local_max = (int)ln[1];
        if ((int)ln[1] < local_min && an == AN_NONE) local_min = (int)ln[1];
    }
    local_max = __reduce_max_sync(0xffffffffu, local_max);
    if ((threadIdx.x & 31) == 0 && local_max > 0) atomicMax((int *)&sc[SC_MAXLEN], local_max);
    local_min = __reduce_min_sync(0xffffffffu, local_min);
    if ((threadIdx.x & 31) == 0 && local_min != 0x7FFFFFFF) atomicMin((int *)&sc[SC_MINLEN], local_min);
    if (LPR == 2) {
        for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&sc[SC_WIN], wsum);
    }
}

// ---- K-NUMQ: numeric quality lines -> quality bytes (value + 33), one thread per record ------------------------
// convert_numeric_quality_score_line (fastx.c:137-167): strtol tokens (leading white space, optional sign, digits), every
// value in [-15, 93], as many values as bases.  Anything else is the host parser's business (it words the message).
__global__ void __launch_bounds__(128) k_numq(const uint8_t *text, RecTable rt, uint32_t n_rec, int stride, uint8_t *qual,
                                              unsigned long long *sc)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const uint8_t *p = text + rt.start[4 * r + 3];
        const uint32_t n = rt.llen[4 * r + 3], nb = rt.llen[4 * r + 1];
        uint8_t *row = qual + (size_t)r * stride;
        uint32_t i = 0, idx = 0;
        bool ok = n > 0;
        while (ok) {
            while (i < n && (p[i] == ' ' || (p[i] >= 9 && p[i] <= 13))) i++;
            bool neg = false;
            if (i < n && (p[i] == '+' || p[i] == '-')) { neg = p[i] == '-'; i++; }
            if (i >= n || p[i] < '0' || p[i] > '9') { ok = false; break; }      // strtol consumed nothing
            int v = 0;
            while (i < n && p[i] >= '0' && p[i] <= '9') { v = v * 10 + (p[i] - '0'); if (v > 100000) v = 100000; i++; }
            if (neg) v = -v;
            if (v > 93 || v < -15 || idx >= nb) { ok = false; break; }
            row[idx++] = (uint8_t)(v + 33);
            if (i >= n) break;                                                   // the line ends right after a number
        }
        if (!ok || idx != nb) { atomicMin(&sc[SC_ANOM], ((unsigned long long)r << 8) | (unsigned long long)AN_BAD_RECORD); idx = 0; }
        for (uint32_t k = idx; k < (uint32_t)stride; k++) row[k] = 0;
    }
}

// ---- K-PACK: one thread per 16-byte destination chunk (both rows) ---------------------------------------------
__device__ __forceinline__ uint4 load_unaligned16(const uint8_t *text, uint64_t src, int nbytes /* 1..16 valid */)
{
    // src is arbitrary: read the aligned words covering it and funnel-shift them into place
    const uint64_t a = src & ~3ull;
    const int sh = (int)(src & 3ull) * 8;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(text + a);
    const int nwords = (nbytes + (int)(src & 3ull) + 3) >> 2;       // words touched (<= 5)
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = (i < nwords) ? __ldg(p + i) : 0u;
    uint4 o;
    o.x = __funnelshift_r(w[0], w[1], sh);
    o.y = __funnelshift_r(w[1], w[2], sh);
    o.z = __funnelshift_r(w[2], w[3], sh);
    o.w = __funnelshift_r(w[3], w[4], sh);
    return o;
}

// qual == NULL: sequence rows only (FASTA, or numeric qualities which K-NUMQ fills)
__global__ void __launch_bounds__(256) k_pack(const uint8_t *text, RecTable rt, uint32_t n_rec, int stride, uint8_t *seq, uint8_t *qual)
{
    const int lpr = rt.lpr;
    const int chunks = stride >> 4;
    const uint64_t total = (uint64_t)n_rec * chunks;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / chunks);
        const int c = (int)(t - (uint64_t)r * chunks);
        const int L = (int)rt.llen[lpr * r + 1];
        const int nb = L - 16 * c;
        uint4 s = make_uint4(0, 0, 0, 0), q = make_uint4(0, 0, 0, 0);
        if (nb > 0) {
            const int n = nb < 16 ? nb : 16;
            s = load_unaligned16(text, (uint64_t)rt.start[lpr * r + 1] + 16u * c, n);
            if (qual) q = load_unaligned16(text, (uint64_t)rt.start[4 * r + 3] + 16u * c, n);
        }
        const size_t off = (size_t)r * stride + (size_t)c * 16;
        *reinterpret_cast<uint4 *>(seq + off) = s;
        if (qual) *reinterpret_cast<uint4 *>(qual + off) = q;
    }
}

// ---- K-EMIT ----------------------------------------------------------------------------------------------------
// out_len[r] < 0: record dropped.  keep_flags != NULL (filter): record kept iff flag, emitted at full length.
// numq != NULL: numeric qualities — the quality line is printed from the slab row (value + 33 per byte) as
// write_numeric_qual_string does (fastx.c:421-438): "%d" joined by single spaces.
__device__ __forceinline__ int numq_digits(int v) { return v < 0 ? (v <= -10 ? 3 : 2) : (v >= 10 ? 2 : 1); }

__global__ void __launch_bounds__(256) k_emit_sizes(RecTable rt, uint32_t n_rec, const int32_t *out_len, const uint8_t *keep_flags,
                                                    const uint8_t *numq, int numq_stride, uint64_t *sizes)
{
    const int lpr = rt.lpr;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        int ol = keep_flags ? (keep_flags[r] ? (int)rt.llen[lpr * r + 1] : -1) : out_len[r];
        uint64_t sz = 0;
        if (ol >= 0) {
            sz = (uint64_t)rt.llen[lpr * r] + 1ull + (uint64_t)ol + 1ull;
            if (lpr == 4) {
                sz += (uint64_t)(rt.llen[4 * r + 2] ? rt.llen[4 * r + 2] : 1u) + 1ull;
                if (numq) {
                    const uint8_t *row = numq + (size_t)r * numq_stride;
                    uint64_t q = ol > 0 ? (uint64_t)(ol - 1) : 0ull;                 // the spaces
                    for (int i = 0; i < ol; i++) q += (uint64_t)numq_digits((int)row[i] - 33);
                    sz += q + 1ull;
                } else {
                    sz += (uint64_t)ol + 1ull;
                }
            }
        }
        sizes[r] = sz;
    }
}

__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, int n, int lane)
{
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// one warp per record
__global__ void __launch_bounds__(256) k_emit(const uint8_t *text, RecTable rt, uint32_t n_rec, const int32_t *out_len,
                                              const uint8_t *keep_flags, const uint64_t *offs, uint8_t *out,
                                              const uint8_t *alt_seq, const uint8_t *alt_qual, int alt_stride,
                                              const uint8_t *numq, int numq_stride)
{
    const int lane = threadIdx.x & 31;
    const int lpr = rt.lpr;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < n_rec; r += nwarps) {
        const int ol = keep_flags ? (keep_flags[r] ? (int)rt.llen[lpr * r + 1] : -1) : out_len[r];
        if (ol < 0) continue;
        uint8_t *o = out + offs[r];
        const int l0 = (int)rt.llen[lpr * r];
        warp_copy(o, text + rt.start[lpr * r], l0, lane);               // "@name" / ">name"
        if (lane == 0) o[l0] = '\n';
        o += l0 + 1;
        warp_copy(o, alt_seq ? alt_seq + (size_t)r * alt_stride : text + rt.start[lpr * r + 1], ol, lane);   // SEQ[:len]
        if (lane == 0) o[ol] = '\n';
        o += ol + 1;
        if (lpr != 4) continue;
        const int l2 = (int)rt.llen[4 * r + 2];
        if (l2 > 0) {                                                   // "+name2": first byte is always written as '+'
            warp_copy(o, text + rt.start[4 * r + 2], l2, lane);
            if (lane == 0) { o[0] = '+'; o[l2] = '\n'; }
            o += l2 + 1;
        } else {
            if (lane == 0) { o[0] = '+'; o[1] = '\n'; }
            o += 2;
        }
        if (!numq) {
            warp_copy(o, alt_qual ? alt_qual + (size_t)r * alt_stride : text + rt.start[4 * r + 3], ol, lane);  // QUAL[:len]
            if (lane == 0) o[ol] = '\n';
        } else {
            const uint8_t *row = numq + (size_t)r * numq_stride;
            int base = 0;                                               // characters written so far
            for (int i0 = 0; i0 < ol; i0 += 32) {
                const int i = i0 + lane;
                const int v = i < ol ? (int)row[i] - 33 : 0;
                const int w = i < ol ? numq_digits(v) + (i + 1 < ol ? 1 : 0) : 0;      // digits (+ the space behind them)
                int incl = w;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
                if (i < ol) {
                    uint8_t *q = o + base + incl - w;
                    int a = v < 0 ? -v : v;
                    if (v < 0) *q++ = '-';
                    if (a >= 10) { *q++ = (uint8_t)('0' + a / 10); a %= 10; }
                    *q++ = (uint8_t)('0' + a);
                    if (i + 1 < ol) *q = ' ';
                }
                base += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) o[base] = '\n';
        }
    }
}

// clipper: what the tool emits per record (fastx_clipper.cpp:280-319): normally the WRITE class at its clipped length;
// with -k only the adapter-only reads, untruncated
__global__ void k_clip_emit_len(const int32_t *clip_len, const uint8_t *cls, const int32_t *seq_len, uint32_t n_rec, int show_adapter_only,
                                int32_t *emit_len)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        int e = -1;
        if (show_adapter_only) { if (cls[r] == FXG_CLIP_ADAPTER_ONLY) e = seq_len[r]; }
        else if (cls[r] == FXG_CLIP_WRITE) e = clip_len[r];
        emit_len[r] = e;
    }
}

// fastx_clipper counts get_reads_count() per class (fastx_clipper.cpp:259,277-312): for FASTA the classes are weighted
__global__ void k_clip_class_weights(const uint8_t *cls, const int32_t *weight, uint32_t n_rec, unsigned long long *sc)
{
    unsigned long long w[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const int c = cls[r];
        const unsigned long long v = (unsigned long long)weight[r];
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] += (c == k) ? v : 0ull;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        for (int o = 16; o; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
        if ((threadIdx.x & 31) == 0 && w[k]) atomicAdd(&sc[SC_CLASS + k], w[k]);
    }
}

// decisions-only output: one int32 per record (surviving length, -1 = dropped)
__global__ void k_decisions(const int32_t *out_len, const uint8_t *keep_flags, const int32_t *seq_len, uint32_t n_rec, int32_t *dec)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x)
        dec[r] = keep_flags ? (keep_flags[r] ? seq_len[r] : -1) : out_len[r];
}

__global__ void k_count_kept(const int32_t *out_len, const uint8_t *keep_flags, const int32_t *weight, uint32_t n_rec, unsigned long long *sc)
{
    unsigned c = 0;
    unsigned long long w = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const bool k = keep_flags ? keep_flags[r] != 0 : out_len[r] >= 0;
        c += k ? 1u : 0u;
        if (k && weight) w += (unsigned long long)weight[r];
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sc[SC_KEPT], (unsigned long long)c);
    if (weight) {
        for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sc[SC_WKEPT], w);
    }
}

}  // namespace fxg

using namespace fxg;

// provided by fxg_api.cu
extern "C" int fxg_internal_scan_on_stream(fxg_ctx *ctx, int mode, const fxg_batch *b, int q_offset, int thr_q, int min_len,
                                           int min_percent, void *out, void *stream);
extern "C" void *fxg_internal_counters(fxg_ctx *ctx);
extern "C" int fxg_internal_revcomp_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *oseq, uint8_t *oqual, void *stream);
extern "C" int fxg_internal_clip_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, const fxg_clip_opts *o, int32_t *out_len,
                                           uint8_t *out_class, void *stream);
extern "C" int fxg_internal_stats_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist, int32_t max_cycles,
                                            const int32_t *weight_dev, void *stream);
extern "C" int fxg_collapse_add_checked(fxg_collapser *c, const fxg_batch *b, const int32_t *weight, int64_t first_base, int64_t *first_bad_row);

struct fxg_text {
    fxg_ctx *ctx;
    int device;
    cudaStream_t st;
    int fasta;                     // input records: 0 = 4-line FASTQ, 1 = 2-line FASTA
    size_t cap_bytes;              // text capacity
    uint8_t *d_text, *d_out;
    uint32_t *d_cnt, *d_scan;      // per-64-byte newline counts
    uint32_t *d_line_end;  size_t cap_lines;
    uint32_t *d_start, *d_llen;    // record table (lpr per record)
    int32_t *d_seq_len, *d_out_len, *d_weight; uint8_t *d_keep;
    uint64_t *d_sizes, *d_offs;
    size_t cap_recs;
    uint8_t *d_seq, *d_qual;  size_t cap_slab;
    uint8_t *d_oseq, *d_oqual; size_t cap_oslab;   // revcomp output rows
    void *d_tmp; size_t tmp_bytes;
    unsigned long long *d_scalars; // SC_WORDS words, see K-RECS
    unsigned long long *h_scalars; // pinned mirror (+ 8 words for the context's counters)
    int64_t launches;
    int64_t n_numeric_chunks, n_fasta_chunks;      // chunks that took the numeric-quality / FASTA forms of the path
    int32_t *decide_len_host;      // fxg_text_decide_host: per-record decision goes here instead of any text
    uint32_t *decide_start_host;
    int deflate;                   // emit DEFLATE blocks instead of plain text (fxg_deflate.cu)
    uint8_t *d_dfl; void *d_dfl_scratch; size_t dfl_scratch_bytes; uint32_t *h_dfl_crc;
    char err[256];
};

#define CKT(t, call)                                                                               \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((t)->err, sizeof((t)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return FXG_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static unsigned tgrid(uint64_t n, unsigned per_block = 256) { uint64_t b = (n + per_block - 1) / per_block; if (b > 148ull * 32) b = 148ull * 32; if (b < 1) b = 1; return (unsigned)b; }

extern "C" const char *fxg_text_error(const fxg_text *t) { return t ? t->err : "no text context"; }
extern "C" int64_t fxg_text_launches(const fxg_text *t) { return t ? t->launches : 0; }
extern "C" int64_t fxg_text_numeric_chunks(const fxg_text *t) { return t ? t->n_numeric_chunks : 0; }
extern "C" int64_t fxg_text_fasta_chunks(const fxg_text *t) { return t ? t->n_fasta_chunks : 0; }
namespace fxg {
size_t deflate_scratch_bytes(size_t max_text_bytes);
size_t deflate_out_bound(size_t text_bytes);
cudaError_t deflate_run(const uint8_t *d_text, size_t bytes, uint8_t *d_out, void *d_scratch, size_t scratch_bytes, uint64_t *h_pinned,
                        uint32_t *h_crc_blocks, size_t *out_bytes, uint32_t *crc_pure, int sm_count, cudaStream_t st, int64_t *launches);
}

// `-z`: the emitted text leaves the GPU as byte-aligned DEFLATE blocks (see fxg_deflate.cu and fxg.h)
extern "C" int fxg_text_set_deflate(fxg_text *t, int on)
{
    if (!t) return FXG_ERR_ARG;
    if (on && !t->d_dfl) {
        CKT(t, cudaSetDevice(t->device));
        const size_t text_cap = t->cap_bytes + t->cap_bytes / 4 + 64;
        t->dfl_scratch_bytes = deflate_scratch_bytes(text_cap);
        CKT(t, cudaMalloc(&t->d_dfl, deflate_out_bound(text_cap)));
        CKT(t, cudaMalloc(&t->d_dfl_scratch, t->dfl_scratch_bytes));
        CKT(t, cudaMallocHost(&t->h_dfl_crc, (text_cap / 65536 + 2) * sizeof(uint32_t)));
    }
    t->deflate = on ? 1 : 0;
    return FXG_OK;
}

extern "C" int fxg_text_set_format(fxg_text *t, int fasta)
{
    if (!t) return FXG_ERR_ARG;
    t->fasta = fasta ? 1 : 0;
    return FXG_OK;
}

extern "C" void fxg_text_free(fxg_text *t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    cudaFree(t->d_text); cudaFree(t->d_out); cudaFree(t->d_cnt); cudaFree(t->d_scan); cudaFree(t->d_line_end);
    cudaFree(t->d_start); cudaFree(t->d_llen); cudaFree(t->d_seq_len); cudaFree(t->d_out_len); cudaFree(t->d_weight); cudaFree(t->d_keep);
    cudaFree(t->d_sizes); cudaFree(t->d_offs); cudaFree(t->d_seq); cudaFree(t->d_qual); cudaFree(t->d_oseq); cudaFree(t->d_oqual); cudaFree(t->d_tmp);
    cudaFree(t->d_scalars); cudaFreeHost(t->h_scalars);
    cudaFree(t->d_dfl); cudaFree(t->d_dfl_scratch); if (t->h_dfl_crc) cudaFreeHost(t->h_dfl_crc);
    if (t->st) cudaStreamDestroy(t->st);
    free(t);
}

extern "C" int fxg_text_new(fxg_ctx *ctx, int device, size_t max_chunk_bytes, fxg_text **out)
{
    if (!ctx || !out || max_chunk_bytes < 1024 || max_chunk_bytes >= 0xFFFFFF00ull) return FXG_ERR_ARG;
    *out = NULL;
    if (cudaSetDevice(device) != cudaSuccess) return FXG_ERR_CUDA;
    fxg_text *t = (fxg_text *)calloc(1, sizeof(fxg_text));
    if (!t) return FXG_ERR_NOMEM;
    t->ctx = ctx; t->device = device; t->cap_bytes = max_chunk_bytes;
    const size_t nthr = (max_chunk_bytes + 63) / 64;
    bool ok = cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&t->d_text, max_chunk_bytes + 64) == cudaSuccess && cudaMalloc(&t->d_out, max_chunk_bytes + max_chunk_bytes / 4 + 64) == cudaSuccess &&
              cudaMalloc(&t->d_cnt, (nthr + 1) * 4) == cudaSuccess && cudaMalloc(&t->d_scan, (nthr + 1) * 4) == cudaSuccess &&
              cudaMalloc(&t->d_scalars, 8 * SC_WORDS) == cudaSuccess && cudaMallocHost(&t->h_scalars, 8 * (SC_WORDS + 8)) == cudaSuccess;
    if (ok) {
        size_t need = 0, best = 0;
        cub::DeviceScan::ExclusiveSum(NULL, need, t->d_cnt, t->d_scan, (int)nthr + 1, t->st); best = need;
        cub::DeviceScan::ExclusiveSum(NULL, need, (uint64_t *)NULL, (uint64_t *)NULL, (int)(max_chunk_bytes / 8 + 1), t->st); if (need > best) best = need;
        t->tmp_bytes = best + 256;
        ok = cudaMalloc(&t->d_tmp, t->tmp_bytes) == cudaSuccess;
    }
    if (!ok) { cudaGetLastError(); fxg_text_free(t); return FXG_ERR_NOMEM; }
    *out = t;
    return FXG_OK;
}

static int ensure(fxg_text *t, void **p, size_t *cap, size_t need_elems, size_t elem)
{
    if (*cap >= need_elems) return FXG_OK;
    if (*p) cudaFree(*p);
    *p = NULL; *cap = 0;
    const size_t n = need_elems + need_elems / 8 + 1024;
    CKT(t, cudaMalloc(p, n * elem));
    *cap = n;
    return FXG_OK;
}

// op: 0 = trim (a0 = threshold, a1 = min_len), 1 = filter (a0 = min_quality, a1 = min_percent), 2 = reverse complement,
//     3 = quality-stats accumulation into hist_dev (no text output)
//     4 = fastx_clipper on a chunk whose reads all have the same length (clip != NULL; a0 = -k flag, a1 = the running
//         maximum read length seen by the caller so far, 0 = none yet)
//     5 = fastx_collapser: the reads (bases only; FASTQ qualities are validated as the reader would) are added to `col`
static int text_run(fxg_text *t, int op, const char *text_host, size_t bytes, int q_offset, int a0, int a1,
                    char *out_host, uint64_t *hist_dev, int32_t max_cycles, const fxg_clip_opts *clip, fxg_collapser *col, int64_t first_base,
                    fxg_text_report *rep)
{
    if (!t || !text_host || !rep || bytes > t->cap_bytes || op < 0 || op > 5 || ((op <= 2 || op == 4) && !out_host) || (op == 3 && !hist_dev) ||
        (op == 4 && !clip) || (op == 5 && !col) || (t->fasta && op <= 1)) return FXG_ERR_ARG;
    memset(rep, 0, sizeof(*rep));
    rep->anomaly_record = -1;
    if (bytes == 0) return FXG_OK;
    CKT(t, cudaSetDevice(t->device));
    cudaStream_t st = t->st;
    const int lpr = t->fasta ? 2 : 4;
    const uint64_t nthr = (bytes + 63) / 64;
    CKT(t, cudaMemcpyAsync(t->d_text, text_host, bytes, cudaMemcpyHostToDevice, st));
    k_nl_count<<<tgrid(nthr), 256, 0, st>>>(t->d_text, bytes, t->d_cnt, nthr);
    size_t need = t->tmp_bytes;
    CKT(t, cub::DeviceScan::ExclusiveSum(t->d_tmp, need, t->d_cnt, t->d_scan, (int)nthr + 1, st));
    uint32_t n_lines = 0;
    CKT(t, cudaMemcpyAsync(&n_lines, t->d_scan + nthr, 4, cudaMemcpyDeviceToHost, st));
    CKT(t, cudaStreamSynchronize(st));
    t->launches += 3;
    const uint32_t n_rec = n_lines / (uint32_t)lpr;
    rep->n_records = n_rec;
    if (n_rec == 0) return FXG_OK;
    int rc;
    { size_t cl = t->cap_lines; rc = ensure(t, (void **)&t->d_line_end, &cl, (size_t)n_lines + 4, 4); t->cap_lines = cl; if (rc) return rc; }
    if (t->cap_recs < n_rec) {
        size_t c;
        c = 0; if ((rc = ensure(t, (void **)&t->d_start, &c, (size_t)n_rec * 4, 4))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_llen, &c, (size_t)n_rec * 4, 4))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_seq_len, &c, n_rec, 4))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_out_len, &c, n_rec, 4))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_weight, &c, n_rec, 4))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_keep, &c, n_rec, 1))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_sizes, &c, (size_t)n_rec + 1, 8))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_offs, &c, (size_t)n_rec + 1, 8))) return rc;
        t->cap_recs = n_rec;
    }
    k_nl_scatter<<<tgrid(nthr), 256, 0, st>>>(t->d_text, bytes, t->d_scan, t->d_line_end, nthr);
    for (int k = 0; k < SC_WORDS; k++) t->h_scalars[k] = 0;
    t->h_scalars[SC_ANOM] = ~0ull; t->h_scalars[SC_MINLEN] = 0x7FFFFFFFull; t->h_scalars[SC_FIRSTNUM] = ~0ull;
    CKT(t, cudaMemcpyAsync(t->d_scalars, t->h_scalars, 8 * SC_WORDS, cudaMemcpyHostToDevice, st));
    RecTable rt = { t->d_start, t->d_llen, lpr };
    if (lpr == 4) k_recs<4><<<tgrid(n_rec), 256, 0, st>>>(t->d_text, t->d_line_end, n_rec, rt, t->d_seq_len, t->d_weight, t->d_scalars);
    else k_recs<2><<<tgrid(n_rec), 256, 0, st>>>(t->d_text, t->d_line_end, n_rec, rt, t->d_seq_len, t->d_weight, t->d_scalars);
    CKT(t, cudaMemcpyAsync(t->h_scalars, t->d_scalars, 8 * SC_WORDS, cudaMemcpyDeviceToHost, st));
    // bytes consumed = end of the last complete record
    uint32_t last_end = 0;
    CKT(t, cudaMemcpyAsync(&last_end, t->d_line_end + (size_t)lpr * n_rec - 1, 4, cudaMemcpyDeviceToHost, st));
    CKT(t, cudaStreamSynchronize(st));
    t->launches += 2;
    rep->consumed_bytes = (int64_t)last_end + 1;
    rep->n_reads = t->fasta ? (int64_t)t->h_scalars[SC_WIN] : (int64_t)n_rec;
    bool numeric = false;
    if (t->h_scalars[SC_NNUM] != 0) {
        // quality lines of another length than the sequence: numeric qualities when EVERY record of the chunk has them
        // (one encoding per chunk), otherwise the first such record is either broken or the chunk mixes the two forms
        if (t->h_scalars[SC_NNUM] == (unsigned long long)n_rec) numeric = true;
        else {
            const unsigned long long an = (t->h_scalars[SC_FIRSTNUM] << 8) | (unsigned long long)AN_QUAL_LEN;
            if (an < t->h_scalars[SC_ANOM]) t->h_scalars[SC_ANOM] = an;
        }
    }
    if (t->h_scalars[SC_ANOM] != ~0ull) {
        rep->anomaly = (int32_t)(t->h_scalars[SC_ANOM] & 0xFF);
        rep->anomaly_record = (int64_t)(t->h_scalars[SC_ANOM] >> 8);
        return FXG_OK;
    }
    const int max_len = (int)*(int *)(t->h_scalars + SC_MAXLEN);
    rep->max_len = max_len;
    rep->min_len = (int)*(int *)(t->h_scalars + SC_MINLEN);
    if (op == 4 && (rep->min_len != max_len || (a1 > 0 && max_len != a1))) {
        // mixed read lengths: the reference's aligner then reads stale bytes of earlier reads —
        // only the host packer reproduces that; hand the chunk back
        rep->anomaly = AN_LINE_COUNT + 1;     /* FXG_TEXT_MIXED_LEN */
        rep->anomaly_record = 0;
        return FXG_OK;
    }
    const int stride = (max_len + 15) & ~15;
    { size_t need_slab = (size_t)n_rec * stride; if (t->cap_slab < need_slab) {
        size_t c = 0; if ((rc = ensure(t, (void **)&t->d_seq, &c, need_slab, 1))) return rc;
        c = 0; if ((rc = ensure(t, (void **)&t->d_qual, &c, need_slab, 1))) return rc;
        t->cap_slab = need_slab; } }
    const bool has_qual = lpr == 4;
    k_pack<<<tgrid((uint64_t)n_rec * (stride >> 4)), 256, 0, st>>>(t->d_text, rt, n_rec, stride, t->d_seq, (has_qual && !numeric) ? t->d_qual : NULL);
    t->launches++;
    if (numeric) {
        k_numq<<<tgrid(n_rec, 128), 128, 0, st>>>(t->d_text, rt, n_rec, stride, t->d_qual, t->d_scalars);
        t->launches++;
        t->n_numeric_chunks++;
    }
    if (t->fasta) t->n_fasta_chunks++;
    const int q_eff = numeric ? 33 : q_offset;            // numeric values are stored as value + 33
    // the op kernel of fxg_kernels.cu on the packed slabs (validation fused; first bad record -> context counters)
    fxg_batch b = { t->d_seq, has_qual ? t->d_qual : NULL, t->d_seq_len, 0, stride, (int64_t)n_rec };
    if ((rc = fxg_report_reset(t->ctx))) { snprintf(t->err, sizeof(t->err), "%s", fxg_last_error(t->ctx)); return rc; }
    const uint8_t *alt_seq = NULL, *alt_qual = NULL;
    if (op <= 1) {
        rc = fxg_internal_scan_on_stream(t->ctx, op, &b, q_eff, a0, op == 0 ? a1 : 0, op == 1 ? a1 : 0,
                                         op == 0 ? (void *)t->d_out_len : (void *)t->d_keep, (void *)st);
    } else if (op == 2) {
        const size_t need_slab = (size_t)n_rec * stride;
        if (t->cap_oslab < need_slab) {
            size_t c = 0; if ((rc = ensure(t, (void **)&t->d_oseq, &c, need_slab, 1))) return rc;
            c = 0; if ((rc = ensure(t, (void **)&t->d_oqual, &c, need_slab, 1))) return rc;
            t->cap_oslab = need_slab;
        }
        rc = fxg_internal_revcomp_on_stream(t->ctx, &b, q_eff, t->d_oseq, has_qual ? t->d_oqual : NULL, (void *)st);
        alt_seq = t->d_oseq; alt_qual = has_qual ? t->d_oqual : NULL;
    } else if (op == 4) {
        // K-CLIP writes its lengths into the sizes scratch (int32 view) and classes into d_keep; emit lengths -> d_out_len
        int32_t *clip_len = reinterpret_cast<int32_t *>(t->d_sizes);
        b.len = NULL; b.uniform_len = max_len;      // equal lengths: the narrow (uniform) K-CLIP instantiation
        rc = fxg_internal_clip_on_stream(t->ctx, &b, q_eff, clip, clip_len, t->d_keep, (void *)st);
        if (!rc) {
            k_clip_emit_len<<<tgrid(n_rec), 256, 0, st>>>(clip_len, t->d_keep, t->d_seq_len, n_rec, a0, t->d_out_len);
            t->launches++;
            if (t->fasta) { k_clip_class_weights<<<tgrid(n_rec), 256, 0, st>>>(t->d_keep, t->d_weight, n_rec, t->d_scalars); t->launches++; }
        }
    } else if (op == 3) {
        rc = fxg_internal_stats_on_stream(t->ctx, &b, q_eff, hist_dev, max_cycles, t->fasta ? t->d_weight : NULL, (void *)st);
    } else {
        // collapser: FASTQ input is validated as the reader validates it (qualities included); the bases are checked by the table
        if (has_qual) rc = fxg_internal_scan_on_stream(t->ctx, 1, &b, q_eff, -100, 0, 100, (void *)t->d_keep, (void *)st);
    }
    if (rc) { snprintf(t->err, sizeof(t->err), "%s", fxg_last_error(t->ctx)); return rc; }
    if (op == 3 || op == 5) {
        // NB (quality stats): an illegal record makes the whole chunk an anomaly, but its earlier reads are already in hist_dev;
        // the caller must treat that as fatal (the reference prints nothing when quality_stats dies).
        CKT(t, cudaMemcpyAsync(t->h_scalars, t->d_scalars, 8 * SC_WORDS, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaMemcpyAsync(t->h_scalars + SC_WORDS, (unsigned long long *)fxg_internal_counters(t->ctx), 16, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaStreamSynchronize(st));
        t->launches += 1;
        if (t->h_scalars[SC_ANOM] != ~0ull) { rep->anomaly = (int32_t)(t->h_scalars[SC_ANOM] & 0xFF); rep->anomaly_record = (int64_t)(t->h_scalars[SC_ANOM] >> 8); return FXG_OK; }
        if (t->h_scalars[SC_WORDS + 1] != ~0ull) { rep->anomaly = AN_BAD_RECORD; rep->anomaly_record = (int64_t)t->h_scalars[SC_WORDS + 1]; return FXG_OK; }
        if (op == 5) {
            fxg_batch kb = { t->d_seq, NULL, t->d_seq_len, 0, stride, (int64_t)n_rec };
            int64_t bad = -1;
            rc = fxg_collapse_add_checked(col, &kb, t->fasta ? t->d_weight : NULL, first_base, &bad);
            if (rc) { snprintf(t->err, sizeof(t->err), "%s", fxg_collapse_error(col)); return rc; }
            if (bad >= 0) { rep->anomaly = AN_BAD_RECORD; rep->anomaly_record = bad; }
        }
        return FXG_OK;
    }
    const int32_t *ol = (op == 0 || op == 4) ? t->d_out_len : (op == 2 ? t->d_seq_len : NULL);   // revcomp keeps every read at full length
    const uint8_t *kf = op == 1 ? t->d_keep : NULL;
    if (t->decide_len_host) {
        // the caller keeps the input text and writes the output itself: only the decisions (and the line table) travel back
        int32_t *dec = reinterpret_cast<int32_t *>(t->d_sizes);
        k_decisions<<<tgrid(n_rec), 256, 0, st>>>(ol, kf, t->d_seq_len, n_rec, dec);
        k_count_kept<<<tgrid(n_rec), 256, 0, st>>>(ol, kf, NULL, n_rec, t->d_scalars);
        CKT(t, cudaMemcpyAsync(t->decide_len_host, dec, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, st));
        if (t->decide_start_host) CKT(t, cudaMemcpyAsync(t->decide_start_host, t->d_start, (size_t)n_rec * lpr * 4, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaMemcpyAsync(t->h_scalars, t->d_scalars, 8 * SC_WORDS, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaMemcpyAsync(t->h_scalars + SC_WORDS, (unsigned long long *)fxg_internal_counters(t->ctx), 64, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaStreamSynchronize(st));
        t->launches += 2;
        if (t->h_scalars[SC_ANOM] != ~0ull) { rep->anomaly = (int32_t)(t->h_scalars[SC_ANOM] & 0xFF); rep->anomaly_record = (int64_t)(t->h_scalars[SC_ANOM] >> 8); return FXG_OK; }
        if (t->h_scalars[SC_WORDS + 1] != ~0ull) { rep->anomaly = AN_BAD_RECORD; rep->anomaly_record = (int64_t)t->h_scalars[SC_WORDS + 1]; return FXG_OK; }
        rep->n_out_records = (int64_t)t->h_scalars[SC_KEPT];
        rep->n_out_reads = rep->n_out_records;
        return FXG_OK;
    }
    const uint8_t *numq = numeric ? (alt_qual ? alt_qual : t->d_qual) : NULL;
    k_emit_sizes<<<tgrid(n_rec), 256, 0, st>>>(rt, n_rec, ol, kf, numq, stride, t->d_sizes);
    CKT(t, cudaMemsetAsync(t->d_sizes + n_rec, 0, 8, st));
    need = t->tmp_bytes;
    CKT(t, cub::DeviceScan::ExclusiveSum(t->d_tmp, need, t->d_sizes, t->d_offs, (int)n_rec + 1, st));
    k_emit<<<tgrid((uint64_t)n_rec * 32), 256, 0, st>>>(t->d_text, rt, n_rec, ol, kf, t->d_offs, t->d_out, alt_seq, alt_qual, stride, numq, stride);
    k_count_kept<<<tgrid(n_rec), 256, 0, st>>>(ol, kf, t->fasta ? t->d_weight : NULL, n_rec, t->d_scalars);
    uint64_t out_bytes = 0;
    CKT(t, cudaMemcpyAsync(&out_bytes, t->d_offs + n_rec, 8, cudaMemcpyDeviceToHost, st));
    CKT(t, cudaMemcpyAsync(t->h_scalars, t->d_scalars, 8 * SC_WORDS, cudaMemcpyDeviceToHost, st));
    CKT(t, cudaMemcpyAsync(t->h_scalars + SC_WORDS, (unsigned long long *)fxg_internal_counters(t->ctx), 64, cudaMemcpyDeviceToHost, st));
    CKT(t, cudaStreamSynchronize(st));
    t->launches += 5;
    const unsigned long long *cnt = t->h_scalars + SC_WORDS;       // the context's counters: CNT_OUT, CNT_FIRST_BAD, CNT_AUX0..
    if (op == 4) for (int k = 0; k < 6; k++)       // CNT_OUT, CNT_AUX0+k; FASTA: weighted by the identifiers' read counts
        rep->clip_class[k] = t->fasta ? (int64_t)t->h_scalars[SC_CLASS + k] : (int64_t)cnt[k == 0 ? 0 : 2 + k];
    if (t->h_scalars[SC_ANOM] != ~0ull) {      // K-NUMQ met a malformed number
        rep->anomaly = (int32_t)(t->h_scalars[SC_ANOM] & 0xFF);
        rep->anomaly_record = (int64_t)(t->h_scalars[SC_ANOM] >> 8);
        return FXG_OK;
    }
    if (cnt[1] != ~0ull) {          // the op kernel found an illegal base / quality: host path decides
        rep->anomaly = AN_BAD_RECORD;
        rep->anomaly_record = (int64_t)cnt[1];
        return FXG_OK;
    }
    rep->n_out_records = (int64_t)t->h_scalars[SC_KEPT];
    rep->n_out_reads = t->fasta ? (int64_t)t->h_scalars[SC_WKEPT] : rep->n_out_records;
    rep->out_bytes = (int64_t)out_bytes;
    rep->raw_out_bytes = (int64_t)out_bytes;
    if (out_bytes && t->deflate) {
        size_t zbytes = 0; uint32_t crc = 0;
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, t->device);
        CKT(t, deflate_run(t->d_out, (size_t)out_bytes, t->d_dfl, t->d_dfl_scratch, t->dfl_scratch_bytes, (uint64_t *)(t->h_scalars + SC_WORDS),
                           t->h_dfl_crc, &zbytes, &crc, sm, st, &t->launches));
        rep->out_bytes = (int64_t)zbytes;
        rep->out_crc32_pure = crc;
        rep->deflated = 1;
        CKT(t, cudaMemcpyAsync(out_host, t->d_dfl, zbytes, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaStreamSynchronize(st));
    } else if (out_bytes) {
        CKT(t, cudaMemcpyAsync(out_host, t->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
        CKT(t, cudaStreamSynchronize(st));
    }
    return FXG_OK;
}

extern "C" int fxg_text_run_host(fxg_text *t, int op, const char *text_host, size_t bytes, int q_offset, int a0, int a1,
                                 char *out_host, fxg_text_report *rep)
{
    if (op < 0 || op > 2) return FXG_ERR_ARG;
    return text_run(t, op, text_host, bytes, q_offset, a0, a1, out_host, NULL, 0, NULL, NULL, 0, rep);
}

// op 0 / 1 with the per-record decision as the only result: out_len_host[r] = surviving length or -1; line_start_host (may be
// NULL) = byte offset of each of the record's 4 lines in the chunk.  For callers that keep the input text and write the output
// from it (device-to-host traffic: 4 to 20 bytes per record instead of the whole text).
extern "C" int fxg_text_decide_host(fxg_text *t, int op, const char *text_host, size_t bytes, int q_offset, int a0, int a1,
                                    int32_t *out_len_host, uint32_t *line_start_host, fxg_text_report *rep)
{
    if (!t || op < 0 || op > 1 || !out_len_host || t->fasta) return FXG_ERR_ARG;
    t->decide_len_host = out_len_host; t->decide_start_host = line_start_host;
    char dummy = 0;
    const int rc = text_run(t, op, text_host, bytes, q_offset, a0, a1, &dummy, NULL, 0, NULL, NULL, 0, rep);
    t->decide_len_host = NULL; t->decide_start_host = NULL;
    return rc;
}

extern "C" int fxg_text_clip_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, const fxg_clip_opts *o,
                                  int show_adapter_only, int expect_len, char *out_host, fxg_text_report *rep)
{
    return text_run(t, 4, text_host, bytes, q_offset, show_adapter_only, expect_len, out_host, NULL, 0, o, NULL, 0, rep);
}

extern "C" int fxg_text_stats_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, uint64_t *hist_dev,
                                   int32_t max_cycles, fxg_text_report *rep)
{
    return text_run(t, 3, text_host, bytes, q_offset, 0, 0, NULL, hist_dev, max_cycles, NULL, NULL, 0, rep);
}

extern "C" int fxg_text_collapse_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, fxg_collapser *col, int64_t first_base,
                                      fxg_text_report *rep)
{
    return text_run(t, 5, text_host, bytes, q_offset, 0, 0, NULL, NULL, 0, NULL, col, first_base, rep);
}
