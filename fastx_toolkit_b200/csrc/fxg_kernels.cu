// fxg_kernels.cu — hand-written sm_100a kernels for the FASTX per-read transform loop.
//
// Common shape ("tile pipeline"): a persistent grid of CTAs; each CTA walks tiles of `tile_reads`
// consecutive reads.  One elected thread streams the tile's seq/qual rows (contiguous in the SoA
// slab) into a shared-memory ring with TMA 1-D bulk copies (cp.async.bulk, completion on an
// mbarrier); G lanes cooperate on one read, each lane pulling 16-byte chunks with LDS.128
// (G is chosen from the stride so that the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank
// groups) and running 4-bytes-per-register SWAR code.  No tensor cores: this is byte/integer work
// bounded by HBM bandwidth and, second, by the INT32 ALU pipe.
//
//   K-TRIM    fastq_quality_trimmer body   src/fastq_quality_trimmer/fastq_quality_trimmer.c:91-103
//   K-FILTER  fastq_quality_filter body    src/fastq_quality_filter/fastq_quality_filter.c:78-129,141-161
//   K-REVCOMP fastx_reverse_complement     src/fastx_reverse_complement/fastx_reverse_complement.c:43-104
//   (every kernel fuses K-VALIDATE: src/libfastx/fastx.c:45-54,118-135,361-362)
#include "fxg_kernels.cuh"

namespace fxg {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void note_bad(unsigned long long *counters, int64_t gidx)
{
    atomicMin(&counters[CNT_FIRST_BAD], (unsigned long long)gidx);
}

template <int G> __device__ __forceinline__ int group_max(int v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int G> __device__ __forceinline__ uint32_t group_sum(uint32_t v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G> __device__ __forceinline__ uint32_t group_or(uint32_t v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// K-TRIM / K-FILTER
// ------------------------------------------------------------------------------------------------
struct ScanAcc {
    uint32_t badq;     // bit7 flags: illegal quality byte seen
    uint32_t bads;     // any bit: illegal base seen
    int      lastc;    // trim: last 16-byte chunk holding a byte with q >= t
    uint32_t lowb;     // filter: per-byte-lane counters of bytes with q < min_q
    uint32_t low;      // filter: flushed total
    uint32_t hasn;     // has-N: bit7 flags of 'N' bytes
    uint32_t ba, bc, bg, bt;   // artifacts: per-byte-lane counters of A / C / G / T
    uint32_t ca, cc, cg, ct;   // artifacts: flushed totals
};

// per-word work of the sequence-only modes (the (f-2)/(f-4) loop bodies on the same tile ring):
//   MODE_HASN      strchr(nucleotides, 'N')                      src/fastq_to_fasta/fastq_to_fasta.c:79-82
//   MODE_ARTIFACT  per-read counts of A, C, G, T                 src/fastx_artifacts_filter/fastx_artifacts_filter.c:56-114
template <int MODE>
__device__ __forceinline__ void scan_seq_word(ScanAcc &a, uint32_t x, uint32_t m)
{
    if (MODE == MODE_HASN) {                      // zero byte of (x ^ "NNNN"), exact for any byte value
        const uint32_t z = x ^ 0x4E4E4E4Eu;
        a.hasn |= ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & HI & m;
    }
    if (MODE == MODE_ARTIFACT) {
        // nucleotide index per byte by table (A 0, C 1, G 2, T 3, N 4; the selector is the one the validation computes):
        // bit 0 = C or T, bit 1 = G or T, bit 2 = N.  C, G, T and N are counted per byte lane, A is what is left of the length.
        const uint32_t nuc4 = __byte_perm(0x01800080u, 0x02048003u, base_selector(x)) & m;
        const uint32_t b0 = nuc4 & ONES, b1 = (nuc4 >> 1) & ONES, t = b0 & b1;
        a.bt += t;
        a.bc += b0 ^ t;
        a.bg += b1 ^ t;
        a.ba += (nuc4 >> 2) & ONES;               // N
    }
}
__device__ __forceinline__ void scan_flush_counts(ScanAcc &a)
{
    a.ca += __dp4a(a.ba, ONES, 0u); a.cc += __dp4a(a.bc, ONES, 0u); a.cg += __dp4a(a.bg, ONES, 0u); a.ct += __dp4a(a.bt, ONES, 0u);
    a.ba = a.bc = a.bg = a.bt = 0;
}

template <int MODE, bool HAS_SEQ>
__device__ __forceinline__ void scan_chunk(ScanAcc &a, const uint4 &q, const uint4 &s, const QualK &k, int c)
{
    const uint32_t qw[4] = { q.x, q.y, q.z, q.w };
    uint32_t anyge = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t x = qw[w], xh = x | HI;
        a.badq |= qual_bad_bits(x, xh, k);
        if (MODE == MODE_TRIM || MODE == MODE_FILTER) {
            const uint32_t ge = qual_ge_bits(xh, k);
            if (MODE == MODE_TRIM) anyge |= ge;
            else a.lowb += (~(ge >> 7)) & ONES;
        }
    }
    if (MODE == MODE_TRIM) { if (anyge & HI) a.lastc = max(a.lastc, c); }
    if (HAS_SEQ) {
        a.bads |= seq_bad_bits(s.x) | seq_bad_bits(s.y);
        a.bads |= seq_bad_bits(s.z) | seq_bad_bits(s.w);
        if (MODE == MODE_HASN || MODE == MODE_ARTIFACT) {
            scan_seq_word<MODE>(a, s.x, 0xFFFFFFFFu); scan_seq_word<MODE>(a, s.y, 0xFFFFFFFFu);
            scan_seq_word<MODE>(a, s.z, 0xFFFFFFFFu); scan_seq_word<MODE>(a, s.w, 0xFFFFFFFFu);
        }
    }
}

// the partial chunk at the end of a read: bytes >= rem are padding and must not count
template <int MODE, bool HAS_SEQ>
__device__ __forceinline__ void scan_tail(ScanAcc &a, const uint4 &q, const uint4 &s, const QualK &k, int c, int rem)
{
    const uint32_t qw[4] = { q.x, q.y, q.z, q.w };
    const uint32_t sw[4] = { s.x, s.y, s.z, s.w };
    uint32_t anyge = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t m = head_mask(rem - 4 * w);
        const uint32_t x = qw[w], xh = x | HI;
        a.badq |= qual_bad_bits(x, xh, k) & m;
        if (MODE == MODE_TRIM || MODE == MODE_FILTER) {
            const uint32_t ge = qual_ge_bits(xh, k);
            if (MODE == MODE_TRIM) anyge |= ge & m;
            else a.lowb += (~(ge >> 7)) & ONES & m;
        }
        if (HAS_SEQ) {
            a.bads |= seq_bad_bits(sw[w]) & m;
            if (MODE == MODE_HASN || MODE == MODE_ARTIFACT) scan_seq_word<MODE>(a, sw[w], m);
        }
    }
    if (MODE == MODE_TRIM) { if (anyge & HI) a.lastc = max(a.lastc, c); }
}

// One read (or this lane's share of it): full 16-byte chunks, then the masked partial chunk.
//   G == 1: the lane owns the whole read and visits its full chunks in an order rotated by `rot`
//           (conflict-free LDS.128 when the row pitch is an even number of 16-byte units);
//   G  > 1: lane j of the group takes chunks j, j+G, ...
template <int G, int MODE, bool HAS_SEQ>
__device__ __forceinline__ void scan_read(ScanAcc &a, const uint8_t *qrow, const uint8_t *srow, int L, int j, int rot,
                                          const QualK &qk)
{
    const int nfull = L >> 4, rem = L & 15;
    int since_flush = 0;
    if (G == 1) {
        const int o = rot < nfull ? rot : 0;
#pragma unroll 2
        for (int k = 0; k < nfull; k++) {
            int c = k + o;
            if (c >= nfull) c -= nfull;
            const uint4 q = lds128(qrow + c * 16);
            uint4 sq = make_uint4(0, 0, 0, 0);
            if (HAS_SEQ) sq = lds128(srow + c * 16);
            scan_chunk<MODE, HAS_SEQ>(a, q, sq, qk, c);
            if (MODE == MODE_FILTER) {
                if (++since_flush == 32) {   // byte lanes hold <= 128: flush before they can wrap
                    a.low += __dp4a(a.lowb, ONES, 0u);
                    a.lowb = 0; since_flush = 0;
                }
            }
            if (MODE == MODE_ARTIFACT) { if (++since_flush == 32) { scan_flush_counts(a); since_flush = 0; } }
        }
    } else {
#pragma unroll 2
        for (int c = j; c < nfull; c += G) {
            const uint4 q = lds128(qrow + c * 16);
            uint4 sq = make_uint4(0, 0, 0, 0);
            if (HAS_SEQ) sq = lds128(srow + c * 16);
            scan_chunk<MODE, HAS_SEQ>(a, q, sq, qk, c);
            if (MODE == MODE_FILTER) {
                if (++since_flush == 32) {
                    a.low += __dp4a(a.lowb, ONES, 0u);
                    a.lowb = 0; since_flush = 0;
                }
            }
            if (MODE == MODE_ARTIFACT) { if (++since_flush == 32) { scan_flush_counts(a); since_flush = 0; } }
        }
    }
    if (rem && (nfull & (G - 1)) == j) {
        const uint4 q = lds128(qrow + nfull * 16);
        uint4 sq = make_uint4(0, 0, 0, 0);
        if (HAS_SEQ) sq = lds128(srow + nfull * 16);
        scan_tail<MODE, HAS_SEQ>(a, q, sq, qk, nfull, rem);
    }
    if (MODE == MODE_FILTER) a.low += __dp4a(a.lowb, ONES, 0u);
    if (MODE == MODE_ARTIFACT) scan_flush_counts(a);
}

// trimmer decision for one read once the last chunk holding a base with q >= t is known:
// exact byte position inside that chunk -> surviving length (fastq_quality_trimmer.c:93-101)
__device__ __forceinline__ int trim_newlen(const uint8_t *qrow, int lastc, int L, const QualK &qk)
{
    if (lastc < 0) return 0;
    const uint4 q = lds128(qrow + lastc * 16);
    const uint32_t qw[4] = { q.x, q.y, q.z, q.w };
    const int valid = (lastc == (L >> 4)) ? (L & 15) : 16;
    int pos = -1;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t f = qual_ge_bits(qw[w] | HI, qk) & HI & head_mask(valid - 4 * w);
        if (f) pos = 4 * w + ((31 - __clz(f)) >> 3);
    }
    return lastc * 16 + pos + 1;
}

// ---- CTA-tile variant: any stride up to the smem limit (long reads); one tile ring per CTA ----------
template <int G, int MODE, bool HAS_SEQ>
__global__ void __launch_bounds__(THREADS) k_scan(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int64_t N = P.n_dev ? *P.n_dev : P.n;       // reads in the batch (device-side count in fused pipelines)
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ unsigned int s_kept;

    const int tid = threadIdx.x;
    const int S = P.stride;
    const int TR = P.tile_reads;
    const int stages = P.stages;
    const uint32_t slab_bytes = (uint32_t)TR * (uint32_t)S;
    const uint32_t stage_bytes = slab_bytes * (HAS_SEQ ? 2u : 1u);
    const int64_t ntiles = (N + TR - 1) / TR;
    const QualK qk = P.qk;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
        s_kept = 0;
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int s) {
        const int64_t r0 = tile * TR;
        const int64_t left = N - r0;
        const uint32_t bytes = (uint32_t)(left < TR ? left : TR) * (uint32_t)S;
        uint8_t *dst = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes * (HAS_SEQ ? 2u : 1u));
        bulk_g2s(dst, P.qual + r0 * S, bytes, &full_bar[s]);
        if (HAS_SEQ) bulk_g2s(dst + slab_bytes, P.seq + r0 * S, bytes, &full_bar[s]);
    };

    if (tid == 0) {
        for (int i = 0; i < stages; i++) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
            if (t < ntiles) issue(t, i);
        }
    }

    const int j = tid & (G - 1);
    const int rsub = tid / G;
    constexpr int RPP = THREADS / G;   // reads per pass
    unsigned kept_local = 0;
    int s = 0;
    uint32_t parity = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&full_bar[s], parity);
        const uint8_t *stage = smem + (size_t)s * stage_bytes;
        const int64_t r0 = tile * TR;
        const int64_t left = N - r0;
        const int nr = (int)(left < TR ? left : TR);

        for (int rb = 0; rb < nr; rb += RPP) {
            const int rr = rb + rsub;
            const bool active = rr < nr;
            const int64_t g = r0 + rr;
            int L = 0;
            if (active) L = P.len ? __ldg(P.len + g) : P.uniform_len;
            const bool lenbad = active && (L <= 0 || L > S);
            if (lenbad) L = 0;
            const uint8_t *qrow = stage + (size_t)rr * S;
            const uint8_t *srow = qrow + slab_bytes;

            ScanAcc a;
            a.badq = 0; a.bads = 0; a.lastc = -1; a.lowb = 0; a.low = 0;
            a.hasn = 0; a.ba = a.bc = a.bg = a.bt = 0; a.ca = a.cc = a.cg = a.ct = 0;
            scan_read<G, MODE, HAS_SEQ>(a, qrow, srow, L, j, 0, qk);
            if (((a.badq & HI) | a.bads) != 0 || lenbad) note_bad(P.counters, P.index_base + g);

            if (MODE == MODE_TRIM) {
                const int lastc = group_max<G>(a.lastc);
                if (j == 0 && active) {
                    const int newlen = trim_newlen(qrow, lastc, L, qk);
                    const bool keep = newlen >= 1 && newlen >= P.min_len;
                    reinterpret_cast<int32_t *>(P.out)[g] = keep ? newlen : -1;
                    kept_local += keep ? 1u : 0u;
                }
            } else {
                const uint32_t low = group_sum<G>(a.low);
                if (j == 0 && active) {
                    const bool keep = !P.force_drop && !lenbad &&
                                      (100ll * (long long)low <= (long long)L * (long long)P.pct_keep);
                    reinterpret_cast<uint8_t *>(P.out)[g] = keep ? 1 : 0;
                    kept_local += keep ? 1u : 0u;
                }
            }
        }

        __syncthreads();   // every lane is done with stage s -> it may be refilled
        if (tid == 0) {
            const int64_t nt = tile + (int64_t)stages * gridDim.x;
            if (nt < ntiles) issue(nt, s);
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }

    // kept-read count: warp reduce -> smem -> one global atomic per CTA
    kept_local = __reduce_add_sync(0xffffffffu, kept_local);
    if ((tid & 31) == 0 && kept_local) atomicAdd(&s_kept, kept_local);
    __syncthreads();
    if (tid == 0 && s_kept) atomicAdd(&P.counters[CNT_OUT], (unsigned long long)s_kept);
}

// ---- warp-private variant (the fast path for short reads) ----------------------------------------
// Every warp owns a private ring of `stages` tiles of R = 32/G reads and its own mbarriers; lane 0
// issues the TMA bulk copies.  No CTA-wide barrier exists: warps drift freely, so one warp's wait for
// HBM is covered by the others' SWAR work, and with G == 1 a lane owns a whole read (no cross-lane
// reduction, and the masked-tail and exact-position steps run with all 32 lanes busy).
template <int G, int MODE, bool HAS_SEQ>
__global__ void __launch_bounds__(W_THREADS) k_scan_w(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int64_t N = P.n_dev ? *P.n_dev : P.n;       // reads in the batch (device-side count in fused pipelines)
    __shared__ __align__(8) uint64_t full_bar[W_WARPS][MAX_STAGES];

    constexpr int R = 32 / G;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int S = P.stride;
    const int stages = P.stages;
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    const uint32_t stage_bytes = slab_bytes * (HAS_SEQ ? 2u : 1u);
    uint8_t *wbase = smem + (size_t)w * stages * stage_bytes;
    uint64_t *bars = full_bar[w];
    const int64_t ntiles = (N + R - 1) / R;
    const int64_t gw = (int64_t)blockIdx.x * W_WARPS + w, GW = (int64_t)gridDim.x * W_WARPS;
    const QualK qk = P.qk;

    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto issue = [&](int64_t tile, int s) {
        const int64_t r0 = tile * R;
        const int64_t left = N - r0;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * (uint32_t)S;
        uint8_t *dst = wbase + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&bars[s], bytes * (HAS_SEQ ? 2u : 1u));
        bulk_g2s(dst, P.qual + r0 * S, bytes, &bars[s]);
        if (HAS_SEQ) bulk_g2s(dst + slab_bytes, P.seq + r0 * S, bytes, &bars[s]);
    };
    if (lane == 0) {
        for (int i = 0; i < stages; i++) {
            const int64_t t = gw + (int64_t)i * GW;
            if (t < ntiles) issue(t, i);
        }
    }

    const int j = lane & (G - 1);
    const int rr = lane / G;
    const int rot = (G == 1) ? ((lane & 7) >> P.rot_shift) : 0;
    const uint32_t row_off = (uint32_t)rr * (uint32_t)S;
    unsigned kept_local = 0;
    int s = 0;
    uint32_t parity = 0;

    for (int64_t tile = gw; tile < ntiles; tile += GW) {
        mbar_wait(&bars[s], parity);
        const uint8_t *qrow = wbase + (size_t)s * stage_bytes + row_off;
        const uint8_t *srow = qrow + slab_bytes;
        const int64_t g = tile * R + rr;
        const bool active = g < N;
        int L = 0;
        if (active) L = P.len ? __ldg(P.len + g) : P.uniform_len;
        const bool lenbad = active && (L <= 0 || L > S);
        if (lenbad) L = 0;

        ScanAcc a;
        a.badq = 0; a.bads = 0; a.lastc = -1; a.lowb = 0; a.low = 0;
        a.hasn = 0; a.ba = a.bc = a.bg = a.bt = 0; a.ca = a.cc = a.cg = a.ct = 0;
        scan_read<G, MODE, HAS_SEQ>(a, qrow, srow, L, j, rot, qk);
        if (((a.badq & HI) | a.bads) != 0 || lenbad) note_bad(P.counters, P.index_base + g);

        if (MODE == MODE_VALIDATE) {
            // the reader's checks alone: nothing per read but the first-bad-read counter
        } else if (MODE == MODE_HASN) {
            const uint32_t hn = group_or<G>(a.hasn);
            if (j == 0 && active) {
                reinterpret_cast<uint8_t *>(P.out)[g] = hn ? 1 : 0;
                kept_local += hn ? 1u : 0u;
            }
        } else if (MODE == MODE_ARTIFACT) {
            const uint32_t cn = group_sum<G>(a.ca), cc = group_sum<G>(a.cc), cg = group_sum<G>(a.cg), ct = group_sum<G>(a.ct);
            const uint32_t ca = (uint32_t)L - cn - cc - cg - ct;       // (a.ca holds the N count, see scan_seq_word)
            if (j == 0 && active) {
                const int lim = L - 3;      // max_allowed_different_bases = 3 (fastx_artifacts_filter.c:66,99-107)
                const bool artifact = (int)ca >= lim || (int)cc >= lim || (int)cg >= lim || (int)ct >= lim;
                reinterpret_cast<uint8_t *>(P.out)[g] = artifact ? 0 : 1;
                kept_local += artifact ? 0u : 1u;
            }
        } else if (MODE == MODE_TRIM) {
            const int lastc = group_max<G>(a.lastc);
            if (j == 0 && active) {
                const int newlen = trim_newlen(qrow, lastc, L, qk);
                const bool keep = newlen >= 1 && newlen >= P.min_len;
                reinterpret_cast<int32_t *>(P.out)[g] = keep ? newlen : -1;
                kept_local += keep ? 1u : 0u;
            }
        } else {
            const uint32_t low = group_sum<G>(a.low);
            if (j == 0 && active) {
                const bool keep = !P.force_drop && !lenbad &&
                                  (100ll * (long long)low <= (long long)L * (long long)P.pct_keep);
                reinterpret_cast<uint8_t *>(P.out)[g] = keep ? 1 : 0;
                kept_local += keep ? 1u : 0u;
            }
        }

        __syncwarp();      // all lanes are done reading stage s -> lane 0 may refill it
        if (lane == 0) {
            const int64_t nt = tile + (int64_t)stages * GW;
            if (nt < ntiles) issue(nt, s);
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }

    kept_local = __reduce_add_sync(0xffffffffu, kept_local);
    if (lane == 0 && kept_local) atomicAdd(&P.counters[CNT_OUT], (unsigned long long)kept_local);
}

// ------------------------------------------------------------------------------------------------
// K-REVCOMP
// ------------------------------------------------------------------------------------------------
// One 16-byte output chunk = 16 input bytes ending at e = L - 16*oc, reversed.  Those bytes lie in
// the 32-byte window made of input chunks ce-1 and ce (ce = e>>4) at byte offset re = e&15, which is
// the same for every chunk of the read, so the word offset (re>>2) is resolved with a warp-uniform
// switch and the byte offset (re&3) with one PRMT per output word (reverse + realign at once).
__device__ __forceinline__ uint4 reverse_window(const uint4 &A, const uint4 &B, int re)
{
    const uint32_t sel = 0x0123u + 0x1111u * (uint32_t)(re & 3);
    uint32_t w0, w1, w2, w3, w4;   // W[wo .. wo+4], ascending
    switch (re >> 2) {
    case 0:  w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
    case 1:  w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
    case 2:  w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
    default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
    }
    uint4 o;
    o.x = __byte_perm(w3, w4, sel);
    o.y = __byte_perm(w2, w3, sel);
    o.z = __byte_perm(w1, w2, sel);
    o.w = __byte_perm(w0, w1, sel);
    return o;
}

// Reverse-complement this lane's share (output chunks j, j+G, ...) of one read from the input rows into
// the output rows (all in shared memory); the whole stride is written so the tile can be bulk-stored.
template <int G, bool HAS_QUAL>
__device__ __forceinline__ void revcomp_read(const uint8_t *srow, const uint8_t *qrow, uint8_t *osrow, uint8_t *oqrow,
                                             int L, int nchunks, int j, const QualK &qk, uint32_t &bads, uint32_t &badq)
{
    const int re = L & 15;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int oc = j; oc < nchunks; oc += G) {
        const int e = L - 16 * oc;     // exclusive end of the input bytes for this chunk
        uint4 os = zero, oq = zero;
        if (e > 0) {
            const int ce = e >> 4;
            const uint4 A = (ce >= 1) ? lds128(srow + (ce - 1) * 16) : zero;
            const uint4 B = (re != 0) ? lds128(srow + ce * 16) : zero;
            const uint4 r = reverse_window(A, B, re);
            const int nreal = e < 16 ? e : 16;
            const uint32_t rw[4] = { r.x, r.y, r.z, r.w };
            uint32_t cw[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t b = 0;
                cw[w] = seq_complement(rw[w], b);
                const uint32_t m = head_mask(nreal - 4 * w);
                bads |= b & m;
                cw[w] &= m;
            }
            os = make_uint4(cw[0], cw[1], cw[2], cw[3]);
            if (HAS_QUAL) {
                const uint4 QA = (ce >= 1) ? lds128(qrow + (ce - 1) * 16) : zero;
                const uint4 QB = (re != 0) ? lds128(qrow + ce * 16) : zero;
                const uint4 rq = reverse_window(QA, QB, re);
                uint32_t qv[4] = { rq.x, rq.y, rq.z, rq.w };
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t m = head_mask(nreal - 4 * w);
                    badq |= qual_bad_bits(qv[w], qv[w] | HI, qk) & m;
                    qv[w] &= m;
                }
                oq = make_uint4(qv[0], qv[1], qv[2], qv[3]);
            }
        }
        *reinterpret_cast<uint4 *>(osrow + oc * 16) = os;
        if (HAS_QUAL) *reinterpret_cast<uint4 *>(oqrow + oc * 16) = oq;
    }
}

// ---- CTA-tile variant (any stride) ----------------------------------------------------------------------
template <int G, bool HAS_QUAL>
__global__ void __launch_bounds__(THREADS) k_revcomp(const __grid_constant__ RevcompParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];

    const int tid = threadIdx.x;
    const int S = P.stride;
    const int TR = P.tile_reads;
    const int stages = P.stages;
    const uint32_t slab_bytes = (uint32_t)TR * (uint32_t)S;
    const uint32_t stage_bytes = slab_bytes * (HAS_QUAL ? 2u : 1u);
    uint8_t *outbuf = smem + (size_t)stages * stage_bytes;   // 2 output buffers of stage_bytes
    const int64_t ntiles = (P.n + TR - 1) / TR;
    const QualK qk = P.qk;
    const int nchunks = S >> 4;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int s) {
        const int64_t r0 = tile * TR;
        const int64_t left = P.n - r0;
        const uint32_t bytes = (uint32_t)(left < TR ? left : TR) * (uint32_t)S;
        uint8_t *dst = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes * (HAS_QUAL ? 2u : 1u));
        bulk_g2s(dst, P.seq + r0 * S, bytes, &full_bar[s]);
        if (HAS_QUAL) bulk_g2s(dst + slab_bytes, P.qual + r0 * S, bytes, &full_bar[s]);
    };
    if (tid == 0) {
        for (int i = 0; i < stages; i++) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
            if (t < ntiles) issue(t, i);
        }
    }

    const int j = tid & (G - 1);
    const int rsub = tid / G;
    constexpr int RPP = THREADS / G;
    int s = 0, ob = 0;
    uint32_t parity = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&full_bar[s], parity);
        const uint8_t *stage = smem + (size_t)s * stage_bytes;
        uint8_t *obuf = outbuf + (size_t)ob * stage_bytes;
        const int64_t r0 = tile * TR;
        const int64_t left = P.n - r0;
        const int nr = (int)(left < TR ? left : TR);

        for (int rb = 0; rb < nr; rb += RPP) {
            const int rr = rb + rsub;
            if (rr >= nr) continue;
            const int64_t g = r0 + rr;
            int L = P.len ? __ldg(P.len + g) : P.uniform_len;
            const bool lenbad = (L <= 0 || L > S);
            if (lenbad) L = 0;
            const uint8_t *srow = stage + (size_t)rr * S;
            uint8_t *osrow = obuf + (size_t)rr * S;
            uint32_t bads = 0, badq = 0;
            revcomp_read<G, HAS_QUAL>(srow, srow + slab_bytes, osrow, osrow + slab_bytes, L, nchunks, j, qk, bads, badq);
            if ((bads | (badq & HI)) != 0 || lenbad) note_bad(P.counters, P.index_base + g);
        }

        fence_async_smem();                 // my smem writes -> visible to the bulk-store engine
        if (tid == 0) bulk_wait_read<0>();  // previous tile's store has drained the other out buffer
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)nr * (uint32_t)S;
            bulk_s2g(P.out_seq + r0 * S, obuf, bytes);
            if (HAS_QUAL) bulk_s2g(P.out_qual + r0 * S, obuf + slab_bytes, bytes);
            bulk_commit();
            const int64_t nt = tile + (int64_t)stages * gridDim.x;
            if (nt < ntiles) issue(nt, s);
        }
        ob ^= 1;
        if (++s == stages) { s = 0; parity ^= 1u; }
    }
    if (tid == 0) bulk_wait_all<0>();       // all stores complete before the CTA (and its smem) retires
}

// ---- warp-private variant: each warp owns one input tile and one output tile of R = 32/G reads --------
// load (TMA) -> wait -> reverse/complement smem->smem -> bulk store (TMA) -> next load; the store of tile i
// drains while the warp waits for tile i+1, and the SM's other warps cover both waits.
template <int G, bool HAS_QUAL>
__global__ void __launch_bounds__(W_THREADS) k_revcomp_w(const __grid_constant__ RevcompParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[W_WARPS];

    constexpr int R = 32 / G;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int S = P.stride;
    const uint32_t slab_bytes = (uint32_t)R * (uint32_t)S;
    const uint32_t stage_bytes = slab_bytes * (HAS_QUAL ? 2u : 1u);
    uint8_t *ibuf = smem + (size_t)w * 2 * stage_bytes;
    uint8_t *obuf = ibuf + stage_bytes;
    uint64_t *bar = &full_bar[w];
    const int64_t ntiles = (P.n + R - 1) / R;
    const int64_t gw = (int64_t)blockIdx.x * W_WARPS + w, GW = (int64_t)gridDim.x * W_WARPS;
    const QualK qk = P.qk;
    const int nchunks = S >> 4;

    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();

    auto issue = [&](int64_t tile) {
        const int64_t r0 = tile * R;
        const int64_t left = P.n - r0;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * (uint32_t)S;
        mbar_arrive_expect_tx(bar, bytes * (HAS_QUAL ? 2u : 1u));
        bulk_g2s(ibuf, P.seq + r0 * S, bytes, bar);
        if (HAS_QUAL) bulk_g2s(ibuf + slab_bytes, P.qual + r0 * S, bytes, bar);
    };
    if (lane == 0 && gw < ntiles) issue(gw);

    const int j = lane & (G - 1);
    const int rr = lane / G;
    uint32_t parity = 0;
    for (int64_t tile = gw; tile < ntiles; tile += GW) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        const int64_t r0 = tile * R;
        const int64_t g = r0 + rr;
        const int64_t left = P.n - r0;
        const int nr = (int)(left < R ? left : R);
        if (lane == 0) bulk_wait_read<0>();     // the previous tile's store has finished reading obuf
        __syncwarp();
        if (rr < nr) {
            int L = P.len ? __ldg(P.len + g) : P.uniform_len;
            const bool lenbad = (L <= 0 || L > S);
            if (lenbad) L = 0;
            const uint8_t *srow = ibuf + (size_t)rr * S;
            uint8_t *osrow = obuf + (size_t)rr * S;
            uint32_t bads = 0, badq = 0;
            revcomp_read<G, HAS_QUAL>(srow, srow + slab_bytes, osrow, osrow + slab_bytes, L, nchunks, j, qk, bads, badq);
            if ((bads | (badq & HI)) != 0 || lenbad) note_bad(P.counters, P.index_base + g);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)nr * (uint32_t)S;
            bulk_s2g(P.out_seq + r0 * S, obuf, bytes);
            if (HAS_QUAL) bulk_s2g(P.out_qual + r0 * S, obuf + slab_bytes, bytes);
            bulk_commit();
            const int64_t nt = tile + GW;
            if (nt < ntiles) issue(nt);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------------------------------------
// synthetic workload generator (include/fxg_synth.h) — one thread per 16-byte chunk
// ------------------------------------------------------------------------------------------------
}  // namespace fxg
#include "fxg_synth.h"
namespace fxg {

__global__ void __launch_bounds__(256) k_synth(const SynthParams P)
{
    const int chunks = P.stride >> 4;
    const int64_t total = P.n * chunks;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        const uint64_t r = fxg_synth_read_key(P.seed, (uint64_t)(P.first_read + i), P.kind, (uint64_t)P.n_total);
        uint32_t sw[4] = { 0, 0, 0, 0 }, qw[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int pos = c * 16 + b;
            if (pos < P.len) {
                sw[b >> 2] |= (uint32_t)fxg_synth_base(r, pos, P.len, P.kind) << (8 * (b & 3));
                qw[b >> 2] |= (uint32_t)(fxg_synth_phred(r, pos, P.len) + P.q_offset) << (8 * (b & 3));
            }
        }
        const size_t off = (size_t)i * P.stride + (size_t)c * 16;
        if (P.seq) *reinterpret_cast<uint4 *>(P.seq + off) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
        if (P.qual) *reinterpret_cast<uint4 *>(P.qual + off) = make_uint4(qw[0], qw[1], qw[2], qw[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// Kernels that may need more than 48 KB of dynamic shared memory opt in right before the launch (a per-launch
// cudaFuncSetAttribute costs microseconds; doing it for every instantiation at start-up made fxg_init() load
// ~100 kernels eagerly).
#define FXG_LAUNCH_DYN(KERNEL, GRID, BLOCK, SMEM, STREAM, PARAMS)                                   \
    do {                                                                                           \
        auto kf_ = KERNEL;                                                                         \
        if ((SMEM) > 40u * 1024u   /* static mbarrier words count against the 48 KB default too */) cudaFuncSetAttribute(kf_, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM); \
        kf_<<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(PARAMS);                                        \
    } while (0)

template <int MODE, bool HAS_SEQ>
static cudaError_t launch_scan_g(const TilePlan &plan, const ScanParams &p, cudaStream_t st)
{
#define FXG_SCAN_CASE(GV)                                                                          \
    case GV:                                                                                       \
        FXG_LAUNCH_DYN((k_scan<GV, MODE, HAS_SEQ>), plan.grid, THREADS, plan.smem_bytes, st, p);       \
        break;
    switch (plan.g) {
        FXG_SCAN_CASE(1) FXG_SCAN_CASE(2) FXG_SCAN_CASE(4) FXG_SCAN_CASE(8) FXG_SCAN_CASE(16) FXG_SCAN_CASE(32)
    default: return cudaErrorInvalidValue;
    }
#undef FXG_SCAN_CASE
    return cudaGetLastError();
}

template <int MODE, bool HAS_SEQ>
static cudaError_t launch_scan_w(const TilePlan &plan, const ScanParams &p, cudaStream_t st)
{
#define FXG_SCANW_CASE(GV)                                                                         \
    case GV:                                                                                       \
        FXG_LAUNCH_DYN((k_scan_w<GV, MODE, HAS_SEQ>), plan.grid, W_THREADS, plan.smem_bytes, st, p);   \
        break;
    switch (plan.g) {
        FXG_SCANW_CASE(1) FXG_SCANW_CASE(2) FXG_SCANW_CASE(4) FXG_SCANW_CASE(8)
    default: return cudaErrorInvalidValue;
    }
#undef FXG_SCANW_CASE
    return cudaGetLastError();
}

cudaError_t launch_scan(int mode, bool has_seq, const TilePlan &plan, const ScanParams &p, cudaStream_t st)
{
    if (mode == MODE_VALIDATE || mode == MODE_HASN || mode == MODE_ARTIFACT) {      // sequence modes: warp-private ring only
        if (!plan.warp_ring || !has_seq) return cudaErrorInvalidValue;
        if (mode == MODE_VALIDATE) return launch_scan_w<MODE_VALIDATE, true>(plan, p, st);
        if (mode == MODE_HASN) return launch_scan_w<MODE_HASN, true>(plan, p, st);
        return launch_scan_w<MODE_ARTIFACT, true>(plan, p, st);
    }
    if (plan.warp_ring) {
        if (mode == MODE_TRIM) return has_seq ? launch_scan_w<MODE_TRIM, true>(plan, p, st) : launch_scan_w<MODE_TRIM, false>(plan, p, st);
        return has_seq ? launch_scan_w<MODE_FILTER, true>(plan, p, st) : launch_scan_w<MODE_FILTER, false>(plan, p, st);
    }
    if (mode == MODE_TRIM) return has_seq ? launch_scan_g<MODE_TRIM, true>(plan, p, st) : launch_scan_g<MODE_TRIM, false>(plan, p, st);
    return has_seq ? launch_scan_g<MODE_FILTER, true>(plan, p, st) : launch_scan_g<MODE_FILTER, false>(plan, p, st);
}

template <bool HAS_QUAL>
static cudaError_t launch_revcomp_g(const TilePlan &plan, const RevcompParams &p, cudaStream_t st)
{
#define FXG_RC_CASE(GV)                                                                            \
    case GV:                                                                                       \
        FXG_LAUNCH_DYN((k_revcomp<GV, HAS_QUAL>), plan.grid, THREADS, plan.smem_bytes, st, p);         \
        break;
    switch (plan.g) {
        FXG_RC_CASE(1) FXG_RC_CASE(2) FXG_RC_CASE(4) FXG_RC_CASE(8) FXG_RC_CASE(16) FXG_RC_CASE(32)
    default: return cudaErrorInvalidValue;
    }
#undef FXG_RC_CASE
    return cudaGetLastError();
}

template <bool HAS_QUAL>
static cudaError_t launch_revcomp_w(const TilePlan &plan, const RevcompParams &p, cudaStream_t st)
{
#define FXG_RCW_CASE(GV)                                                                           \
    case GV:                                                                                       \
        FXG_LAUNCH_DYN((k_revcomp_w<GV, HAS_QUAL>), plan.grid, W_THREADS, plan.smem_bytes, st, p);     \
        break;
    switch (plan.g) {
        FXG_RCW_CASE(1) FXG_RCW_CASE(2) FXG_RCW_CASE(4) FXG_RCW_CASE(8)
    default: return cudaErrorInvalidValue;
    }
#undef FXG_RCW_CASE
    return cudaGetLastError();
}

cudaError_t launch_revcomp(bool has_qual, const TilePlan &plan, const RevcompParams &p, cudaStream_t st)
{
    if (plan.warp_ring) return has_qual ? launch_revcomp_w<true>(plan, p, st) : launch_revcomp_w<false>(plan, p, st);
    return has_qual ? launch_revcomp_g<true>(plan, p, st) : launch_revcomp_g<false>(plan, p, st);
}

cudaError_t launch_synth(const SynthParams &p, cudaStream_t st)
{
    const int64_t total = p.n * (p.stride >> 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    k_synth<<<(unsigned)blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fxg
