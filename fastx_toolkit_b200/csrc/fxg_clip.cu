// fxg_clip.cu — K-CLIP-ALIGN: fastx_clipper's half-local adapter alignment + cut-off rules + discard
// cascade, one thread per read.
//
//   HalfLocalSequenceAlignment::reset_matrix / populate_matrix   src/libfastx/sequence_alignment.cpp:340-428
//   find_optimal_alignment_from_point (backtrace)                src/libfastx/sequence_alignment.cpp:496-604
//   adapter_cutoff_index                                         src/fastx_clipper/fastx_clipper.cpp:159-241
//   discard cascade                                              src/fastx_clipper/fastx_clipper.cpp:280-319
//
// The reference fills a W x H score matrix and an origin matrix, then walks back from the best cell.
// Here the walk is replaced by a payload carried FORWARD with every cell — {matches, mismatches, neutral,
// tstart | gaps, qstart} packed in two 32-bit words so that no field straddles a word and every update
// is one 32-bit add — so only the previous column lives in registers and nothing is stored.  The score
// arithmetic is the reference's: fp32 adds of {+1, -1, 0.1f, 0, -5} in the same order, strict '>' in the
// same candidate order (diagonal, up, left), first maximum in (x outer, y inner) scan order wins.
// ALU-bound (13*L dependent cell updates per ~L bytes): reported as cell updates/s, not against HBM.
#include <stdlib.h>

#include "fxg_kernels.cuh"
#include "fxg_clip_dpx.cuh"

namespace fxg {

namespace {
constexpr uint32_t INC_M = 1u, INC_X = 1u << 7, INC_N = 1u << 14;   // lo word: matches, mismatches, neutral, tstart<<21
constexpr int TSTART_SHIFT = 21, QSTART_SHIFT = 15;                 // hi word: gaps, qstart<<15
constexpr float GAP = -5.0f;
}

struct Cell { float s; uint32_t lo, hi; };

__device__ __forceinline__ float target_border(int y) { return y <= 3 ? 0.0f : GAP * (float)(y - 3); }

// One DP column x (query character qc) over adapter rows 0..H-1.  prev* hold column x-1 on entry
// (FIRST: the virtual border column) and column x on exit.
template <int HMAX, bool FIRST>
__device__ __forceinline__ void clip_column(const ClipParams &P, int H, int x, uint32_t qc,
                                            float (&ps)[HMAX], uint32_t (&plo)[HMAX], uint32_t (&phi)[HMAX],
                                            float &best, int &bx, int &by, uint32_t &blo, uint32_t &bhi)
{
    const bool qn = qc == (uint32_t)'N';
    const uint32_t fresh_hi = (uint32_t)x << QSTART_SHIFT;
    Cell up;                // cell (x, y-1) of the current column
    up.s = 0.0f; up.lo = 0; up.hi = 0;
    float diag_s = 0.0f;    // score (x-1, y-1); for y == 0: query_border[x-1] = 0, or target_border[-1] (= 0.0) when x == 0
    uint32_t diag_lo = 0, diag_hi = 0;
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        if (y >= H) break;
        const uint32_t tc = P.adapter[y];
        const bool tn = tc == (uint32_t)'N';
        const bool eq = qc == tc;
        // nucleotide_match_score (sequence_alignment.h:157-169) and match_value (:125-131)
        const float ms = (qn && tn) ? 0.0f : ((qn || tn) ? 0.1f : (eq ? 1.0f : -1.0f));
        const uint32_t inc = (qn || tn) ? INC_N : (eq ? INC_M : INC_X);
        const uint32_t fresh_lo = (uint32_t)y << TSTART_SHIFT;

        // left neighbour (x-1, y): the border column when FIRST
        float left_s = (FIRST ? target_border(y) : ps[y]) + GAP;
        if (y > 3 && y - 3 > x) left_s = -100000.0f;                  // sequence_alignment.cpp:388-390
        const uint32_t left_lo = FIRST ? fresh_lo : plo[y];
        const uint32_t left_hi = (FIRST ? fresh_hi : phi[y]) + 1u;    // one more gap

        // upper neighbour (x, y-1): query_border[x] = 0 when y == 0
        const float up_s = (y == 0 ? 0.0f : up.s) + GAP;
        const uint32_t up_lo = (y == 0) ? fresh_lo : up.lo;
        const uint32_t up_hi = ((y == 0) ? fresh_hi : up.hi) + 1u;

        // diagonal (x-1, y-1)
        const float d_in = (y == 0) ? 0.0f : (FIRST ? target_border(y - 1) : diag_s);
        const float ul_s = d_in + ms;
        const bool d_out = (y == 0) || FIRST;                         // predecessor outside the matrix
        const uint32_t ul_lo = (d_out ? fresh_lo : diag_lo) + inc;
        const uint32_t ul_hi = d_out ? fresh_hi : diag_hi;

        // remember (x-1, y) as the next row's diagonal before overwriting it
        if (!FIRST) { diag_s = ps[y]; diag_lo = plo[y]; diag_hi = phi[y]; }

        Cell c;
        c.s = ul_s; c.lo = ul_lo; c.hi = ul_hi;                       // FROM_UPPER_LEFT first,
        if (up_s > c.s) { c.s = up_s; c.lo = up_lo; c.hi = up_hi; }   // then FROM_UPPER,
        if (left_s > c.s) { c.s = left_s; c.lo = left_lo; c.hi = left_hi; }   // then FROM_LEFT: strict '>'
        ps[y] = c.s; plo[y] = c.lo; phi[y] = c.hi;
        up = c;
        if (c.s > best) { best = c.s; bx = x; by = y; blo = c.lo; bhi = c.hi; }
    }
}

// adapter_cutoff_index (fastx_clipper.cpp:159-241); size_t arithmetic of the reference kept (query_size-2
// wraps for 1-base reads).
__device__ __forceinline__ int cutoff_index(uint32_t lo, uint32_t hi, int qend, int qsize, int min_adapter_len)
{
    const unsigned long long matches = lo & 127u, mism = (lo >> 7) & 127u, neutral = (lo >> 14) & 127u;
    const unsigned long long tstart = (lo >> TSTART_SHIFT) & 127u, gaps = hi & 0x7FFFu;
    const int qstart = (int)((hi >> QSTART_SHIFT) & 0x7FFFu);
    const int asz = (int)(neutral + matches + mism + gaps);
    const unsigned long long qe = (unsigned long long)qend, qs = (unsigned long long)qsize;
    if (asz == 0) return -1;
    if (min_adapter_len > 0 && asz < min_adapter_len) return -1;
    if (qe == qs - 1ull && mism == 0) return qstart;
    if (asz > 5 && tstart == 0 && (matches * 100ull / (unsigned long long)asz) >= 75ull) return qstart;
    if (asz > 11 && (matches * 100ull / (unsigned long long)asz) >= 80ull) return qstart;
    if (qe >= qs - 2ull && asz <= 5 && matches >= 3) return qstart;
    return -1;
}

enum { CLS_WRITE = 0, CLS_ADAPTER_ONLY = 1, CLS_TOO_SHORT = 2, CLS_NON_CLIPPED = 3, CLS_CLIPPED = 4, CLS_HAS_N = 5 };

// cut-off + discard cascade (fastx_clipper.cpp:280-319) of one read; writes its outputs, returns its class
__device__ __forceinline__ int clip_epilogue(const ClipParams &P, int64_t g, int L, bool bad, uint32_t lo, uint32_t hi, int bx, int firstN)
{
    const int cut = bad ? -1 : cutoff_index(lo, hi, bx, L, P.min_adapter_len);
    int newL = L, cls;
    if (cut > 0) { const int at = cut + P.keep_delta; if (at < newL) newL = at; }
    if (cut == 0) cls = CLS_ADAPTER_ONLY;
    else if ((unsigned)newL < (unsigned)P.min_length) cls = CLS_TOO_SHORT;
    else if (cut == -1 && P.discard_non_clipped) cls = CLS_NON_CLIPPED;
    else if (cut > 0 && P.discard_clipped) cls = CLS_CLIPPED;
    else if (P.discard_unknown && firstN < newL) cls = CLS_HAS_N;
    else cls = CLS_WRITE;
    P.out_len[g] = (cls == CLS_WRITE) ? newL : -1;
    if (P.out_class) P.out_class[g] = (uint8_t)cls;
    if (P.out_cut) P.out_cut[g] = cut;
    return cls;
}
// reader's quality range check for one FASTQ read (fastx.c:118-135); non-zero when a byte is illegal
__device__ __forceinline__ uint32_t clip_qual_bad(const ClipParams &P, int64_t g, int L)
{
    const uint8_t *qrow = P.qual + (size_t)g * P.stride;
    uint32_t badbits = 0;
    for (int c = 0; c * 16 < L; c++) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(qrow + c * 16));
        const uint32_t qw[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int wd = 0; wd < 4; wd++)
            badbits |= qual_bad_bits(qw[wd], qw[wd] | HI, P.qk) & HI & head_mask(L - 16 * c - 4 * wd);
    }
    return badbits;
}

template <int HMAX>
__global__ void __launch_bounds__(128) k_clip(const __grid_constant__ ClipParams P)
{
    const int H = P.alen;
    const int lane = threadIdx.x & 31;
    const int64_t N = P.n_dev ? *P.n_dev : P.n;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; base < N; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = base + lane;
        const bool active = g < N;
        int cls = -1;
        if (active) {
            int L = P.len ? __ldg(P.len + g) : P.uniform_len;
            int W = P.width ? __ldg(P.width + g) : L;
            bool bad = (L <= 0 || L > P.stride || W < L || W > P.stride);
            if (bad) { L = 0; W = 0; }
            const uint8_t *row = P.seq + (size_t)g * P.stride;

            float ps[HMAX]; uint32_t plo[HMAX], phi[HMAX];
            float best = -1000000.0f;
            int bx = 0, by = 0, firstN = 0x7FFFFFFF;
            uint32_t blo = 0, bhi = 0, badbits = 0;
            for (int x0 = 0; x0 < W; x0 += 4) {
                uint32_t wq = __ldg(reinterpret_cast<const uint32_t *>(row + x0));
                const int nb = (W - x0 < 4) ? (W - x0) : 4;
                for (int k = 0; k < nb; k++) {
                    const uint32_t qc = wq & 0xFFu;
                    wq >>= 8;
                    const int x = x0 + k;
                    if (x < L) {
                        const bool ok = qc == 'A' || qc == 'C' || qc == 'G' || qc == 'T' || qc == 'N';
                        if (!ok) badbits = 1;
                        if (qc == 'N' && x < firstN) firstN = x;
                    }
                    if (x == 0) clip_column<HMAX, true>(P, H, x, qc, ps, plo, phi, best, bx, by, blo, bhi);
                    else clip_column<HMAX, false>(P, H, x, qc, ps, plo, phi, best, bx, by, blo, bhi);
                }
            }
            if (P.qual && !bad) {           // FASTQ input: the reader validates qualities too (fastx.c:118-135)
                const uint8_t *qrow = P.qual + (size_t)g * P.stride;
                for (int c = 0; c * 16 < L; c++) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(qrow + c * 16));
                    const uint32_t qw[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                    for (int wd = 0; wd < 4; wd++)
                        badbits |= qual_bad_bits(qw[wd], qw[wd] | HI, P.qk) & HI & head_mask(L - 16 * c - 4 * wd);
                }
            }
            if (badbits || bad) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));

            // cut-off + cascade (fastx_clipper.cpp:280-319)
            const int cut = bad ? -1 : cutoff_index(blo, bhi, bx, L, P.min_adapter_len);
            int newL = L;
            if (cut > 0) { const int at = cut + P.keep_delta; if (at < newL) newL = at; }
            if (cut == 0) cls = CLS_ADAPTER_ONLY;
            else if ((unsigned)newL < (unsigned)P.min_length) cls = CLS_TOO_SHORT;
            else if (cut == -1 && P.discard_non_clipped) cls = CLS_NON_CLIPPED;
            else if (cut > 0 && P.discard_clipped) cls = CLS_CLIPPED;
            else if (P.discard_unknown && firstN < newL) cls = CLS_HAS_N;
            else cls = CLS_WRITE;
            P.out_len[g] = (cls == CLS_WRITE) ? newL : -1;
            if (P.out_class) P.out_class[g] = (uint8_t)cls;
            if (P.out_cut) P.out_cut[g] = cut;
        }
        // class counters for the -v report, one atomic per class per warp
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (lane == 0 && m) atomicAdd(&P.counters[c == 0 ? CNT_OUT : CNT_AUX0 + c], (unsigned long long)__popc(m));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast variant: score column in registers + a bit-packed ORIGIN matrix (2 bits per cell, one word per
// query column, in per-thread local memory) + the reference's own backtrace from the best cell
// (sequence_alignment.cpp:496-604).  ~3x fewer instructions per cell than carrying the payload forward,
// because a cell costs three fp32 adds, two compares and their selects, nothing else.
// Used when the matrix width fits MAXW columns and the adapter fits one word per column.
// ------------------------------------------------------------------------------------------------
// HMAX is the adapter length rounded up to a multiple of 4; ALL HMAX rows are evaluated (rows >= H depend on
// rows < H but never feed back, and are excluded from the arg-max), so the unrolled column has no early exits
// and no predicated state merging.
template <int HMAX, typename WordT, bool FIRST>
__device__ __forceinline__ WordT clip_column_bits(const ClipParams &P, int H, int x, uint32_t qc, float (&ps)[HMAX],
                                                  float &best, int &bx, int &by)
{
    const bool qn = qc == (uint32_t)'N';
    WordT word = 0;
    float up_s = 0.0f;      // score (x, y-1)
    float diag_s = 0.0f;    // score (x-1, y-1)
#pragma unroll
    for (int y = 0; y < HMAX; y++) {
        const uint32_t tc = P.adapter[y];
        const bool tn = tc == (uint32_t)'N';
        const float ms = (qn && tn) ? 0.0f : ((qn || tn) ? 0.1f : ((qc == tc) ? 1.0f : -1.0f));
        float left = (FIRST ? target_border(y) : ps[y]) + GAP;
        if (y > 3 && y - 3 > x) left = -100000.0f;
        const float up = (y == 0 ? 0.0f : up_s) + GAP;
        const float ul = ((y == 0) ? 0.0f : (FIRST ? target_border(y - 1) : diag_s)) + ms;
        if (!FIRST) diag_s = ps[y];
        float sc = ul;
        WordT o = 3;                                   // FROM_UPPER_LEFT
        if (up > sc) { sc = up; o = 1; }               // FROM_UPPER
        if (left > sc) { sc = left; o = 2; }           // FROM_LEFT
        word |= o << (2 * y);
        ps[y] = sc;
        up_s = sc;
        const bool live = (y < HMAX - 3) || (y < H);   // only the last 3 rows can lie beyond the adapter
        if (live && sc > best) { best = sc; bx = x; by = y; }
    }
    return word;
}

template <int HMAX, int MAXW, typename WordT>
__global__ void __launch_bounds__(128) k_clip_bits(const __grid_constant__ ClipParams P)
{
    const int H = P.alen;
    const int lane = threadIdx.x & 31;
    // list mode (second pass of the integer fast path): only the reads whose indices k_clip_dpx put on the list
    const int64_t nn = P.list ? (int64_t)*P.list_count : (P.n_dev ? *P.n_dev : P.n);
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; base < nn; base += (int64_t)gridDim.x * blockDim.x) {
        const bool active = base + lane < nn;
        const int64_t g = !active ? 0 : (P.list ? (int64_t)P.list[base + lane] : base + lane);
        int cls = -1;
        if (active) {
            int L = P.len ? __ldg(P.len + g) : P.uniform_len;
            int W = P.width ? __ldg(P.width + g) : L;
            const bool bad = (L <= 0 || L > P.stride || W < L || W > P.stride || W > MAXW);
            if (bad) { L = 0; W = 0; }
            const uint8_t *row = P.seq + (size_t)g * P.stride;

            float ps[HMAX];
            WordT origin[MAXW];
            float best = -1000000.0f;
            int bx = 0, by = 0, firstN = 0x7FFFFFFF;
            uint32_t badbits = 0;
            for (int x0 = 0; x0 < W; x0 += 4) {
                uint32_t wq = __ldg(reinterpret_cast<const uint32_t *>(row + x0));
                const int nb = (W - x0 < 4) ? (W - x0) : 4;
                for (int k = 0; k < nb; k++) {
                    const uint32_t qc = wq & 0xFFu;
                    wq >>= 8;
                    const int x = x0 + k;
                    if (x < L) {
                        const bool ok = qc == 'A' || qc == 'C' || qc == 'G' || qc == 'T' || qc == 'N';
                        if (!ok) badbits = 1;
                        if (qc == 'N' && x < firstN) firstN = x;
                    }
                    origin[x] = (x == 0) ? clip_column_bits<HMAX, WordT, true>(P, H, x, qc, ps, best, bx, by)
                                         : clip_column_bits<HMAX, WordT, false>(P, H, x, qc, ps, best, bx, by);
                }
            }
            // backtrace from the best cell: find_optimal_alignment_from_point, sequence_alignment.cpp:496-604
            uint32_t lo = 0, hi = 0;
            if (!bad) {
                int qi = bx, ti = by, matches = 0, mism = 0, neutral = 0, gaps = 0, qstart = bx, tstart = by;
                while (qi >= 0 && ti >= 0) {
                    qstart = qi; tstart = ti;
                    const uint32_t o = (uint32_t)(origin[qi] >> (2 * ti)) & 3u;
                    if (o == 2u) { gaps++; qi--; }
                    else if (o == 3u) {
                        const uint32_t q = __ldg(row + qi), t = P.adapter[ti];
                        if (q == 'N' || t == 'N') neutral++; else if (q == t) matches++; else mism++;
                        qi--; ti--;
                    } else { gaps++; ti--; }
                }
                lo = (uint32_t)matches | ((uint32_t)mism << 7) | ((uint32_t)neutral << 14) | ((uint32_t)tstart << TSTART_SHIFT);
                hi = (uint32_t)gaps | ((uint32_t)qstart << QSTART_SHIFT);
            }
            if (P.qual && !bad) badbits |= clip_qual_bad(P, g, L);
            if (badbits || bad) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g));
            cls = clip_epilogue(P, g, L, bad, lo, hi, bx, firstN);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (lane == 0 && m) atomicAdd(&P.counters[c == 0 ? CNT_OUT : CNT_AUX0 + c], (unsigned long long)__popc(m));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Integer fast path (fxg_clip_dpx.cuh): two reads per thread in packed s16x2 DPX arithmetic, for batches of
// uniform length and adapters without 'N'.  Reads with an 'N' or an illegal character are appended to
// P.list and redone by k_clip_bits (fp32, exact tie behaviour of the 0.1f score) in a second launch.
// ------------------------------------------------------------------------------------------------
template <int HMAX>
__global__ void __launch_bounds__(128) k_clip_dpx(const __grid_constant__ ClipParams P)
{
    const int H = P.alen, L = P.uniform_len;
    const int lane = threadIdx.x & 31;
    const int64_t N = P.n_dev ? *P.n_dev : P.n;
    const int64_t npairs = (N + 1) >> 1;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; base < npairs; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pair = base + lane;
        int cls0 = -1, cls1 = -1;
        if (pair < npairs) {
            const int64_t g0 = 2 * pair;
            const bool has1 = g0 + 1 < N;
            const int64_t g1 = has1 ? g0 + 1 : g0;
            const uint8_t *row0 = P.seq + (size_t)g0 * P.stride, *row1 = P.seq + (size_t)g1 * P.stride;
            dpx::PairOut o;
            dpx::align_pair<HMAX, 256>(row0, row1, L, P.adapter, H, o);
            if (P.qual) {
                if (clip_qual_bad(P, g0, L)) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g0));
                if (has1 && clip_qual_bad(P, g1, L)) atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + g1));
            }
            if (o.exact & 1u) P.list[atomicAdd(P.list_count, 1ull)] = (int32_t)g0;
            else cls0 = clip_epilogue(P, g0, L, false, o.lo[0], o.hi[0], o.bx[0], 0x7FFFFFFF);
            if (has1) {
                if (o.exact & 2u) P.list[atomicAdd(P.list_count, 1ull)] = (int32_t)g1;
                else cls1 = clip_epilogue(P, g1, L, false, o.lo[1], o.hi[1], o.bx[1], 0x7FFFFFFF);
            }
        }
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const unsigned m0 = __ballot_sync(0xffffffffu, cls0 == c), m1 = __ballot_sync(0xffffffffu, cls1 == c);
            if (lane == 0 && (m0 | m1)) atomicAdd(&P.counters[c == 0 ? CNT_OUT : CNT_AUX0 + c], (unsigned long long)(__popc(m0) + __popc(m1)));
        }
    }
}

cudaError_t launch_clip(const ClipParams &p, int sm_count, int max_width, cudaStream_t st)
{
    int64_t blocks = (p.n + 127) / 128;
    if (blocks < 1) blocks = 1;
    const int64_t cap = (int64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    const unsigned b = (unsigned)blocks;
    if (p.list) {
        // first pass of the integer fast path (the caller launches the list pass afterwards with list_pass = 1)
        if (!p.list_pass) {
            int64_t pb = ((p.n + 1) / 2 + 127) / 128;
            if (pb > cap) pb = cap;
            if (pb < 1) pb = 1;
            const unsigned pbu = (unsigned)pb;
            switch ((p.alen + 3) / 4) {
            case 1: k_clip_dpx<4><<<pbu, 128, 0, st>>>(p); break;
            case 2: k_clip_dpx<8><<<pbu, 128, 0, st>>>(p); break;
            case 3: k_clip_dpx<12><<<pbu, 128, 0, st>>>(p); break;
            default: k_clip_dpx<16><<<pbu, 128, 0, st>>>(p); break;
            }
            return cudaGetLastError();
        }
    }
    const char *force = getenv("FXG_CLIP_PAYLOAD");   // experimentation: force the forward-payload kernels
    if ((p.list_pass || !(force && force[0] == '1')) && p.alen <= 32 && max_width <= 1024) {
        const int hb = (p.alen + 3) / 4;      // adapter length bucket (multiple of 4 rows)
#define FXG_CLIPB(HM, WT)                                                                          \
        if (max_width <= 256) k_clip_bits<HM, 256, WT><<<b, 128, 0, st>>>(p);                      \
        else k_clip_bits<HM, 1024, WT><<<b, 128, 0, st>>>(p);                                      \
        return cudaGetLastError();
        switch (hb) {
        case 1: { FXG_CLIPB(4, uint32_t) }
        case 2: { FXG_CLIPB(8, uint32_t) }
        case 3: { FXG_CLIPB(12, uint32_t) }
        case 4: { FXG_CLIPB(16, uint32_t) }
        case 5: { FXG_CLIPB(20, unsigned long long) }
        case 6: { FXG_CLIPB(24, unsigned long long) }
        case 7: { FXG_CLIPB(28, unsigned long long) }
        default: { FXG_CLIPB(32, unsigned long long) }
        }
#undef FXG_CLIPB
    }
    if (p.alen <= 16) k_clip<16><<<b, 128, 0, st>>>(p);
    else if (p.alen <= 32) k_clip<32><<<b, 128, 0, st>>>(p);
    else k_clip<100><<<b, 128, 0, st>>>(p);   // column spills to local memory: slow but exact
    return cudaGetLastError();
}

}  // namespace fxg
