/*
 * fastx_oracle.c — CPU restatement of the FASTX-Toolkit 0.0.14 hot path (plain C).
 *
 * *** TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See fastx_oracle.h. ***
 *
 * Written from the behaviour of the reference (file:line cited per function, paths relative to
 * /root/reference); no reference source is copied.  Parity status: pinned against the reference's
 * Galaxy fixtures and the oracle/_ref binaries (tests/test_oracle_*.py), except collapser tie
 * order and stats "-N", which no reference fixture pins (binary-only pinning).
 */
#include "fastx_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ a1: validation ---------- */

/* src/libfastx/fastx.c:45-84 — all six hot-path tools open the reader with ALLOW_N,
 * REQUIRE_UPPERCASE, so the allowed set is exactly {A,C,G,T,N}. */
int fxo_seq_first_invalid(const uint8_t *seq, int len)
{
    for (int i = 0; i < len; i++) {
        uint8_t c = seq[i];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N')) return i;
    }
    return -1;
}

/* src/libfastx/fastx.c:118-135 — q = (char)c - offset must lie in [-15, 93]; `char` is signed
 * on x86-64, so bytes >= 128 are negative. */
int fxo_qual_first_invalid(const uint8_t *qual, int len, int q_offset)
{
    for (int i = 0; i < len; i++) {
        int q = (int)(signed char)qual[i] - q_offset;
        if (q < FXO_MIN_Q || q > FXO_MAX_Q) return i;
    }
    return -1;
}

/* ------------------------------------------------------------------ a2: trimmer -------------- */

/* src/fastq_quality_trimmer/fastq_quality_trimmer.c:93-102 — scan backwards while q < t; keep
 * iff a base survived and the surviving prefix is at least min_len long. */
int fxo_trim_record(const uint8_t *qual, int len, int q_offset, int threshold, int min_len)
{
    int i;
    for (i = len - 1; i >= 0; i--) {
        int q = (int)(signed char)qual[i] - q_offset;
        if (q < threshold) continue;
        break;
    }
    if (i >= 0 && i + 1 >= min_len) return i + 1;
    return -1;
}

/* ------------------------------------------------------------------ a3: filter --------------- */

/* src/fastq_quality_filter/fastq_quality_filter.c:78-108 — walk a counting-sort histogram to the
 * n-th smallest element.  array_size is QUALITY_VALUES_RANGE = 108 in the reference although 109
 * values are legal; we keep slack so the reference's one-past-the-end reads stay defined. */
static int nth_element_index(const int *array, int array_size, int n)
{
    int pos = 0;
    while (pos < array_size && array[pos] == 0) pos++;
    if (pos == array_size) return -1; /* reference: errx "bug: got empty array" */
    while (n > 0) {
        if (array[pos] > n) break;
        n -= array[pos];
        pos++;
        while (array[pos] == 0 && pos < array_size) pos++;
    }
    return pos;
}

/* src/fastq_quality_filter/fastq_quality_filter.c:110-129 + decision at :150-156 */
int fxo_filter_record(const uint8_t *qual, int len, int q_offset, int min_quality, int min_percent)
{
    int bins[FXO_QBINS + 3];
    memset(bins, 0, sizeof bins);
    int count = 0;
    for (int i = 0; i < len; i++) {
        count++;
        bins[((int)(signed char)qual[i] - q_offset) - FXO_MIN_Q]++;
    }
    int idx = nth_element_index(bins, FXO_MAX_Q - FXO_MIN_Q /* 108 */, count * (100 - min_percent) / 100);
    int value = idx + FXO_MIN_Q;
    return value >= min_quality;
}

/* ------------------------------------------------------------------ a4: reverse complement --- */

/* src/fastx_reverse_complement/fastx_reverse_complement.c:43-72 (uppercase arms only: lowercase
 * cannot pass the reader) and :74-104 (complement in place, then swap ends; qualities swapped too). */
static uint8_t complement_base(uint8_t b)
{
    switch (b) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    default:  return b; /* 'N' */
    }
}

void fxo_revcomp_record(const uint8_t *seq, const uint8_t *qual, int len, uint8_t *oseq, uint8_t *oqual)
{
    for (int i = 0; i < len; i++) oseq[i] = complement_base(seq[len - 1 - i]);
    if (qual && oqual)
        for (int i = 0; i < len; i++) oqual[i] = qual[len - 1 - i];
}

/* ------------------------------------------------------------------ batch wrappers ----------- */

static int rec_len(const int32_t *len, int uniform_len, int64_t i) { return len ? len[i] : uniform_len; }

static int record_is_bad(const uint8_t *s, const uint8_t *q, int L, int q_offset)
{
    if (L <= 0) return 1; /* src/libfastx/fastx.c:361-362: empty sequence is fatal */
    if (s && fxo_seq_first_invalid(s, L) >= 0) return 1;
    if (q && fxo_qual_first_invalid(q, L, q_offset) >= 0) return 1;
    return 0;
}

void fxo_trim_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                    int stride, int64_t n, int q_offset, int threshold, int min_len,
                    int32_t *out_len, int64_t *first_bad)
{
    int64_t bad = -1;
    for (int64_t i = 0; i < n; i++) {
        int L = rec_len(len, uniform_len, i);
        const uint8_t *q = qual + i * (int64_t)stride;
        if (bad < 0 && record_is_bad(seq ? seq + i * (int64_t)stride : NULL, q, L, q_offset)) bad = i;
        out_len[i] = fxo_trim_record(q, L, q_offset, threshold, min_len);
    }
    if (first_bad) *first_bad = bad;
}

void fxo_filter_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                      int stride, int64_t n, int q_offset, int min_quality, int min_percent,
                      uint8_t *keep, int64_t *first_bad)
{
    int64_t bad = -1;
    for (int64_t i = 0; i < n; i++) {
        int L = rec_len(len, uniform_len, i);
        const uint8_t *q = qual + i * (int64_t)stride;
        if (record_is_bad(seq ? seq + i * (int64_t)stride : NULL, q, L, q_offset)) {
            if (bad < 0) bad = i;
            keep[i] = 0;
            continue;
        }
        keep[i] = (uint8_t)fxo_filter_record(q, L, q_offset, min_quality, min_percent);
    }
    if (first_bad) *first_bad = bad;
}

void fxo_revcomp_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                       int stride, int64_t n, uint8_t *oseq, uint8_t *oqual)
{
    for (int64_t i = 0; i < n; i++) {
        int L = rec_len(len, uniform_len, i);
        int64_t o = i * (int64_t)stride;
        fxo_revcomp_record(seq + o, qual ? qual + o : NULL, L, oseq + o, oqual ? oqual + o : NULL);
    }
}

/* ------------------------------------------------------------------ a5: quality stats -------- */

/* src/fastx_quality_stats/fastx_quality_stats.c:115-133: one nucleotide_data per (cycle, nuc);
 * nuc index 0 = ALL, 1..5 = A,C,G,T,N (:100-108). */
typedef struct {
    int min, max, count;
    unsigned long long sum;
    int bins[FXO_QBINS + 3];
} nuc_data;

struct fxo_stats {
    int max_cycles;
    nuc_data *d; /* [max_cycles][6] */
};

static int nuc_index(uint8_t c)
{ /* :142-152 — lookup table, 0 (= ALL) for anything else */
    switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    case 'N': case 'n': return 5;
    default: return 0;
    }
}

fxo_stats *fxo_stats_new(int max_cycles)
{
    fxo_stats *s = (fxo_stats *)calloc(1, sizeof *s);
    s->max_cycles = max_cycles;
    s->d = (nuc_data *)calloc((size_t)max_cycles * 6, sizeof(nuc_data));
    for (int i = 0; i < max_cycles * 6; i++) { s->d[i].min = 100; s->d[i].max = -100; } /* :157-162 */
    return s;
}

void fxo_stats_free(fxo_stats *s) { if (s) { free(s->d); free(s); } }

/* :166-216 — per base: two count bumps, and for FASTQ two (min,max,sum,bin) updates.
 * NB the sum is NOT weighted by reads_count although counts are. */
void fxo_stats_add(fxo_stats *s, const uint8_t *seq, const uint8_t *qual, int len, int q_offset, int weight)
{
    for (int c = 0; c < len && c < s->max_cycles; c++) {
        nuc_data *all = &s->d[(size_t)c * 6];
        nuc_data *nd = &s->d[(size_t)c * 6 + nuc_index(seq[c])];
        all->count += weight;
        nd->count += weight;
        if (qual) {
            int q = (int)(signed char)qual[c] - q_offset;
            if (q < all->min) all->min = q;
            if (q > all->max) all->max = q;
            all->sum += (unsigned long long)(long long)q;
            all->bins[q - FXO_MIN_Q] += weight;
            if (q < nd->min) nd->min = q;
            if (q > nd->max) nd->max = q;
            nd->sum += (unsigned long long)(long long)q;
            nd->bins[q - FXO_MIN_Q] += weight;
        }
    }
}

void fxo_stats_add_batch(fxo_stats *s, const uint8_t *seq, const uint8_t *qual, const int32_t *len,
                         int uniform_len, int stride, int64_t n, int q_offset)
{
    for (int64_t i = 0; i < n; i++) {
        int64_t o = i * (int64_t)stride;
        fxo_stats_add(s, seq + o, qual ? qual + o : NULL, rec_len(len, uniform_len, i), q_offset, 1);
    }
}

int fxo_stats_export_hist(const fxo_stats *s, uint64_t *hist, int max_cycles)
{
    int cycles = 0;
    for (int c = 0; c < s->max_cycles && c < max_cycles; c++) {
        if (s->d[(size_t)c * 6].count == 0) break;
        for (int nuc = 0; nuc < 5; nuc++)
            for (int b = 0; b < FXO_QBINS; b++)
                hist[((size_t)c * 5 + nuc) * FXO_QBINS + b] = (uint64_t)s->d[(size_t)c * 6 + 1 + nuc].bins[b];
        cycles = c + 1;
    }
    return cycles;
}

/* :218-247 */
static int nth_value(const fxo_stats *s, int cycle, int nuc, int n)
{
    const nuc_data *d = &s->d[(size_t)cycle * 6 + nuc];
    if (n == 0) return d->min;
    if (n < 0 || n >= d->count) return -9999; /* reference: "Internal error" + exit(1) */
    int pos = 0;
    while (n > 0) {
        if (d->bins[pos] > n) break;
        n -= d->bins[pos];
        pos++;
        while (d->bins[pos] == 0) pos++;
    }
    return pos + FXO_MIN_Q;
}

static void whiskers(const fxo_stats *s, int cycle, int nuc, int *Q1, int *Q3, int *IQR, int *lw, int *rw)
{ /* :281-298 and :366-381 */
    const nuc_data *d = &s->d[(size_t)cycle * 6 + nuc];
    *Q1 = nth_value(s, cycle, nuc, d->count / 4);
    *Q3 = nth_value(s, cycle, nuc, d->count * 3 / 4);
    *IQR = *Q3 - *Q1;
    *lw = ((*Q1 - *IQR * 3 / 2) < d->min) ? d->min : (*Q1 - *IQR * 3 / 2);
    *rw = ((*Q3 + *IQR * 3 / 2) > d->max) ? d->max : (*Q3 + *IQR * 3 / 2);
}

static double runtime_div(double a, double b)
{ /* keep 0/0 a run-time division so its NaN sign matches the reference binary ("-nan") */
    volatile double x = a, y = b;
    return x / y;
}

int fxo_stats_print(const fxo_stats *s, FILE *out, int new_format)
{
    static const char *names[6] = { "ALL", "A", "C", "G", "T", "N" };
    static const char *cols[11] = { "count", "min", "max", "sum", "mean", "Q1", "med", "Q3", "IQR", "lW", "rW" };
    int Q1, Q3, IQR, lw, rw;
    if (new_format) { /* :316-347 */
        fprintf(out, "cycle\tmax_count");
        for (int nuc = 0; nuc < 6; nuc++)
            for (int k = 0; k < 11; k++) fprintf(out, "\t%s_%s", names[nuc], cols[k]);
        fprintf(out, "\n");
        int max_count = s->d[0].count;
        for (int c = 0; c < s->max_cycles; c++) {
            if (s->d[(size_t)c * 6].count == 0) break;
            fprintf(out, "%d\t%d", c + 1, max_count);
            for (int nuc = 0; nuc < 6; nuc++) {
                const nuc_data *d = &s->d[(size_t)c * 6 + nuc];
                whiskers(s, c, nuc, &Q1, &Q3, &IQR, &lw, &rw);
                fprintf(out, "\t%d\t%d\t%d\t%lld\t", d->count, d->min, d->max, (long long)d->sum);
                fprintf(out, "%3.2f\t%d\t%d\t%d\t", runtime_div((double)d->sum, (double)d->count), Q1,
                        nth_value(s, c, nuc, d->count / 2), Q3);
                fprintf(out, "%d\t%d\t%d", IQR, lw, rw);
            }
            fprintf(out, "\n");
        }
    } else { /* :349-417 */
        fprintf(out, "column\tcount\tmin\tmax\tsum\tmean\tQ1\tmed\tQ3\tIQR\tlW\trW\t"
                     "A_Count\tC_Count\tG_Count\tT_Count\tN_Count\tMax_count\n");
        for (int c = 0; c < s->max_cycles; c++) {
            const nuc_data *d = &s->d[(size_t)c * 6];
            if (d->count == 0) break;
            whiskers(s, c, 0, &Q1, &Q3, &IQR, &lw, &rw);
            fprintf(out, "%d\t", c + 1);
            fprintf(out, "%d\t%d\t%d\t%lld\t", d->count, d->min, d->max, (long long)d->sum);
            fprintf(out, "%3.2f\t%d\t%d\t%d\t", runtime_div((double)d->sum, (double)d->count), Q1,
                    nth_value(s, c, 0, d->count / 2), Q3);
            fprintf(out, "%d\t%d\t%d\t", IQR, lw, rw);
            fprintf(out, "%d\t%d\t%d\t%d\t%d\t", d[1].count, d[2].count, d[3].count, d[4].count, d[5].count);
            fprintf(out, "%d\n", s->d[0].count);
        }
    }
    return 0;
}

int fxo_stats_print_path(const fxo_stats *s, const char *path, int new_format)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fxo_stats_print(s, f, new_format);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ a6: half-local aligner --- */

enum { O_UP = 1, O_LEFT = 2, O_UPLEFT = 3 }; /* src/libfastx/sequence_alignment.h:78-84 */

/* src/libfastx/sequence_alignment.h:125-131 */
static char match_kind(uint8_t q, uint8_t t)
{
    if (q == 'N' || t == 'N') return 'N';
    return (q == t) ? 'M' : 'x';
}

/* src/libfastx/sequence_alignment.h:157-169; penalties from sequence_alignment.cpp:88-93 */
static float match_score(uint8_t q, uint8_t t)
{
    if (q == 'N' && t == 'N') return 0.0f;
    if (q == 'N' || t == 'N') return 0.1f;
    return (q == t) ? 1.0f : -1.0f;
}

void fxo_align(const uint8_t *Q, int len, int W, const uint8_t *T, int H, fxo_align_result *res)
{
    (void)len;
    const float gap = -5.0f;
    float *S = (float *)malloc(sizeof(float) * (size_t)W * (size_t)H);
    uint8_t *O = (uint8_t *)malloc((size_t)W * (size_t)H);
    float *tb = (float *)malloc(sizeof(float) * (size_t)(H + 1));
#define SC(x, y) S[(size_t)(x) * (size_t)H + (size_t)(y)]
#define OR(x, y) O[(size_t)(x) * (size_t)H + (size_t)(y)]
    /* reset_matrix, sequence_alignment.cpp:340-363: query border all 0; target border 0 for
     * y<=3 then gap*(y-3).  tb[-1] is an out-of-bounds read in the reference (safe_score(-1,-1),
     * sequence_alignment.h:147-155) that yields 0.0 in practice. */
    float *tbp = tb + 1;
    tbp[-1] = 0.0f;
    for (int y = 0; y < H; y++) tbp[y] = (y <= 3) ? 0.0f : gap * (float)(y - 3);

    /* populate_matrix, sequence_alignment.cpp:365-428 */
    float best = -1000000.0f;
    int bx = 0, by = 0;
    for (int x = 0; x < W; x++) {
        for (int y = 0; y < H; y++) {
            float up_in = (y > 0) ? SC(x, y - 1) : 0.0f /* query_border[x] */;
            float left_in = (x > 0) ? SC(x - 1, y) : tbp[y];
            float ul_in = (x > 0) ? ((y > 0) ? SC(x - 1, y - 1) : 0.0f /* query_border[x-1] */) : tbp[y - 1];
            float up = up_in + gap;
            float left = left_in + gap;
            float ul = ul_in + match_score(Q[x], T[y]);
            if (y > 3 && y - 3 > x) left = -100000.0f;
            float s = -100000000.0f;
            int o = O_LEFT;
            if (ul > s) { s = ul; o = O_UPLEFT; }
            if (up > s) { s = up; o = O_UP; }
            if (left > s) { s = left; o = O_LEFT; }
            SC(x, y) = s;
            OR(x, y) = (uint8_t)o;
            if (s > best) { best = s; bx = x; by = y; }
        }
    }

    /* find_optimal_alignment_from_point(bx, by), sequence_alignment.cpp:496-604.  The later
     * heuristics in find_optimal_alignment (:606-650) always keep this result. */
    memset(res, 0, sizeof *res);
    res->query_end = bx;
    res->target_end = by;
    res->score_at_best = best;
    int qi = bx, ti = by;
    while (qi >= 0 && ti >= 0) {
        res->query_start = qi;
        res->target_start = ti;
        switch (OR(qi, ti)) {
        case O_LEFT: res->gaps++; qi--; break;
        case O_UPLEFT:
            switch (match_kind(Q[qi], T[ti])) {
            case 'N': res->neutral++; break;
            case 'M': res->matches++; break;
            default:  res->mismatches++; break;
            }
            qi--; ti--;
            break;
        default /* O_UP */: res->gaps++; ti--; break;
        }
    }
#undef SC
#undef OR
    free(S); free(O); free(tb);
}

/* ------------------------------------------------------------------ a7: clipper decisions ---- */

/* src/fastx_clipper/fastx_clipper.cpp:159-241.  The reference mixes size_t fields with int
 * literals; the unsigned wrap of query_size-2 for 1-base reads is preserved. */
int fxo_adapter_cutoff_index(const fxo_align_result *r, int query_size_i, int min_adapter_len)
{
    size_t matches = (size_t)r->matches, mismatches = (size_t)r->mismatches;
    size_t query_end = (size_t)r->query_end, query_size = (size_t)query_size_i;
    size_t target_start = (size_t)r->target_start;
    int asz = (int)((size_t)r->neutral + matches + mismatches + (size_t)r->gaps);
    if (asz == 0) return -1;
    if (min_adapter_len > 0 && asz < min_adapter_len) return -1;
    if (query_end == query_size - 1 && mismatches == 0) return r->query_start;
    if (asz > 5 && target_start == 0 && (matches * 100 / (size_t)asz) >= 75) return r->query_start;
    if (asz > 11 && (matches * 100 / (size_t)asz) >= 80) return r->query_start;
    if (query_end >= query_size - 2 && asz <= 5 && matches >= 3) return r->query_start;
    return -1;
}

/* src/fastx_clipper/fastx_clipper.cpp:257-320 */
int fxo_clip_record(const uint8_t *row, int len, int width, const uint8_t *adapter, int alen,
                    const fxo_clip_opts *o, int *new_len, int *cut)
{
    fxo_align_result r;
    fxo_align(row, len, width, adapter, alen, &r);
    int i = fxo_adapter_cutoff_index(&r, len, o->min_adapter_len);
    int L = len;
    if (i != -1 && i > 0) {
        int at = i + o->keep_delta;
        if (at < L) L = at; /* nucleotides[at] = 0 only shortens when inside the string */
    }
    if (cut) *cut = i;
    if (new_len) *new_len = L;
    if (i == 0) return FXO_CLIP_ADAPTER_ONLY;
    if ((unsigned int)L < (unsigned int)o->min_length) return FXO_CLIP_TOO_SHORT;
    if (i == -1 && o->discard_non_clipped) return FXO_CLIP_NON_CLIPPED;
    if (i > 0 && o->discard_clipped) return FXO_CLIP_CLIPPED;
    if (o->discard_unknown && memchr(row, 'N', (size_t)L) != NULL) return FXO_CLIP_HAS_N;
    return FXO_CLIP_WRITE;
}

void fxo_clip_batch(const uint8_t *seq, const int32_t *len, const int32_t *width, int uniform_len,
                    int stride, int64_t n, const uint8_t *adapter, int alen, const fxo_clip_opts *o,
                    int32_t *out_len, uint8_t *out_class, int32_t *out_cut)
{
    for (int64_t i = 0; i < n; i++) {
        int L = rec_len(len, uniform_len, i);
        int W = width ? width[i] : L;
        int nl, cut;
        int cls = fxo_clip_record(seq + i * (int64_t)stride, L, W, adapter, alen, o, &nl, &cut);
        if (out_len) out_len[i] = nl;
        if (out_class) out_class[i] = (uint8_t)cls;
        if (out_cut) out_cut[i] = cut;
    }
}

/* ------------------------------------------------------------------ a8/a9: collapser --------- */

/* libstdc++ 13.3 libsupc++/hash_bytes.cc, 64-bit _Hash_bytes (Murmur-style).
 * std::hash<std::string> calls it with seed 0xc70f6907 (bits/functional_hash.h). */
uint64_t fxo_hash_bytes(const void *ptr, size_t len, uint64_t seed)
{
    const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
    const uint8_t *buf = (const uint8_t *)ptr;
    const size_t len_aligned = len & ~(size_t)7;
    uint64_t hash = seed ^ (len * mul);
    for (size_t p = 0; p < len_aligned; p += 8) {
        uint64_t w;
        memcpy(&w, buf + p, 8);
        uint64_t d = w * mul;
        d ^= d >> 47;
        d *= mul;
        hash ^= d;
        hash *= mul;
    }
    if (len & 7) {
        uint64_t d = 0;
        for (int k = (int)(len & 7) - 1; k >= 0; k--) d = (d << 8) + buf[len_aligned + (size_t)k];
        hash ^= d;
        hash *= mul;
    }
    hash ^= hash >> 47;
    hash *= mul;
    hash ^= hash >> 47;
    return hash;
}

/* Bucket-count ladder produced by _Prime_rehash_policy::_M_need_rehash/_M_next_bkt when a map
 * grows from empty by single insertions at max_load_factor 1.0 (hashtable_c++0x.cc): first
 * insertion asks for >= 12 buckets -> 13; afterwards next prime in __prime_list >= 2*buckets. */
static const uint64_t bucket_ladder[] = {
    13ull, 29ull, 59ull, 127ull, 257ull, 541ull, 1109ull, 2357ull, 5087ull, 10273ull, 20753ull, 42043ull,
    85229ull, 172933ull, 351061ull, 712697ull, 1447153ull, 2938679ull, 5967347ull, 12117689ull,
    24607243ull, 49969847ull, 101473717ull, 206062531ull, 418451333ull, 849749479ull, 1725587117ull,
    3504151727ull
};

typedef struct cnode {
    struct cnode *next;
    uint64_t hash, count;
    int64_t first;
    int len;
    uint8_t *key;
} cnode;

struct fxo_collapser {
    cnode before_begin;  /* _M_before_begin */
    cnode **buckets;     /* each holds the node BEFORE the bucket's first node */
    uint64_t nb;         /* _M_bucket_count (starts at 1: the single bucket) */
    uint64_t size;       /* _M_element_count */
    int ladder_pos;
    int64_t adds;
};

fxo_collapser *fxo_collapser_new(void)
{
    fxo_collapser *c = (fxo_collapser *)calloc(1, sizeof *c);
    c->nb = 1;
    c->buckets = (cnode **)calloc(1, sizeof(cnode *));
    c->ladder_pos = -1;
    return c;
}

void fxo_collapser_free(fxo_collapser *c)
{
    if (!c) return;
    for (cnode *p = c->before_begin.next; p;) { cnode *n = p->next; free(p->key); free(p); p = n; }
    free(c->buckets);
    free(c);
}

/* hashtable.h _M_insert_bucket_begin: front of a non-empty bucket, else new global head */
static void insert_bucket_begin(fxo_collapser *c, cnode **buckets, uint64_t nb, uint64_t bkt, cnode *node)
{
    if (buckets[bkt]) {
        node->next = buckets[bkt]->next;
        buckets[bkt]->next = node;
    } else {
        node->next = c->before_begin.next;
        c->before_begin.next = node;
        if (node->next) buckets[node->next->hash % nb] = node;
        buckets[bkt] = &c->before_begin;
    }
}

/* hashtable.h _M_rehash_aux(unique keys): re-thread nodes in current list order with the same rule */
static void rehash(fxo_collapser *c, uint64_t nb)
{
    cnode **nbk = (cnode **)calloc((size_t)nb, sizeof(cnode *));
    cnode *p = c->before_begin.next;
    c->before_begin.next = NULL;
    uint64_t bbegin_bkt = 0;
    while (p) {
        cnode *next = p->next;
        uint64_t bkt = p->hash % nb;
        if (!nbk[bkt]) {
            p->next = c->before_begin.next;
            c->before_begin.next = p;
            nbk[bkt] = &c->before_begin;
            if (p->next) nbk[bbegin_bkt] = p;
            bbegin_bkt = bkt;
        } else {
            p->next = nbk[bkt]->next;
            nbk[bkt]->next = p;
        }
        p = next;
    }
    free(c->buckets);
    c->buckets = nbk;
    c->nb = nb;
}

/* src/fastx_collapser/fastx_collapser.cpp:112-114: collapsed_sequences[string(seq)] += reads_count */
void fxo_collapser_add(fxo_collapser *c, const uint8_t *seq, int len, uint64_t weight)
{
    uint64_t h = fxo_hash_bytes(seq, (size_t)len, 0xc70f6907ull);
    uint64_t bkt = h % c->nb;
    int64_t idx = c->adds++;
    cnode *prev = c->buckets[bkt];
    if (prev) { /* _M_find_before_node */
        for (cnode *p = prev->next;; p = p->next) {
            if (p->hash == h && p->len == len && memcmp(p->key, seq, (size_t)len) == 0) { p->count += weight; return; }
            if (!p->next || p->next->hash % c->nb != bkt) break;
        }
    }
    /* _M_insert_unique_node: rehash first when size+1 exceeds the load limit (== bucket count) */
    if (c->ladder_pos < 0 || c->size + 1 > c->nb) {
        c->ladder_pos++;
        rehash(c, bucket_ladder[c->ladder_pos]);
        bkt = h % c->nb;
    }
    cnode *node = (cnode *)calloc(1, sizeof *node);
    node->hash = h; node->count = weight; node->first = idx; node->len = len;
    node->key = (uint8_t *)malloc((size_t)len + 1);
    memcpy(node->key, seq, (size_t)len);
    node->key[len] = 0;
    insert_bucket_begin(c, c->buckets, c->nb, bkt, node);
    c->size++;
}

void fxo_collapser_add_batch(fxo_collapser *c, const uint8_t *seq, const int32_t *len, int uniform_len,
                             int stride, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        fxo_collapser_add(c, seq + i * (int64_t)stride, rec_len(len, uniform_len, i), 1);
}

int64_t fxo_collapser_unique(const fxo_collapser *c) { return (int64_t)c->size; }

typedef struct { uint64_t count; int64_t iterpos; cnode *n; } ord_t;

static int ord_cmp(const void *a, const void *b)
{ /* count descending; ties: later iteration position first (stable ascending sort printed in
     reverse: fastx_collapser.cpp:116-122) */
    const ord_t *x = (const ord_t *)a, *y = (const ord_t *)b;
    if (x->count != y->count) return (x->count > y->count) ? -1 : 1;
    if (x->iterpos != y->iterpos) return (x->iterpos > y->iterpos) ? -1 : 1;
    return 0;
}

static ord_t *collapser_sorted(fxo_collapser *c)
{
    ord_t *o = (ord_t *)malloc(sizeof(ord_t) * (size_t)(c->size ? c->size : 1));
    int64_t k = 0;
    for (cnode *p = c->before_begin.next; p; p = p->next, k++) { o[k].count = p->count; o[k].iterpos = k; o[k].n = p; }
    qsort(o, (size_t)c->size, sizeof(ord_t), ord_cmp);
    return o;
}

void fxo_collapser_order(fxo_collapser *c, int64_t *first_index, uint64_t *count)
{
    ord_t *o = collapser_sorted(c);
    for (uint64_t k = 0; k < c->size; k++) { first_index[k] = o[k].n->first; count[k] = o[k].count; }
    free(o);
}

int fxo_collapser_print_path(fxo_collapser *c, const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    ord_t *o = collapser_sorted(c);
    /* fastx_collapser.cpp:80-85: ">rank-count\nSEQ\n"; the count is narrowed to int */
    for (uint64_t k = 0; k < c->size; k++)
        fprintf(f, ">%llu-%d\n%s\n", (unsigned long long)(k + 1), (int)o[k].count, (const char *)o[k].n->key);
    free(o);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ (f-2) rows ---------------- */

/* src/fastq_masker/fastq_masker.c:92-107 */
void fxo_mask_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len, int stride, int64_t n,
                    int q_offset, int min_quality, int mask_char, uint8_t *out_seq, uint8_t *masked_flag,
                    int64_t *masked_reads, int64_t *masked_bases)
{
    int64_t mr = 0, mb = 0;
    for (int64_t i = 0; i < n; i++) {
        const int L = rec_len(len, uniform_len, i);
        const int64_t o = i * (int64_t)stride;
        int masked = 0;
        memset(out_seq + o, 0, (size_t)stride);
        for (int k = 0; k < L; k++) {
            const int q = (int)(signed char)qual[o + k] - q_offset;
            if (q < min_quality) { out_seq[o + k] = (uint8_t)mask_char; masked = 1; mb++; }
            else out_seq[o + k] = seq[o + k];
        }
        if (masked_flag) masked_flag[i] = (uint8_t)masked;
        mr += masked;
    }
    if (masked_reads) *masked_reads = mr;
    if (masked_bases) *masked_bases = mb;
}

/* src/fastx_artifacts_filter/fastx_artifacts_filter.c:56-114 */
void fxo_artifacts_batch(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint8_t *keep)
{
    for (int64_t i = 0; i < n; i++) {
        const int L = rec_len(len, uniform_len, i);
        int a = 0, c = 0, g = 0, t = 0, total = 0;
        for (int k = 0; k < L; k++) {
            total++;
            switch (seq[i * (int64_t)stride + k]) {
            case 'A': a++; break;
            case 'C': c++; break;
            case 'G': g++; break;
            case 'T': t++; break;
            default: break;   /* 'N' */
            }
        }
        const int lim = total - 3;
        keep[i] = (a >= lim || c >= lim || g >= lim || t >= lim) ? 0 : 1;
    }
}

/* scripts/fastx_barcode_splitter.pl:296 — length(a) - (number of NUL bytes of the string XOR): the XOR is as long as the
 * longer string and a position counts as equal only where both strings have the same character */
static int fxo_perl_mismatch_count(const uint8_t *a, int la, const uint8_t *b, int lb)
{
    int zeros = 0;
    const int lx = la > lb ? la : lb;
    for (int p = 0; p < lx; p++) {
        const uint8_t ca = p < la ? a[p] : 0, cb = p < lb ? b[p] : 0;
        if ((ca ^ cb) == 0) zeros++;
    }
    return la - zeros;
}

/* scripts/fastx_barcode_splitter.pl:232-275 */
int fxo_barcode_match(const uint8_t *fragment, int frag_len, const uint8_t *const *entries, const int32_t *entry_len, int n_entries,
                      int barcode_len, int allowed_mismatches)
{
    int best_mm = barcode_len, best = -1;
    for (int e = 0; e < n_entries; e++) {
        int mm = fxo_perl_mismatch_count(fragment, frag_len, entries[e], entry_len[e]);
        mm += barcode_len - entry_len[e];                    /* partial entries: the missing bases are mismatches */
        if (mm < best_mm) { best_mm = mm; best = e; }
    }
    if (best < 0 || best_mm > allowed_mismatches) return -1;
    return best;
}

/* src/fastq_to_fasta/fastq_to_fasta.c:79-82: the discard test of fastq_to_fasta */
void fxo_has_n_batch(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint8_t *has_n)
{
    for (int64_t i = 0; i < n; i++) {
        const int L = rec_len(len, uniform_len, i);
        uint8_t f = 0;
        for (int k = 0; k < L; k++)
            if (seq[i * (int64_t)stride + k] == 'N') { f = 1; break; }
        has_n[i] = f;
    }
}

/* src/fastx_trimmer/fastx_trimmer.c:120-148 — pointer arithmetic only */
int fxo_fastx_trimmer_record(int len, int first, int last, int trim_last, int min_len, int *start)
{
    int L = len, s0 = 0;
    if (last != 0 && last < L) L = last;                 /* nucleotides[keep_last_base] = 0 */
    if (first != 1) {
        if (L < first) return -1;                        /* sequence too short - remove it */
        s0 = first - 1;
        L = L - first + 1;
    }
    if (trim_last > 0) {
        if (L <= trim_last) return -1;
        const int i = L - trim_last;
        if (i < min_len) return -1;
        L = i;
    }
    if (start) *start = s0;
    return L;
}
