/*
 * fastx_oracle.h — CPU restatement of the FASTX-Toolkit 0.0.14 hot path.
 *
 * *** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libfxg.so, bin/ tools) never links, loads or executes it and
 * has no CPU fallback.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Pinning: tests/test_oracle_golden.py replays the reference's own Galaxy fixtures
 * (tests/golden/reference_fixtures/, copied from galaxy/test-data/) through these functions and,
 * where oracle/_ref/ (the unmodified reference compiled by oracle/Makefile) is present, compares
 * them with the real binaries on seeded synthetic input.  The collapser's tie order and the
 * stats "-N" format are NOT pinned by any reference fixture; they are pinned only
 * against the oracle/_ref binaries (libstdc++ 13.3).
 */
#ifndef FASTX_ORACLE_H
#define FASTX_ORACLE_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXO_MIN_Q (-15)   /* src/libfastx/fastx.h:28 */
#define FXO_MAX_Q 93      /* src/libfastx/fastx.h:29 */
#define FXO_QBINS 109     /* -15..93 inclusive; the reference's arrays are 108 long */

/* ---- a1: record validation (src/libfastx/fastx.c:45-54, 118-135) ---- */
int fxo_seq_first_invalid(const uint8_t *seq, int len);               /* -1 if all in {A,C,G,T,N} */
int fxo_qual_first_invalid(const uint8_t *qual, int len, int q_offset); /* -1 if all in [-15,93]   */

/* ---- a2: fastq_quality_trimmer body (src/fastq_quality_trimmer/fastq_quality_trimmer.c:93-102) */
int fxo_trim_record(const uint8_t *qual, int len, int q_offset, int threshold, int min_len);

/* ---- a3: fastq_quality_filter (src/fastq_quality_filter/fastq_quality_filter.c:78-129,150-156) */
int fxo_filter_record(const uint8_t *qual, int len, int q_offset, int min_quality, int min_percent);

/* ---- a4: fastx_reverse_complement (src/fastx_reverse_complement/fastx_reverse_complement.c:43-104) */
void fxo_revcomp_record(const uint8_t *seq, const uint8_t *qual, int len, uint8_t *oseq, uint8_t *oqual);

/* ---- batch forms over a fixed-stride slab (read i at base + i*stride; len==NULL => uniform_len) */
void fxo_trim_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                    int stride, int64_t n, int q_offset, int threshold, int min_len,
                    int32_t *out_len, int64_t *first_bad);
void fxo_filter_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                      int stride, int64_t n, int q_offset, int min_quality, int min_percent,
                      uint8_t *keep, int64_t *first_bad);
void fxo_revcomp_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len,
                       int stride, int64_t n, uint8_t *oseq, uint8_t *oqual);

/* ---- a5: fastx_quality_stats (src/fastx_quality_stats/fastx_quality_stats.c:115-247,276-417) */
typedef struct fxo_stats fxo_stats;
fxo_stats *fxo_stats_new(int max_cycles);
void fxo_stats_free(fxo_stats *s);
/* qual == NULL => FASTA record (counts only), weight = get_reads_count() */
void fxo_stats_add(fxo_stats *s, const uint8_t *seq, const uint8_t *qual, int len, int q_offset, int weight);
void fxo_stats_add_batch(fxo_stats *s, const uint8_t *seq, const uint8_t *qual, const int32_t *len,
                         int uniform_len, int stride, int64_t n, int q_offset);
/* Export as u64 hist[cycles][5 (A,C,G,T,N)][109]; returns cycles written */
int fxo_stats_export_hist(const fxo_stats *s, uint64_t *hist, int max_cycles);
int fxo_stats_print(const fxo_stats *s, FILE *out, int new_format);
int fxo_stats_print_path(const fxo_stats *s, const char *path, int new_format);

/* ---- a6/a7: fastx_clipper (src/libfastx/sequence_alignment.cpp:340-428,496-650,
 *                            src/fastx_clipper/fastx_clipper.cpp:159-241,257-320) ---- */
typedef struct {
    int matches, mismatches, neutral, gaps;
    int query_start, query_end, target_start, target_end;
    float score_at_best;
} fxo_align_result;

/* query row has `width` readable columns (width >= len; columns >= len hold NUL then stale bytes:
 * SURVEY.md Appendix D.1).  Restates the DP + origin matrix + backtrace literally. */
void fxo_align(const uint8_t *query_row, int len, int width, const uint8_t *adapter, int alen,
               fxo_align_result *res);
int fxo_adapter_cutoff_index(const fxo_align_result *r, int query_size, int min_adapter_len);

/* Outcome classes of the discard cascade, in the reference's order */
enum { FXO_CLIP_WRITE = 0, FXO_CLIP_ADAPTER_ONLY = 1, FXO_CLIP_TOO_SHORT = 2, FXO_CLIP_NON_CLIPPED = 3,
       FXO_CLIP_CLIPPED = 4, FXO_CLIP_HAS_N = 5 };
typedef struct {
    int min_length;          /* -l, default 5 */
    int keep_delta;          /* -d N (>0 => N + strlen(adapter) added by caller, as parse_commandline does) */
    int discard_non_clipped; /* -c */
    int discard_clipped;     /* -C */
    int discard_unknown;     /* default 1; -n => 0 */
    int min_adapter_len;     /* -M */
} fxo_clip_opts;
/* returns class; *new_len = length the writer would emit (strlen after truncation); *cut = cutoff index */
int fxo_clip_record(const uint8_t *query_row, int len, int width, const uint8_t *adapter, int alen,
                    const fxo_clip_opts *o, int *new_len, int *cut);
void fxo_clip_batch(const uint8_t *seq, const int32_t *len, const int32_t *width, int uniform_len,
                    int stride, int64_t n, const uint8_t *adapter, int alen, const fxo_clip_opts *o,
                    int32_t *out_len, uint8_t *out_class, int32_t *out_cut);

/* ---- a8/a9: fastx_collapser (src/fastx_collapser/fastx_collapser.cpp:51-54,80-91,112-122)
 *      + libstdc++ 13.3 std::unordered_map<std::string,size_t> (hashtable.h, hashtable_policy.h,
 *      hash_bytes.cc) — third-party, not under /root/reference; restated from its published algorithm. */
uint64_t fxo_hash_bytes(const void *p, size_t len, uint64_t seed);  /* std::_Hash_bytes, 64-bit */
typedef struct fxo_collapser fxo_collapser;
fxo_collapser *fxo_collapser_new(void);
void fxo_collapser_free(fxo_collapser *c);
void fxo_collapser_add(fxo_collapser *c, const uint8_t *seq, int len, uint64_t weight);
void fxo_collapser_add_batch(fxo_collapser *c, const uint8_t *seq, const int32_t *len, int uniform_len,
                             int stride, int64_t n);
int64_t fxo_collapser_unique(const fxo_collapser *c);
/* Final order (count desc, ties = reverse map iteration order).  For rank k (0-based):
 * first_index[k] = index of the add() call that first inserted the key, count[k] = total weight. */
void fxo_collapser_order(fxo_collapser *c, int64_t *first_index, uint64_t *count);
int fxo_collapser_print_path(fxo_collapser *c, const char *path);

/* ---- (f-2) rows: fastq_masker (src/fastq_masker/fastq_masker.c:92-107), fastx_artifacts_filter
 * (src/fastx_artifacts_filter/fastx_artifacts_filter.c:56-114), fastx_trimmer (src/fastx_trimmer/fastx_trimmer.c:120-148) */
void fxo_mask_batch(const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len, int stride, int64_t n,
                    int q_offset, int min_quality, int mask_char, uint8_t *out_seq, uint8_t *masked_flag,
                    int64_t *masked_reads, int64_t *masked_bases);
void fxo_artifacts_batch(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint8_t *keep);
/* (f-4) barcode splitter, scripts/fastx_barcode_splitter.pl:208-290 + mismatch_count (:296): index of the entry that gets
 * the fragment, -1 = 'unmatched'.  entries: the script's @barcodes in order (each barcode, then its --partial forms). */
int fxo_barcode_match(const uint8_t *fragment, int frag_len, const uint8_t *const *entries, const int32_t *entry_len, int n_entries,
                      int barcode_len, int allowed_mismatches);
/* (f-4) fastq_to_fasta (src/fastq_to_fasta/fastq_to_fasta.c:79-82): has_n[i] = strchr(nucleotides, 'N') != NULL */
void fxo_has_n_batch(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint8_t *has_n);
/* first/last: -f/-l (1-based, last 0 = none); trim_last/min_len: -t/-m.  Returns the new length (>=0) and *start, or -1 = discard */
int fxo_fastx_trimmer_record(int len, int first, int last, int trim_last, int min_len, int *start);

#ifdef __cplusplus
}
#endif
#endif
