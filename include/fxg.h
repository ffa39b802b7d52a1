/*
 * fxg.h — thin C ABI of the B200-native FASTX hot path (libfxg.so).
 *
 * The reference (agordon/fastx_toolkit 0.0.14) has no plugin/FFI interface: its only seams are the
 * six executables and libfastx's record API (src/libfastx/fastx.h:120-142).  This header is the
 * boundary the drop-in executables in bin/ (host C, fastx_toolkit_b200/csrc/host/) call instead of
 * running the per-read transform loop on the CPU.  Each entry point names the reference loop body
 * it replaces.  Plain C structs, plain pointers and sizes, int error codes, no exceptions.
 *
 * Conventions
 *   - A *batch* is a structure-of-arrays slab: read i occupies bytes [i*stride, i*stride+len_i) of
 *     `seq` and of `qual` (raw FASTQ bytes, i.e. ASCII quality, not yet offset by -Q).  `stride`
 *     is a multiple of 16 and both base pointers are 16-byte aligned.  Padding bytes are ignored.
 *   - `*_dev` entry points take DEVICE pointers and only enqueue work on the context's stream;
 *     results are valid after fxg_sync().  `*_host` entry points take HOST pointers (pinned for
 *     full speed: fxg_alloc_pinned / fxg_host_register), pipeline H2D copy -> kernel -> D2H copy
 *     over chunks on side streams, and return when the results are in host memory.
 *   - A context is bound to one GPU and is single-threaded; contexts are independent (one per GPU).
 *   - There is NO CPU fallback: without a usable GPU every call fails with FXG_ERR_CUDA.
 *   - Supported quality offsets: 15 <= q_offset <= 127 (Phred+33 and +64 are the real ones).
 */
#ifndef FXG_H
#define FXG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXG_OK               0
#define FXG_ERR_CUDA         1   /* CUDA runtime/driver error (no device, launch failure, ...) */
#define FXG_ERR_ARG          2   /* bad argument (alignment, stride, NULL, unsupported -Q)      */
#define FXG_ERR_NOMEM        3
#define FXG_ERR_UNSUPPORTED  4
#define FXG_ERR_NCCL         5

#define FXG_QBINS            109 /* quality values -15..93 (src/libfastx/fastx.h:28-29)         */
#define FXG_MAX_ADAPTER      100 /* src/fastx_clipper/fastx_clipper.cpp:38 MAX_ADAPTER_LEN       */

typedef struct fxg_ctx fxg_ctx;

/* SoA slab of reads (see Conventions). len == NULL means every read has uniform_len bases. */
typedef struct {
    const uint8_t *seq;
    const uint8_t *qual;      /* NULL for FASTA input (only ops that accept it)                 */
    const int32_t *len;
    int32_t        uniform_len;
    int32_t        stride;
    int64_t        n;
} fxg_batch;

/* Filled by fxg_sync() / by the *_host calls. Counters accumulate until fxg_report_reset(). */
typedef struct {
    int64_t n_in;             /* reads processed                                                */
    int64_t n_out;            /* reads kept / emitted                                           */
    int64_t first_bad_read;   /* smallest batch-relative index (+base) of a read that fails the
                                 reader's validation (fastx.c:45-54,118-135,361-362); -1 = none */
    int64_t aux[6];           /* op specific (clipper: discard classes, see fxg_clip_*)         */
} fxg_report;

/* ---- context, memory -------------------------------------------------------------------- */
int         fxg_init(int device, fxg_ctx **out);
void        fxg_destroy(fxg_ctx *ctx);
const char *fxg_strerror(int code);
const char *fxg_last_error(const fxg_ctx *ctx);          /* detail of the last failure           */
int         fxg_device_info(fxg_ctx *ctx, int *sm_count, size_t *hbm_bytes, int *cc_major, int *cc_minor);
int         fxg_set_stream(fxg_ctx *ctx, void *cuda_stream);  /* adopt the caller's cudaStream_t for *_dev calls
                                                                 (NULL is the legacy default stream)      */
int         fxg_use_own_stream(fxg_ctx *ctx);                 /* back to the context's private stream  */
int         fxg_sync(fxg_ctx *ctx);                      /* wait, then refresh the report        */
int         fxg_get_report(fxg_ctx *ctx, fxg_report *out);
int         fxg_report_reset(fxg_ctx *ctx);
int64_t     fxg_kernel_launches(const fxg_ctx *ctx);     /* kernels launched by this context     */

void       *fxg_alloc_pinned(size_t bytes);
void        fxg_free_pinned(void *p);
int         fxg_host_register(void *p, size_t bytes);
int         fxg_host_unregister(void *p);
void       *fxg_alloc_device(fxg_ctx *ctx, size_t bytes);
void        fxg_free_device(fxg_ctx *ctx, void *p);
int         fxg_memcpy_h2d(fxg_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int         fxg_memcpy_d2h(fxg_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int         fxg_memset_dev(fxg_ctx *ctx, void *dst_dev, int value, size_t bytes);

/* Tuning knobs (0 = library default). tile_reads: reads per TMA tile; stages: smem ring depth;
 * ctas_per_sm: persistent CTAs per SM. */
int         fxg_set_tuning(fxg_ctx *ctx, int tile_reads, int stages, int ctas_per_sm);

/* ---- synthetic workload (include/fxg_synth.h), generated on the device ------------------- */
int fxg_synth_dev(fxg_ctx *ctx, uint8_t *seq_dev, uint8_t *qual_dev, int64_t n, int64_t first_read,
                  int64_t n_total, int32_t len, int32_t stride, uint64_t seed, int kind, int q_offset);

/* ---- a2: fastq_quality_trimmer loop body (src/fastq_quality_trimmer/fastq_quality_trimmer.c:91-103)
 * out_len[i] = surviving prefix length, or -1 when the read is discarded.  batch->seq may be NULL
 * to skip base validation ("decide-only", L+4 bytes/read); the full form reads 2L+4 bytes/read. */
int fxg_trim_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, int threshold, int min_len,
                  int32_t *out_len_dev, int64_t index_base);
int fxg_trim_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int threshold, int min_len,
                  int32_t *out_len_host, fxg_report *report);

/* ---- a3: fastq_quality_filter loop body (src/fastq_quality_filter/fastq_quality_filter.c:78-129,141-161)
 * keep[i] = 1 iff the read passes (-q min_quality, -p min_percent). */
int fxg_filter_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int min_percent,
                    uint8_t *keep_dev, int64_t index_base);
int fxg_filter_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int min_percent,
                    uint8_t *keep_host, fxg_report *report);

/* ---- a4: fastx_reverse_complement loop body (src/fastx_reverse_complement/fastx_reverse_complement.c:43-104)
 * Writes the reverse complement (and reversed qualities) with the same stride; padding is zeroed. */
int fxg_revcomp_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *out_seq_dev,
                     uint8_t *out_qual_dev, int64_t index_base);
int fxg_revcomp_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *out_seq_host,
                     uint8_t *out_qual_host, fxg_report *report);

/* ---- a5: fastx_quality_stats accumulation (src/fastx_quality_stats/fastx_quality_stats.c:166-216)
 * hist[cycle][nuc][q+15] += 1 (u64, nuc order A,C,G,T,N, FXG_QBINS bins) for every base of every read;
 * the table lives in device memory (fxg_alloc_device, zeroed by the caller) and is ACCUMULATED, so it can
 * be all-reduced across GPUs before the host derives count/min/max/sum/quartiles from it.  For FASTA
 * batches (qual == NULL) every count lands in the q = 0 bin.  weight (may be NULL) = per-read count for
 * collapsed FASTA ids (get_reads_count, src/libfastx/fastx.c:475-497). */
int fxg_stats_accum_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist_dev, int32_t max_cycles,
                         const int32_t *weight_dev, int64_t index_base);
int fxg_stats_accum_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist_dev, int32_t max_cycles,
                         const int32_t *weight_host, fxg_report *report);

/* ---- a6/a7: fastx_clipper loop body: HalfLocalSequenceAlignment::align (src/libfastx/sequence_alignment.cpp:
 * 113-129,340-428,496-650) + adapter_cutoff_index (src/fastx_clipper/fastx_clipper.cpp:159-241) + the discard
 * cascade (:280-319).  width[i] (may be NULL => len) is the DP matrix width of read i: the reference's matrix
 * only grows, so for mixed-length input it is the running maximum length and the row must hold the bytes the
 * reference would read there (NUL at len, then stale bytes of earlier reads — SURVEY.md Appendix D.1).
 * out_len[i] = length to emit, or -1 when discarded; out_class[i] (may be NULL) = FXG_CLIP_* class. */
#define FXG_CLIP_WRITE        0
#define FXG_CLIP_ADAPTER_ONLY 1
#define FXG_CLIP_TOO_SHORT    2
#define FXG_CLIP_NON_CLIPPED  3
#define FXG_CLIP_CLIPPED      4
#define FXG_CLIP_HAS_N        5
typedef struct {
    const char *adapter;          /* -a, NUL terminated, 1..99 characters                              */
    int32_t min_length;           /* -l (default 5)                                                    */
    int32_t keep_delta;           /* -d N > 0 ? N + strlen(adapter) : 0   (fastx_clipper.cpp:153-154)  */
    int32_t discard_non_clipped;  /* -c                                                                */
    int32_t discard_clipped;      /* -C                                                                */
    int32_t discard_unknown;      /* 1 unless -n                                                       */
    int32_t min_adapter_len;      /* -M                                                                */
} fxg_clip_opts;
/* report.n_out = reads written; report.aux[FXG_CLIP_*] = reads in each discard class */
int fxg_clip_dev (fxg_ctx *ctx, const fxg_batch *b, const int32_t *width_dev, int q_offset, const fxg_clip_opts *o,
                  int32_t *out_len_dev, uint8_t *out_class_dev, int32_t *out_cut_dev, int64_t index_base);
int fxg_clip_host(fxg_ctx *ctx, const fxg_batch *b, const int32_t *width_host, int q_offset, const fxg_clip_opts *o,
                  int32_t *out_len_host, uint8_t *out_class_host, fxg_report *report);

/* ---- a8/a9: fastx_collapser (src/fastx_collapser/fastx_collapser.cpp:51-54,80-91,112-122)
 * K-HASH: hash_dev[i] = std::hash<std::string>(read i) (libstdc++ _Hash_bytes, seed 0xc70f6907) — the value
 * that fixes both the owner GPU (hash mod G) and the reference's tie order. */
int fxg_hash_dev(fxg_ctx *ctx, const fxg_batch *b, uint64_t *hash_dev);

/* Exact dedup table resident on one GPU.  add(): append the batch's rows (HOST or DEVICE pointers) and count
 * them: weight[i] (NULL = 1) is added to the key's count (collapsed-FASTA "N-COUNT" ids, or partial counts from
 * another GPU); first[i] (NULL = index_base + i) is the key's first-occurrence index candidate (the minimum
 * wins).  finish(order=1) computes the reference's output order (count descending, ties in reverse
 * std::unordered_map iteration order); fetch() copies the uniques out in that order. */
typedef struct fxg_collapser fxg_collapser;
/* max_reads and stride are starting sizes: add() grows the row store (and its stride, when a batch with longer rows
 * arrives) and rebuilds the table when it would get more than half full. */
int         fxg_collapse_new(int device, int64_t max_reads, int32_t stride, fxg_collapser **out);
void        fxg_collapse_free(fxg_collapser *c);
int         fxg_collapse_add(fxg_collapser *c, const fxg_batch *b, const int32_t *weight, const int64_t *first, int64_t index_base);
int         fxg_collapse_reserve(fxg_collapser *c, int64_t rows, int32_t stride);   /* grow now (e.g. to a common stride) */
int         fxg_collapse_add_next(fxg_collapser *c, const fxg_batch *b);   /* add(): weight 1, first = rows added so far + i */
int         fxg_collapse_finish(fxg_collapser *c, int order, int64_t *n_unique, int64_t *first_bad_read);
int         fxg_collapse_fetch(fxg_collapser *c, uint8_t *out_seq, int32_t *out_len, uint64_t *out_count, int64_t *out_first,
                               uint64_t *out_hash);
const char *fxg_collapse_error(const fxg_collapser *c);
int64_t     fxg_collapse_launches(const fxg_collapser *c);
int32_t     fxg_collapse_stride(const fxg_collapser *c);      /* row pitch of fetch()'s out_seq (may have grown)  */
/* K-ORDER alone: perm_dev[k] = index of the unique printed at rank k, from (hash, first, count) triples that
 * may have been gathered from several GPUs (device pointers on `device`). */
int         fxg_collapse_order_dev(int device, const uint64_t *hash_dev, const uint64_t *first_dev, const uint64_t *count_dev,
                                   int64_t n_unique, uint32_t *perm_dev);

/* ---- the native collectives ------------------------------------------------------------------------
 * A communicator drives the GPUs THIS PROCESS owns: every GPU of the box for the drop-in tools (fxg_comm_init_all:
 * ncclCommInitAll over `devices`), or one GPU per process for torchrun / mpirun style jobs (fxg_comm_init_rank: rank 0
 * makes the 128-byte id with fxg_comm_unique_id() and whatever launched the job hands it to the other ranks).  NCCL is
 * dlopen()ed at the first call.  Collectives are enqueued on the communicator's own streams (after a device-wide wait
 * for earlier work), or on streams adopted from the caller (fxg_comm_set_stream, then plain stream order applies);
 * fxg_comm_sync() waits for them.
 *   fxg_comm_allreduce_u64  the fastx_quality_stats reduction: in-place ncclSum of one u64 histogram per local GPU
 *                           (src/fastx_quality_stats/fastx_quality_stats.c:166-216 accumulates ONE table; G GPUs hold G partials)
 *   fxg_comm_allgather      small fixed-size all-gather (count matrices)
 *   fxg_comm_alltoallv      the collapser's owner exchange: grouped ncclSend/ncclRecv with per-peer counts/offsets
 *                           (elements of elem_bytes bytes; arrays indexed [local GPU * nranks + peer], host memory)
 *   fxg_comm_gatherv        variable-size gather to one rank (cnt/off: nranks entries, elements)                          */
#define FXG_COMM_ID_BYTES 128
typedef struct fxg_comm fxg_comm;
int         fxg_comm_init_all(int ndev, const int *devices, fxg_comm **out);
int         fxg_comm_unique_id(void *id_out /* FXG_COMM_ID_BYTES */);
int         fxg_comm_init_rank(int device, int nranks, int rank, const void *id, fxg_comm **out);
int         fxg_comm_nranks(const fxg_comm *c);
int         fxg_comm_nlocal(const fxg_comm *c);
int         fxg_comm_rank(const fxg_comm *c, int local_index);
int         fxg_comm_device(const fxg_comm *c, int local_index);
int         fxg_comm_set_stream(fxg_comm *c, int local_index, void *cuda_stream, int adopt);
int         fxg_comm_sync(fxg_comm *c);
int         fxg_comm_allreduce_u64(fxg_comm *c, uint64_t *const *bufs_dev, size_t count);
int         fxg_comm_allgather(fxg_comm *c, const void *const *send_dev, void *const *recv_dev, size_t bytes);
int         fxg_comm_alltoallv(fxg_comm *c, const void *const *send_dev, const int64_t *send_off, const int64_t *send_cnt,
                               void *const *recv_dev, const int64_t *recv_off, const int64_t *recv_cnt, size_t elem_bytes);
int         fxg_comm_gatherv(fxg_comm *c, const void *const *send_dev, const int64_t *cnt, const int64_t *off, void *root_recv_dev,
                             int root, size_t elem_bytes);
int64_t     fxg_comm_bytes_sent(const fxg_comm *c);      /* payload bytes this process sent to OTHER ranks              */
int64_t     fxg_comm_collectives(const fxg_comm *c);     /* NCCL groups issued                                            */
void        fxg_comm_free(fxg_comm *c);
const char *fxg_comm_error(const fxg_comm *c);

/* ---- a8/a9 across GPUs: the collapser's global count map (src/fastx_collapser/fastx_collapser.cpp:112-114) partitioned by
 * owner = std::hash(sequence) mod nranks, then the reference's output order (:116-122) computed once on the root GPU.
 * run(): batches[i] = DEVICE slabs on the communicator's i-th local GPU (seq only; len == NULL: uniform_len), global read
 * index of row r = index_base[i] + r (or first_dev[i][r] when given: rows that are already partial results, e.g. the uniques of
 * a per-GPU fxg_collapser with their first indices), weight_dev (or weight_dev[i]) NULL = 1 per read.  Phases: K-ROUTE (hash, owner, send
 * slabs) -> exchange of key rows + 16-byte {first, weight, len} records -> K-DEDUP on the owners -> gather of the uniques'
 * (hash, first, count) to the root -> K-ORDER.  Key rows never leave their owner: fetch_local() returns an owner's uniques
 * in its table order, fetch_order() (root's process) says which (owner, index) is printed at every rank of the output.    */
typedef struct {
    int64_t n_unique;          /* uniques in the whole job                                                              */
    int64_t first_bad_read;    /* smallest global index of a read the reader would reject, -1 = none                    */
    int64_t n_reads_local;     /* reads this process put in                                                             */
    int64_t rows_received;     /* rows this process's GPUs own after the exchange                                       */
    int64_t n_unique_local;    /* uniques this process's GPUs own                                                       */
    int64_t bytes_sent;        /* payload bytes sent to other ranks (NVLink)                                            */
    float   ms[5];             /* route, exchange, dedup, gather, order — CUDA events on the root's (else first) local GPU */
    float   reserved;
} fxg_dcollapse_report;
typedef struct fxg_dcollapse fxg_dcollapse;
int         fxg_dcollapse_new(fxg_comm *comm, int32_t stride, fxg_dcollapse **out);
void        fxg_dcollapse_free(fxg_dcollapse *d);
int         fxg_dcollapse_run(fxg_dcollapse *d, const fxg_batch *batches, const int64_t *index_base, const int32_t *const *weight_dev,
                              const int64_t *const *first_dev, int root, fxg_dcollapse_report *rep);
int         fxg_dcollapse_fetch_local(fxg_dcollapse *d, int local_index, uint8_t *out_seq, int32_t *out_len, uint64_t *out_count,
                                      int64_t *out_first, uint64_t *out_hash);
int         fxg_dcollapse_fetch_order(fxg_dcollapse *d, int32_t *perm_owner_host, uint32_t *perm_index_host, int64_t *ordered_first_host,
                                      uint64_t *ordered_count_host);
const char *fxg_dcollapse_error(const fxg_dcollapse *d);
int64_t     fxg_dcollapse_launches(const fxg_dcollapse *d);

/* ---- (f-2) three more loop bodies on the same slabs ------------------------------------------------------------------
 * fxg_validate_*: the reader's checks alone (fastx.c:45-54,118-135,361-362) — all fastx_trimmer needs, its body being
 *                 pointer arithmetic (src/fastx_trimmer/fastx_trimmer.c:120-148).  Result: report.first_bad_read.
 * fxg_mask_*:     fastq_masker body (src/fastq_masker/fastq_masker.c:92-107): bases with q < min_quality become
 *                 mask_char; masked_flag[i] = 1 iff read i had one; report.n_out = masked reads, aux[0] = masked bases.
 * fxg_artifacts_*: fastx_artifacts_filter decision (src/fastx_artifacts_filter/fastx_artifacts_filter.c:56-114):
 *                 keep[i] = 0 iff one of A,C,G,T makes up at least len-3 bases; report.n_out = kept reads. */
int fxg_validate_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, int64_t index_base);
int fxg_validate_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, fxg_report *report);
int fxg_mask_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int mask_char, uint8_t *out_seq_dev,
                  uint8_t *masked_flag_dev, int64_t index_base);
int fxg_mask_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int mask_char, uint8_t *out_seq_host,
                  uint8_t *masked_flag_host, fxg_report *report);
int fxg_artifacts_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *keep_dev, int64_t index_base);
int fxg_artifacts_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *keep_host, fxg_report *report);

/* ---- (f-4) fastq_to_fasta loop body (src/fastq_to_fasta/fastq_to_fasta.c:79-82): `strchr(fastx.nucleotides,'N') != NULL`
 * with the reader's checks fused (fastx.c:45-54,118-135).  has_n[i] = 1 iff read i holds an 'N'; report.n_out = such reads.
 * Dropping the quality line and renaming (-r, fastq_to_fasta.c:84-85) are writer-side and stay on the host. */
int fxg_has_n_dev (fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *has_n_dev, int64_t index_base);
int fxg_has_n_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *has_n_host, fxg_report *report);

/* ---- (f-4) barcode splitter matching loop (scripts/fastx_barcode_splitter.pl:208-290, mismatch_count :296) ---------
 * `fragments` is an ordinary batch whose rows are the read fragments the script compares (the first --bol / last --eol
 * barcode_len characters of each sequence line; shorter reads give shorter fragments, an empty line length 0): seq =
 * fragment bytes, len = fragment lengths (required), stride = 16/32/48/64.  The table holds the script's @barcodes list in
 * order: every barcode followed by its --partial forms (:170-176), rows of the same stride, zero padded.
 * best[i] = index of the FIRST entry with the lowest mismatch count if that count <= allowed_mismatches, else -1
 * ('unmatched'); the count is length(fragment) - equal positions + (barcode_len - entry length), exactly as the Perl. */
typedef struct {
    const uint8_t *entries;       /* host memory, n_entries x stride */
    const int32_t *entry_len;     /* host memory */
    int32_t n_entries;
    int32_t barcode_len;          /* length of the full barcodes */
    int32_t allowed_mismatches;   /* --mismatches (0 with --exact) */
} fxg_barcode_table;
int fxg_barcode_dev (fxg_ctx *ctx, const fxg_batch *fragments, const fxg_barcode_table *t, int32_t *best_dev);
int fxg_barcode_host(fxg_ctx *ctx, const fxg_batch *fragments, const fxg_barcode_table *t, int32_t *best_host, fxg_report *report);

/* ---- (f-3) fused pipelines (parity-checked on B200 against the composed oracle and the reference shell pipe) ----------
 * The tools chained on the device, e.g. fastq_quality_trimmer | fastx_clipper | fastq_quality_filter | fastx_collapser:
 * every stage runs the tool's own kernel on the survivors of the stage before (compacted in HBM, exactly the records the
 * next process of a shell pipe would read).  The map-type tools only shorten reads at the 3' end, so their result is one
 * length per ORIGINAL read: final_len[i] = length after the last stage, -1 = dropped by some stage.
 * FXG_STAGE_CLIP on mixed lengths (after a trimming stage, or on a ragged batch) reproduces the reference aligner's
 * grow-only query buffer (src/libfastx/sequence_alignment.cpp:131-153) with one scan over the
 * survivors; that path needs stride <= 160.  The buffer's history starts with the batch: one call = one input stream
 * (a second call does not see the first call's reads, exactly like a second run of the shell pipe).
 * FXG_STAGE_COLLAPSE must be the last stage: the survivors are added, in input order, to stage.collapser
 * (fxg_collapse_new; finish/fetch it afterwards); several calls may feed one collapser.
 * Blocking, with ONE device-to-host read at the end of the call (the survivor counts of all stages; they stay on the device
 * between the stages); report.first_bad_read refers to the input batch. */
enum { FXG_STAGE_TRIM = 0, FXG_STAGE_FILTER = 1, FXG_STAGE_CLIP = 2, FXG_STAGE_COLLAPSE = 3 };
typedef struct {
    int32_t op;                   /* FXG_STAGE_*                                                            */
    int32_t a0, a1;               /* TRIM: -t, -l    FILTER: -q, -p                                         */
    const fxg_clip_opts *clip;    /* CLIP                                                                   */
    fxg_collapser *collapser;     /* COLLAPSE                                                               */
} fxg_stage;
int fxg_pipeline_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, const fxg_stage *stages, int n_stages, int32_t *final_len_dev,
                     int64_t *n_survivors);

/* ---- next to the loop: FASTQ text in, FASTQ text out, parsed / packed / emitted on the GPU --------
 * fxg_text_run_host(): `text_host` holds raw 4-line FASTQ (any number of bytes; an incomplete trailing record is left
 * alone, see consumed_bytes).  The GPU indexes the lines (fastx.c:324-378 fgets/chomp), checks the record structure
 * (fastx.c:331-347,361-362,382-390), packs the slabs, runs op 0 = fastq_quality_trimmer (a0 = -t, a1 = -l),
 * op 1 = fastq_quality_filter (a0 = -q, a1 = -p) or op 2 = fastx_reverse_complement with fused validation, and writes
 * the surviving records as text (fastx.c:440-473) into out_host (capacity >= 1.25 x max_chunk_bytes).
 * fxg_text_stats_host() does the same up to the slabs and accumulates the quality-stats histogram (fxg_stats_accum_*)
 * into hist_dev instead of emitting text.  Anything it cannot reproduce bit-exactly
 * by construction (broken structure, a chunk mixing ASCII and numeric quality lines, malformed numbers, illegal bytes,
 * over-long lines) is reported as anomaly != 0 with nothing emitted, so the caller can re-read that chunk with the host
 * parser. */
typedef struct {
    int64_t n_records;        /* complete records found in the chunk                          */
    int64_t n_out_records;    /* records emitted                                              */
    int64_t consumed_bytes;   /* bytes of input covered by those records                      */
    int64_t out_bytes;        /* bytes written to out_host                                    */
    int32_t max_len;          /* longest read                                                 */
    int32_t anomaly;          /* 0 = none, else FXG_TEXT_* class                              */
    int64_t anomaly_record;   /* first record (0-based within the chunk) with the anomaly     */
    int32_t min_len;          /* shortest read                                                */
    int32_t reserved;
    int64_t clip_class[6];    /* fxg_text_clip_host: reads per FXG_CLIP_* class               */
    int64_t n_reads;          /* sum of get_reads_count() over the records (FASTA "N-COUNT" ids; = n_records for FASTQ) */
    int64_t n_out_reads;      /* the same over the emitted records                            */
    int64_t raw_out_bytes;    /* bytes of emitted TEXT (= out_bytes unless the chunk left the GPU deflated) */
    uint32_t out_crc32_pure;  /* deflated chunks: CRC-32 register of that text (init 0, no final xor)       */
    uint32_t deflated;        /* 1: out_host holds byte-aligned DEFLATE blocks                               */
} fxg_text_report;
#define FXG_TEXT_PREFIX     1
#define FXG_TEXT_EMPTY_SEQ  2
#define FXG_TEXT_QUAL_LEN   3
#define FXG_TEXT_LONG_LINE  4
#define FXG_TEXT_BAD_RECORD 5
#define FXG_TEXT_MIXED_LEN  7   /* clipper: read lengths differ (the reference then reads stale bytes: host path) */
typedef struct fxg_text fxg_text;
int         fxg_text_new(fxg_ctx *ctx, int device, size_t max_chunk_bytes, fxg_text **out);
void        fxg_text_free(fxg_text *t);
int         fxg_text_run_host(fxg_text *t, int op, const char *text_host, size_t bytes, int q_offset, int a0, int a1,
                              char *out_host, fxg_text_report *rep);
/* The trimmer / filter decision alone: out_len_host[r] = surviving length of record r or -1; line_start_host (may be NULL)
 * receives the byte offset of each of the record's 4 lines inside the chunk.  For a caller that keeps the input text and
 * writes the output from it: 4 to 20 bytes per record come back instead of the whole text. */
int         fxg_text_decide_host(fxg_text *t, int op, const char *text_host, size_t bytes, int q_offset, int a0, int a1,
                                 int32_t *out_len_host, uint32_t *line_start_host, fxg_text_report *rep);
/* fastx_clipper on a chunk of equal-length reads (expect_len = the length of every earlier read, 0 = none yet); other
 * chunks come back as FXG_TEXT_MIXED_LEN because the reference's aligner then depends on earlier reads' bytes. */
int         fxg_text_clip_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, const fxg_clip_opts *o,
                               int show_adapter_only, int expect_len, char *out_host, fxg_text_report *rep);
int         fxg_text_stats_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, uint64_t *hist_dev,
                                int32_t max_cycles, fxg_text_report *rep);
/* Record format of the input text: 0 = 4-line FASTQ (default), 1 = 2-line FASTA (fastx.c:348-352; ops 2 = reverse
 * complement, clip, stats, collapse; the read weights are get_reads_count() of the identifiers, fastx.c:475-497).
 * FASTQ chunks whose records all carry NUMERIC quality lines (fastx.c:137-167,382-390) are parsed on the GPU too and
 * written back in numeric form (fastx.c:421-438); a chunk that mixes the two forms is an anomaly (host parser). */
int         fxg_text_set_format(fxg_text *t, int fasta);
/* fastx_collapser: the reads of the chunk (FASTA or FASTQ; FASTQ records are validated as the reader validates them) are
 * added to `col` (fxg_collapse_new; finish / fetch it when the input is exhausted).  first_base = first-occurrence index
 * of the chunk's record 0 (any values that grow with the input order, e.g. chunk number << 32, so chunks may be added
 * in any order); < 0: the number of rows added so far. */
int         fxg_text_collapse_host(fxg_text *t, const char *text_host, size_t bytes, int q_offset, fxg_collapser *col, int64_t first_base,
                                   fxg_text_report *rep);
/* `-z` (src/libfastx/fastx.c:214-248 pipes the text through a forked gzip): with deflate on, the emitted text of every
 * chunk leaves the GPU as non-final, byte-aligned DEFLATE blocks (dynamic Huffman codes over the literals, one per 64 KB),
 * so chunks concatenate.  The caller frames them: 10-byte gzip header, the chunks, a final empty stored block
 * (01 00 00 FF FF), then CRC-32 and ISIZE.  fxg_crc32_concat() chains the chunks' pure CRCs, fxg_crc32_finish() turns the
 * result into the CRC-32 gzip stores.  Any gunzip yields exactly the text the plain path emits. */
int         fxg_text_set_deflate(fxg_text *t, int on);
uint32_t    fxg_crc32_concat(uint32_t crc_pure_a, uint32_t crc_pure_b, uint64_t len_b);
uint32_t    fxg_crc32_finish(uint32_t crc_pure, uint64_t total_len);
const char *fxg_text_error(const fxg_text *t);
int64_t     fxg_text_launches(const fxg_text *t);
int64_t     fxg_text_numeric_chunks(const fxg_text *t);   /* chunks that went through the numeric-quality form of the path */
int64_t     fxg_text_fasta_chunks(const fxg_text *t);     /* chunks that went through the FASTA form of the path          */

#ifdef __cplusplus
}
#endif
#endif /* FXG_H */
