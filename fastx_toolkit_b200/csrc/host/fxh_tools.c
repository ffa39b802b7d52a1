/*
 * fxh_tools.c — the drop-in executables as one multi-call program (dispatch on basename(argv[0])):
 *   fastq_quality_trimmer  fastq_quality_filter  fastx_reverse_complement
 *   fastx_clipper          fastx_collapser       fastx_quality_stats
 *   + the SURVEY §8(f-2) rows:  fastx_trimmer  fastq_masker  fastx_artifacts_filter
 * Same flags, streams, messages and exit status as FASTX-Toolkit 0.0.14; the per-read loop bodies run on the
 * GPU through include/fxg.h (no CPU fallback).  Reference mains:
 *   src/fastq_quality_trimmer/fastq_quality_trimmer.c:52-123   src/fastq_quality_filter/fastq_quality_filter.c:54-178
 *   src/fastx_reverse_complement/fastx_reverse_complement.c:106-128   src/fastx_clipper/fastx_clipper.cpp:90-350
 *   src/fastx_collapser/fastx_collapser.cpp:93-138   src/fastx_quality_stats/fastx_quality_stats.c:420-463
 */
#define _GNU_SOURCE
#include <err.h>
#include <errno.h>
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fxh.h"
#include "fxh_stream.h"
#include "fxh_usage.h"

static void *pinned(size_t bytes)
{
    void *p = fxg_alloc_pinned(bytes);
    if (!p) errx(1, "cannot allocate %zu bytes of pinned host memory", bytes);
    return p;
}

/* batch-sized pinned result arrays, grown on demand */
typedef struct { void *p; size_t cap; } pbuf;
static void *pbuf_get(pbuf *b, size_t bytes)
{
    if (bytes > b->cap) { fxg_free_pinned(b->p); b->p = pinned(bytes); b->cap = bytes; }
    return b->p;
}

static int batch_q(const fxh_batch *b) { return b->numeric_qual ? 33 : fxh_q_offset(); }

/* GPU text path (fxg_text_*) driven by the streaming engine (fxh_stream.c): whole chunks of raw FASTQ / FASTA text go to
 * the GPUs (FASTX_GPUS of them), which parse, pack, run the op and emit the output text; the host only moves blocks, with
 * read(2), the copies, the kernels and write(2) overlapped.  The engine stops at the first chunk holding anything the GPU
 * path does not reproduce bit-exactly by construction (broken structure, illegal bytes, mixed quality encodings) and
 * leaves the reader positioned there, so the record-by-record host path that follows produces the reference's output
 * and message. */
static int first_gpu(void) { return fxh_first_device(); }

static void text_fast_path(fxh_reader *rd, fxh_writer *wr, int op, int a0, int a1)
{
    fxs_job job;
    memset(&job, 0, sizeof job);
    job.op = op; job.a0 = a0; job.a1 = a1;
    job.ngpu = fxh_gpu_count(); job.first_dev = first_gpu();
    (void)fxs_run(&job, rd, wr);
}

/* ================================================================================ fastq_quality_trimmer */
static int tr_min_quality = 0, tr_min_length = 0;

static int tr_args(int oi, int optc, char *oa)
{
    (void)oi;
    switch (optc) {
    case 'l':
        if (oa == NULL) errx(1, "[-l] parameter requires an argument value");
        tr_min_length = (int)strtoul(oa, NULL, 10);
        if (tr_min_length < 0) errx(1, "Invalid minimum length value (-l %s)", oa);
        break;
    case 't':
        if (oa == NULL) errx(1, "[-t] parameter requires an argument value");
        tr_min_quality = (int)strtol(oa, NULL, 10);
        break;
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

static int main_trimmer(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "t:l:", tr_args, fxh_usage_fastq_quality_trimmer);
    if (tr_min_quality == 0) errx(1, "Missing minimum quality threshold value (-t)");
    const double t_start = fxh_now();
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTQ_ONLY, fxh_q_offset(), 0);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), 1, fxh_compress_output());
    const double t_opened = fxh_now();
    fxg_ctx *ctx = fxh_gpu_open();
    if (getenv("FASTX_TIMING")) fprintf(stderr, "[timing] open+first read %.3f s, GPU context %.3f s\n", t_opened - t_start, fxh_now() - t_opened);
    pbuf out = { 0, 0 };
    fxh_batch *b;
    text_fast_path(rd, wr, FXS_TRIM, tr_min_quality, tr_min_length);
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        int32_t *out_len = (int32_t *)pbuf_get(&out, (size_t)b->n * sizeof(int32_t));
        fxg_batch gb = fxh_as_fxg_batch(b, 1);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_trim_host(ctx, &gb, batch_q(b), tr_min_quality, tr_min_length, out_len, &rep), "fxg_trim_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++)
            if (out_len[i] >= 0)
                fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride, b->qual + (size_t)i * b->stride, out_len[i]);
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (getenv("FASTX_TIMING")) fprintf(stderr, "[timing] total %.3f s\n", fxh_now() - t_start);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Minimum Quality Threshold: %d\n", tr_min_quality);
        if (tr_min_length > 0) fprintf(f, "Minimum Length: %d\n", tr_min_length);
        else fprintf(f, "No minimum Length\n");
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
        size_t discarded = fxh_num_input_reads(rd) - fxh_num_output_reads(wr);
        fprintf(f, "discarded %zu (%zu%%) too-short reads.\n", discarded, (discarded * 100) / fxh_num_input_reads(rd));
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastq_quality_filter */
static int fl_min_quality = 0, fl_min_percent = 0;

static int fl_args(int oi, int optc, char *oa)
{
    (void)oi;
    switch (optc) {
    case 'q':
        if (oa == NULL) errx(1, "[-q] parameter requires an argument value");
        fl_min_quality = (int)strtoul(oa, NULL, 10);
        break;
    case 'p':
        if (oa == NULL) errx(1, "[-l] parameter requires an argument value");
        fl_min_percent = (int)strtoul(oa, NULL, 10);
        if (fl_min_percent <= 0 || fl_min_percent > 100) errx(1, "Invalid percent value (-p %s)", oa);
        break;
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

static int main_filter(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "q:p:", fl_args, fxh_usage_fastq_quality_filter);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTQ_ONLY, fxh_q_offset(), 0);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), 1, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    pbuf out = { 0, 0 };
    fxh_batch *b;
    text_fast_path(rd, wr, FXS_FILTER, fl_min_quality, fl_min_percent);
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        uint8_t *keep = (uint8_t *)pbuf_get(&out, (size_t)b->n);
        fxg_batch gb = fxh_as_fxg_batch(b, 1);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_filter_host(ctx, &gb, batch_q(b), fl_min_quality, fl_min_percent, keep, &rep), "fxg_filter_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++)
            if (keep[i])
                fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride, b->qual + (size_t)i * b->stride, b->len[i]);
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Quality cut-off: %d\n", fl_min_quality);
        fprintf(f, "Minimum percentage: %d\n", fl_min_percent);
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
        size_t discarded = fxh_num_input_reads(rd) - fxh_num_output_reads(wr);
        fprintf(f, "discarded %zu (%zu%%) low-quality reads.\n", discarded, (discarded * 100) / fxh_num_input_reads(rd));
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastx_reverse_complement */
static int main_revcomp(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "", NULL, fxh_usage_fastx_reverse_complement);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    pbuf os = { 0, 0 }, oq = { 0, 0 };
    fxh_batch *b;
    text_fast_path(rd, wr, FXS_REVCOMP, 0, 0);
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        const size_t bytes = (size_t)b->n * b->stride;
        uint8_t *oseq = (uint8_t *)pbuf_get(&os, bytes), *oqual = fastq ? (uint8_t *)pbuf_get(&oq, bytes) : NULL;
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_revcomp_host(ctx, &gb, batch_q(b), oseq, oqual, &rep), "fxg_revcomp_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++)
            fxh_write_record(wr, b, i, oseq + (size_t)i * b->stride, fastq ? oqual + (size_t)i * b->stride : NULL, b->len[i]);
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Printing Reverse-Complement Sequences.\n");
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastx_clipper */
static char cl_adapter[FXG_MAX_ADAPTER] = "CCTTAAGG";
static unsigned int cl_min_length = 5;
static int cl_discard_unknown = 1, cl_keep_delta = 0, cl_discard_non_clipped = 0, cl_discard_clipped = 0;
static int cl_show_adapter_only = 0, cl_debug = 0, cl_min_adapter_len = 0;

static int cl_args(int oi, int optc, char *oa)
{
    (void)oi;
    switch (optc) {
    case 'M':
        if (oa == NULL) errx(1, "[-M] parameter requires an argument value");
        cl_min_adapter_len = atoi(oa);
        if (cl_min_adapter_len <= 0) errx(1, "Invalid minimum adapter length (-M %s)", oa);
        break;
    case 'k': cl_show_adapter_only = 1; break;
    case 'D': cl_debug++; break;
    case 'c': cl_discard_non_clipped = 1; break;
    case 'C': cl_discard_clipped = 1; break;
    case 'd':
        if (oa == NULL) errx(1, "[-d] parameter requires an argument value");
        cl_keep_delta = (int)strtoul(oa, NULL, 10);
        if (cl_keep_delta < 0) errx(1, "Invalid number bases to keep (-d %s)", oa);
        break;
    case 'a': strncpy(cl_adapter, oa, sizeof(cl_adapter) - 1); break;
    case 'l':
        if (oa == NULL) errx(1, "[-l] parameter requires an argument value");
        cl_min_length = (unsigned int)strtoul(oa, NULL, 10);
        break;
    case 'n': cl_discard_unknown = 0; break;
    case 's': break;        /* accepted and ignored, like the reference's getopt string (fastx_clipper.cpp:151) */
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

/* fastx_clipper on the GPU text path: chunks whose reads all have one and the same length (the usual raw-reads case).
 * The first chunk with another length hands over to the record path, whose packer reproduces the reference aligner's
 * stale query buffer — the engine seeds it with the last read it consumed. */
static void clip_text_path(fxh_reader *rd, fxh_writer *wr, const fxg_clip_opts *o, unsigned int *count_input, unsigned int *cnt)
{
    fxs_job job;
    memset(&job, 0, sizeof job);
    job.op = FXS_CLIP; job.a0 = cl_show_adapter_only; job.clip = o;
    job.ngpu = fxh_gpu_count(); job.first_dev = first_gpu();
    (void)fxs_run(&job, rd, wr);
    *count_input += (unsigned int)job.reads;          /* get_reads_count() per record (fastx_clipper.cpp:259,277) */
    for (int k = 0; k < 6; k++) cnt[k] += job.clip_class[k];
}

static int main_clipper(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "M:kDCcd:a:s:l:n", cl_args, fxh_usage_fastx_clipper);
    if (cl_keep_delta > 0) cl_keep_delta += (int)strlen(cl_adapter);
    if (cl_debug) warnx("[-D] alignment dumps are not produced by the GPU build; the flag is ignored");
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 1 /* aligner's stale rows */);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    fxg_clip_opts o;
    o.adapter = cl_adapter; o.min_length = (int32_t)cl_min_length; o.keep_delta = cl_keep_delta;
    o.discard_non_clipped = cl_discard_non_clipped; o.discard_clipped = cl_discard_clipped;
    o.discard_unknown = cl_discard_unknown; o.min_adapter_len = cl_min_adapter_len;
    unsigned int count_input = 0, cnt[6] = { 0, 0, 0, 0, 0, 0 };
    pbuf ol = { 0, 0 }, oc = { 0, 0 };
    fxh_batch *b;
    clip_text_path(rd, wr, &o, &count_input, cnt);
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        int32_t *out_len = (int32_t *)pbuf_get(&ol, (size_t)b->n * sizeof(int32_t));
        uint8_t *out_cls = (uint8_t *)pbuf_get(&oc, (size_t)b->n);
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_clip_host(ctx, &gb, b->width, batch_q(b), &o, out_len, out_cls, &rep), "fxg_clip_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++) {
            const uint8_t *s = b->seq + (size_t)i * b->stride, *q = b->qual + (size_t)i * b->stride;
            count_input += (unsigned int)b->weight[i];
            cnt[out_cls[i]] += (unsigned int)b->weight[i];
            if (out_cls[i] == FXG_CLIP_ADAPTER_ONLY) {
                if (cl_show_adapter_only) fxh_write_record(wr, b, i, s, q, b->len[i]);      /* untruncated */
            } else if (out_cls[i] == FXG_CLIP_WRITE && !cl_show_adapter_only)
                fxh_write_record(wr, b, i, s, q, out_len[i]);
        }
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Clipping Adapter: %s\n", cl_adapter);
        fprintf(f, "Min. Length: %d\n", cl_min_length);
        if (cl_discard_clipped) fprintf(f, "Clipped reads - discarded.\n");
        if (cl_discard_non_clipped) fprintf(f, "Non-Clipped reads - discarded.\n");
        fprintf(f, "Input: %u reads.\n", count_input);
        fprintf(f, "Output: %u reads.\n", count_input - cnt[FXG_CLIP_TOO_SHORT] - cnt[FXG_CLIP_NON_CLIPPED] - cnt[FXG_CLIP_CLIPPED] -
                                              cnt[FXG_CLIP_HAS_N] - cnt[FXG_CLIP_ADAPTER_ONLY]);
        fprintf(f, "discarded %u too-short reads.\n", cnt[FXG_CLIP_TOO_SHORT]);
        fprintf(f, "discarded %u adapter-only reads.\n", cnt[FXG_CLIP_ADAPTER_ONLY]);
        if (cl_discard_non_clipped) fprintf(f, "discarded %u non-clipped reads.\n", cnt[FXG_CLIP_NON_CLIPPED]);
        if (cl_discard_clipped) fprintf(f, "discarded %u clipped reads.\n", cnt[FXG_CLIP_CLIPPED]);
        if (cl_discard_unknown) fprintf(f, "discarded %u N reads.\n", cnt[FXG_CLIP_HAS_N]);
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastx_collapser */
static int main_collapser(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "", NULL, fxh_usage_fastx_collapser);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    FILE *out = stdout;
    if (strcmp(fxh_output_filename(), "-") != 0) {
        out = fopen(fxh_output_filename(), "w");
        if (!out) errx(1, "Failed to create output file (%s)", fxh_output_filename());
    }
    /* the count map lives on the GPU(s) from the first read on (fastx_collapser.cpp:112-114) and grows as the input arrives.
     * FASTX_GPUS=N: every GPU dedups the chunks it gets; afterwards the partial maps are merged by owner = std::hash mod N
     * over NVLink and the output order is computed once (fxg_dcollapse_*). */
    const int ngpu = fxh_gpu_count(), dev0 = first_gpu();
    fxg_ctx *ctxs[64];
    fxg_collapser *cols[64];
    int devs[64];
    for (int g = 0; g < ngpu; g++) {
        devs[g] = dev0 + g;
        ctxs[g] = fxh_gpu_open_dev(devs[g]);
        int rc0 = fxg_collapse_new(devs[g], 1 << 20, 64, &cols[g]);
        if (rc0 != FXG_OK) errx(1, "fxg_collapse_new failed: %s", fxg_strerror(rc0));
    }
    fxg_ctx *ctx = ctxs[0];
    fxg_collapser *col = cols[0];
    int rc;
    fxs_job job;
    memset(&job, 0, sizeof job);
    job.op = FXS_COLLAPSE; job.ngpu = ngpu; job.first_dev = dev0; job.collapsers = cols;
    (void)fxs_run(&job, rd, NULL);
    /* whatever the GPU text path handed back: record by record, validated as the reader validates (K-VALIDATE) */
    int64_t host_rows = 0;
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_validate_host(ctx, &gb, batch_q(b), &rep), "fxg_validate_host");
        if (rep.first_bad_read >= 0) fxh_die_bad_record(rd, b, rep.first_bad_read);
        gb.qual = NULL;
        rc = fxg_collapse_add(col, &gb, b->weight, NULL, ((job.chunks + 1) << 32) + host_rows);
        if (rc != FXG_OK) errx(1, "fxg_collapse_add failed: %s (%s)", fxg_strerror(rc), fxg_collapse_error(col));
        host_rows += b->n;
    }
    int64_t U = 0, bad = -1;
    int32_t stride = 0;
    uint8_t *useq = NULL; int32_t *ulen = NULL; uint64_t *ucnt = NULL;
    if (ngpu == 1) {
        rc = fxg_collapse_finish(col, 1, &U, &bad);
        if (rc != FXG_OK) errx(1, "fxg_collapse_finish failed: %s (%s)", fxg_strerror(rc), fxg_collapse_error(col));
        if (bad >= 0) errx(1, "internal error: the count map rejected a read the validation had accepted (index %lld)", (long long)bad);
        stride = fxg_collapse_stride(col);
        useq = (uint8_t *)malloc((size_t)(U > 0 ? U : 1) * (size_t)stride);
        ulen = (int32_t *)malloc((size_t)(U > 0 ? U : 1) * sizeof(int32_t));
        ucnt = (uint64_t *)malloc((size_t)(U > 0 ? U : 1) * sizeof(uint64_t));
        if (!useq || !ulen || !ucnt) err(1, "out of memory");
        rc = fxg_collapse_fetch(col, useq, ulen, ucnt, NULL, NULL);
        if (rc != FXG_OK) errx(1, "fxg_collapse_fetch failed: %s (%s)", fxg_strerror(rc), fxg_collapse_error(col));
    } else {
        /* the partial maps' uniques (rows, lengths, counts, first indices) stay in HBM and go through the owner exchange */
        int64_t ul[64];
        fxg_batch pb[64]; int64_t base[64];
        const int32_t *wdev[64]; const int64_t *fdev[64];
        for (int g = 0; g < ngpu; g++) {
            rc = fxg_collapse_finish(cols[g], 0, &ul[g], &bad);
            if (rc != FXG_OK) errx(1, "fxg_collapse_finish failed: %s (%s)", fxg_strerror(rc), fxg_collapse_error(cols[g]));
            if (fxg_collapse_stride(cols[g]) > stride) stride = fxg_collapse_stride(cols[g]);
        }
        for (int g = 0; g < ngpu; g++) {
            const size_t u = (size_t)(ul[g] > 0 ? ul[g] : 1);
            if ((rc = fxg_collapse_reserve(cols[g], 0, stride)) != FXG_OK) errx(1, "fxg_collapse_reserve failed: %s", fxg_collapse_error(cols[g]));
            uint8_t *d_rows = (uint8_t *)fxg_alloc_device(ctxs[g], u * (size_t)stride);
            int32_t *d_len = (int32_t *)fxg_alloc_device(ctxs[g], u * 4), *d_w = (int32_t *)fxg_alloc_device(ctxs[g], u * 4);
            uint64_t *d_cnt = (uint64_t *)fxg_alloc_device(ctxs[g], u * 8);
            int64_t *d_first = (int64_t *)fxg_alloc_device(ctxs[g], u * 8);
            if (!d_rows || !d_len || !d_w || !d_cnt || !d_first) errx(1, "out of device memory on GPU %d", devs[g]);
            rc = fxg_collapse_fetch(cols[g], d_rows, d_len, d_cnt, d_first, NULL);
            if (rc != FXG_OK) errx(1, "fxg_collapse_fetch failed: %s (%s)", fxg_strerror(rc), fxg_collapse_error(cols[g]));
            /* counts as the 32-bit weights the exchange carries */
            uint64_t *hc = (uint64_t *)malloc(u * 8); int32_t *hw = (int32_t *)malloc(u * 4);
            if (!hc || !hw) err(1, "out of memory");
            fxh_gpu_check(ctxs[g], fxg_memcpy_d2h(ctxs[g], hc, d_cnt, (size_t)ul[g] * 8), "fxg_memcpy_d2h");
            for (int64_t k = 0; k < ul[g]; k++) hw[k] = (int32_t)hc[k];
            fxh_gpu_check(ctxs[g], fxg_memcpy_h2d(ctxs[g], d_w, hw, (size_t)ul[g] * 4), "fxg_memcpy_h2d");
            free(hc); free(hw);
            fxg_collapse_free(cols[g]); cols[g] = NULL;
            pb[g].seq = d_rows; pb[g].qual = NULL; pb[g].len = d_len; pb[g].uniform_len = 0; pb[g].stride = stride; pb[g].n = ul[g];
            base[g] = 0; wdev[g] = d_w; fdev[g] = d_first;
        }
        fxg_comm *comm = NULL;
        if ((rc = fxg_comm_init_all(ngpu, devs, &comm)) != FXG_OK) errx(1, "fxg_comm_init_all failed: %s (%s)", fxg_strerror(rc), fxg_comm_error(NULL));
        fxg_dcollapse *dc = NULL;
        if ((rc = fxg_dcollapse_new(comm, stride, &dc)) != FXG_OK) errx(1, "fxg_dcollapse_new failed: %s", fxg_strerror(rc));
        fxg_dcollapse_report drep;
        if ((rc = fxg_dcollapse_run(dc, pb, base, wdev, fdev, 0, &drep)) != FXG_OK) errx(1, "fxg_dcollapse_run failed: %s (%s)", fxg_strerror(rc), fxg_dcollapse_error(dc));
        U = drep.n_unique;
        const size_t u = (size_t)(U > 0 ? U : 1);
        int32_t *powner = (int32_t *)malloc(u * 4); uint32_t *pidx = (uint32_t *)malloc(u * 4);
        ucnt = (uint64_t *)malloc(u * 8);
        useq = (uint8_t *)malloc(u * (size_t)stride);
        ulen = (int32_t *)malloc(u * sizeof(int32_t));
        if (!powner || !pidx || !ucnt || !useq || !ulen) err(1, "out of memory");
        if ((rc = fxg_dcollapse_fetch_order(dc, powner, pidx, NULL, ucnt)) != FXG_OK) errx(1, "fxg_dcollapse_fetch_order failed: %s", fxg_dcollapse_error(dc));
        /* every owner's rows come back over its own PCIe link; the output is gathered in the order the root computed */
        uint8_t *orow[64]; int32_t *olen[64];
        for (int g = 0; g < ngpu; g++) {
            orow[g] = (uint8_t *)malloc(u * (size_t)stride); olen[g] = (int32_t *)malloc(u * 4);
            if (!orow[g] || !olen[g]) err(1, "out of memory");
            if ((rc = fxg_dcollapse_fetch_local(dc, g, orow[g], olen[g], NULL, NULL, NULL)) != FXG_OK) errx(1, "fxg_dcollapse_fetch_local failed: %s", fxg_dcollapse_error(dc));
        }
        for (int64_t k = 0; k < U; k++) {
            memcpy(useq + (size_t)k * stride, orow[powner[k]] + (size_t)pidx[k] * stride, (size_t)stride);
            ulen[k] = olen[powner[k]][pidx[k]];
        }
        for (int g = 0; g < ngpu; g++) { free(orow[g]); free(olen[g]); }
        free(powner); free(pidx);
        fxg_dcollapse_free(dc);
        fxg_comm_free(comm);
    }
    static char obuf[1 << 22];
    setvbuf(out, obuf, _IOFBF, sizeof obuf);
    size_t total_reads = 0;
    for (int64_t k = 0; k < U; k++) {   /* PrintCollapsedSequence, fastx_collapser.cpp:80-85 (count narrowed to int) */
        const int c = (int)ucnt[k];
        total_reads += (size_t)c;
        fprintf(out, ">%zu-%d\n%.*s\n", (size_t)(k + 1), c, ulen[k], (const char *)useq + (size_t)k * stride);
    }
    if (out != stdout) fclose(out); else fflush(out);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Input: %zu sequences (representing %zu reads)\n", fxh_num_input_sequences(rd), fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu sequences (representing %zu reads)\n", (size_t)U, total_reads);
    }
    for (int g = 0; g < ngpu; g++) { if (cols[g]) fxg_collapse_free(cols[g]); fxg_destroy(ctxs[g]); }
    return 0;
}

/* ================================================================================ fastx_quality_stats */
static int st_new_format = 0;
static int st_args(int oi, int optc, char *oa)
{
    (void)oi; (void)oa;
    if (optc == 'N') st_new_format = 1;
    else errx(1, "Unknown argument (%c)", optc);
    return 1;
}

/* Everything the tool prints is derived from one (cycle, nucleotide) histogram h[q+15]
 * (fastx_quality_stats.c:218-247 get_nth_value, :276-417 printers). */
typedef struct { int count, min, max; long long sum; const uint64_t *h; int fasta; int cycle, nuc; } nucstat;

/* FASTA input has no quality bins, so the reference's get_nth_value() walks past the end of its
 * bases_values_count[] into the fields of the following table entries (min, max, count, padding, sum, bins ...).
 * The result is garbage but deterministic; to stay byte-identical we replay the walk over the same memory
 * image: struct nucleotide_data = { int min, max, count; unsigned long long sum; int bins[108]; } laid out under the
 * `#pragma pack(1)` that fastx.h:60 leaves active = 113 ints (no padding), 6 per cycle (fastx_quality_stats.c:115-133).  fa_counts[cycle*6+nuc] are the only non-constant fields. */
static const int *fa_counts = NULL;
static int fa_cycles = 0;
static int fa_int(long long idx)
{
    const long long e = idx / 113, f = idx % 113;
    if (f == 0) return 100;                       /* min  */
    if (f == 1) return -100;                      /* max  */
    if (f == 2) return (e / 6 < fa_cycles) ? fa_counts[e] : 0;   /* count */
    return 0;                                     /* sum (never updated for FASTA), bins */
}

static int st_nth(const nucstat *s, int n)
{
    if (n == 0) return s->min;
    if (s->fasta) {
        const long long base = ((long long)s->cycle * 6 + s->nuc) * 113 + 5;
        /* Beyond the cycles that hold data every entry is {100,-100,0,...}: the reference keeps walking (its
         * result then depends on whatever follows its static table); we stop at the end of the data. */
        const long long end = (long long)fa_cycles * 6 * 113;
        long long pos = 0;
        while (n > 0 && base + pos < end) {
            if (fa_int(base + pos) > n) break;
            n -= fa_int(base + pos);
            pos++;
            while (base + pos < end && fa_int(base + pos) == 0) pos++;
        }
        return (int)(pos - 15);
    }
    int pos = 0;
    while (n > 0) {
        if ((long long)s->h[pos] > n) break;
        n -= (int)s->h[pos];
        pos++;
        while (pos < FXG_QBINS && s->h[pos] == 0) pos++;
    }
    return pos - 15;
}

static void st_fill(nucstat *s, const uint64_t *h, int fasta, int cycle, int nuc)
{
    s->h = h; s->fasta = fasta; s->count = 0; s->min = 100; s->max = -100; s->sum = 0; s->cycle = cycle; s->nuc = nuc;
    for (int b = 0; b < FXG_QBINS; b++) {
        if (!h[b]) continue;
        s->count += (int)h[b];
        if (!fasta) {
            if (s->min == 100) s->min = b - 15;
            s->max = b - 15;
            s->sum += (long long)(b - 15) * (long long)h[b];
        }
    }
}

static void st_print_fields(FILE *f, const nucstat *s, const char *lead)
{
    const int Q1 = st_nth(s, s->count / 4), Q3 = st_nth(s, s->count * 3 / 4), IQR = Q3 - Q1;
    const int lw = ((Q1 - IQR * 3 / 2) < s->min) ? s->min : (Q1 - IQR * 3 / 2);
    const int rw = ((Q3 + IQR * 3 / 2) > s->max) ? s->max : (Q3 + IQR * 3 / 2);
    /* the reference keeps `sum` in an unsigned long long: a negative total becomes a huge mean (kept); the
     * division happens at run time so that 0/0 prints "-nan" like the reference binary */
    volatile double num = (double)(unsigned long long)s->sum, den = (double)s->count;
    fprintf(f, "%s%d\t%d\t%d\t%lld\t", lead, s->count, s->min, s->max, s->sum);
    fprintf(f, "%3.2f\t%d\t%d\t%d\t", num / den, Q1, st_nth(s, s->count / 2), Q3);
    fprintf(f, "%d\t%d\t%d", IQR, lw, rw);
}

static int main_stats(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "N", st_args, fxh_usage_fastx_quality_stats);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    FILE *out = stdout;
    if (strcmp(fxh_output_filename(), "-") != 0) {
        out = fopen(fxh_output_filename(), "w+");
        if (out == NULL) err(1, "Failed to create output file (%s)", fxh_output_filename());
    }
    /* FASTX_GPUS=N: chunks go to the GPUs round-robin, each GPU keeps a partial histogram, and one NCCL all-reduce
     * (fxg_comm_allreduce_u64) merges them before printing — the only collective this tool needs. */
    const int ngpu = fxh_gpu_count();
    const int dev0 = first_gpu();
    fxg_ctx *ctxs[64]; uint64_t *hists[64]; int devs[64];
    const int max_cycles = FXH_MAX_LINE;
    const size_t hist_bytes = (size_t)max_cycles * 5 * FXG_QBINS * sizeof(uint64_t);
    for (int g = 0; g < ngpu; g++) {
        devs[g] = dev0 + g;
        ctxs[g] = fxh_gpu_open_dev(devs[g]);
        hists[g] = (uint64_t *)fxg_alloc_device(ctxs[g], hist_bytes);
        if (!hists[g]) errx(1, "cannot allocate the histogram on GPU %d: %s", devs[g], fxg_last_error(ctxs[g]));
        fxh_gpu_check(ctxs[g], fxg_memset_dev(ctxs[g], hists[g], 0, hist_bytes), "fxg_memset_dev");
        fxh_gpu_check(ctxs[g], fxg_sync(ctxs[g]), "fxg_sync");
    }
    fxg_ctx *ctx = ctxs[0];
    uint64_t *d_hist = hists[0];
    int maxlen = 0, turn = 0;
    fxh_batch *b;
    {   /* GPU text path: parse + pack + accumulate on the devices, chunks to whichever GPU is free */
        fxs_job job;
        memset(&job, 0, sizeof job);
        job.op = FXS_STATS; job.ngpu = ngpu; job.first_dev = dev0; job.hist_dev = hists; job.max_cycles = max_cycles;
        (void)fxs_run(&job, rd, NULL);
        if (job.max_len > maxlen) maxlen = job.max_len;
    }
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        const int g = turn++ % ngpu;
        fxh_gpu_check(ctxs[g], fxg_stats_accum_host(ctxs[g], &gb, batch_q(b), hists[g], max_cycles, fastq ? NULL : b->weight, &rep), "fxg_stats_accum_host");
        if (rep.first_bad_read >= 0) fxh_die_bad_record(rd, b, rep.first_bad_read);
        for (int64_t i = 0; i < b->n; i++) if (b->len[i] > maxlen) maxlen = b->len[i];
    }
    if (ngpu > 1 && maxlen > 0) {
        fxg_comm *comm = NULL;
        int rc = fxg_comm_init_all(ngpu, devs, &comm);
        if (rc != FXG_OK) errx(1, "fxg_comm_init_all failed: %s (%s)", fxg_strerror(rc), fxg_comm_error(NULL));
        rc = fxg_comm_allreduce_u64(comm, hists, (size_t)maxlen * 5 * FXG_QBINS);
        if (rc != FXG_OK) errx(1, "fxg_comm_allreduce_u64 failed: %s (%s)", fxg_strerror(rc), fxg_comm_error(comm));
        fxg_comm_free(comm);
    }
    const size_t cyc_words = (size_t)5 * FXG_QBINS;
    uint64_t *hist = (uint64_t *)calloc((size_t)(maxlen + 1) * cyc_words, sizeof(uint64_t));
    if (!hist) err(1, "out of memory");
    if (maxlen > 0) fxh_gpu_check(ctx, fxg_memcpy_d2h(ctx, hist, d_hist, (size_t)maxlen * cyc_words * sizeof(uint64_t)), "fxg_memcpy_d2h");

    static const char *names[6] = { "ALL", "A", "C", "G", "T", "N" };
    static const char *cols[11] = { "count", "min", "max", "sum", "mean", "Q1", "med", "Q3", "IQR", "lW", "rW" };
    uint64_t all[FXG_QBINS];
    nucstat s[6];
    int max_count = 0;
    if (!fastq) {   /* table of counts for the FASTA walk emulation */
        int *fc = (int *)calloc((size_t)(maxlen + 1) * 6, sizeof(int));
        if (!fc) err(1, "out of memory");
        for (int c = 0; c < maxlen; c++)
            for (int nuc = 0; nuc < 5; nuc++) {
                const int v = (int)hist[(size_t)c * cyc_words + (size_t)nuc * FXG_QBINS + 15];
                fc[c * 6 + 1 + nuc] = v;
                fc[c * 6] += v;
            }
        fa_counts = fc; fa_cycles = maxlen;
    }
    if (st_new_format) {
        fprintf(out, "cycle\tmax_count");
        for (int nuc = 0; nuc < 6; nuc++) for (int k = 0; k < 11; k++) fprintf(out, "\t%s_%s", names[nuc], cols[k]);
        fprintf(out, "\n");
    } else {
        fprintf(out, "column\tcount\tmin\tmax\tsum\tmean\tQ1\tmed\tQ3\tIQR\tlW\trW\tA_Count\tC_Count\tG_Count\tT_Count\tN_Count\tMax_count\n");
    }
    for (int c = 0; c < maxlen; c++) {
        const uint64_t *hc = hist + (size_t)c * cyc_words;
        for (int q = 0; q < FXG_QBINS; q++) all[q] = hc[q] + hc[FXG_QBINS + q] + hc[2 * FXG_QBINS + q] + hc[3 * FXG_QBINS + q] + hc[4 * FXG_QBINS + q];
        st_fill(&s[0], all, !fastq, c, 0);
        for (int nuc = 0; nuc < 5; nuc++) st_fill(&s[1 + nuc], hc + (size_t)nuc * FXG_QBINS, !fastq, c, 1 + nuc);
        if (s[0].count == 0) break;
        if (c == 0) max_count = s[0].count;
        if (st_new_format) {
            fprintf(out, "%d\t%d", c + 1, max_count);
            for (int nuc = 0; nuc < 6; nuc++) st_print_fields(out, &s[nuc], "\t");
            fprintf(out, "\n");
        } else {
            fprintf(out, "%d\t", c + 1);
            st_print_fields(out, &s[0], "");
            fprintf(out, "\t%d\t%d\t%d\t%d\t%d\t%d\n", s[1].count, s[2].count, s[3].count, s[4].count, s[5].count, max_count);
        }
    }
    if (out != stdout) fclose(out); else fflush(out);
    for (int g = 0; g < ngpu; g++) { fxg_free_device(ctxs[g], hists[g]); fxg_destroy(ctxs[g]); }
    return 0;
}

/* ================================================================================ fastx_trimmer   */
static int ft_first = 1, ft_last = 0, ft_by_position = 0, ft_from_end = 0;
static unsigned int ft_trim_last = 0, ft_min_len = 0;

static int ft_args(int oi, int optc, char *oa)
{
    (void)oi;
    switch (optc) {
    case 'f':
        if (oa == NULL) errx(1, "[-f] parameter requires an argument value");
        ft_first = (int)strtoul(oa, NULL, 10);
        if (ft_first <= 0 || ft_first >= FXH_MAX_LINE) errx(1, "Invalid number bases to keep (-f %s)", oa);
        ft_by_position = 1;
        break;
    case 'l':
        if (oa == NULL) errx(1, "[-l] parameter requires an argument value");
        ft_last = (int)strtoul(oa, NULL, 10);
        if (ft_last <= 0 || ft_last >= FXH_MAX_LINE) errx(1, "Invalid number bases to keep (-l %s)", oa);
        ft_by_position = 1;
        break;
    case 't':
        if (oa == NULL) errx(1, "[-t] parameter requires an argument value");
        ft_trim_last = (unsigned int)strtoul(oa, NULL, 10);
        if (ft_trim_last <= 0 || ft_trim_last >= FXH_MAX_LINE) errx(1, "Invalid number bases to trim (-t %s)", oa);
        ft_from_end = 1;
        break;
    case 'm':
        if (oa == NULL) errx(1, "[-t] parameter requires an argument value");
        ft_min_len = (unsigned int)strtoul(oa, NULL, 10);
        if (ft_min_len <= 0 || ft_min_len >= FXH_MAX_LINE) errx(1, "Invalid minimum length value (-m %s)", oa);
        break;
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

/* src/fastx_trimmer/fastx_trimmer.c:120-148: which slice of the read survives; -1 = the read is dropped */
static int ft_slice(int len, int *start)
{
    size_t L = (size_t)len, s0 = 0;
    if (ft_last != 0 && (size_t)ft_last < L) L = (size_t)ft_last;
    if (ft_first != 1) {
        if (L < (size_t)ft_first) return -1;
        s0 = (size_t)ft_first - 1;
        L = L - (size_t)ft_first + 1;
    }
    if (ft_trim_last > 0) {
        if (L <= ft_trim_last) return -1;
        const size_t i = L - ft_trim_last;
        if (i < ft_min_len) return -1;
        L = i;
    }
    *start = (int)s0;
    return (int)L;
}

static int main_fastx_trimmer(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "l:f:t:m:", ft_args, fxh_usage_fastx_trimmer);
    if (ft_by_position && ft_from_end) errx(1, "[-t], [-f] and [-l] options can not be used together. Use [-t] or [-l,-f]");
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_validate_host(ctx, &gb, batch_q(b), &rep), "fxg_validate_host");   /* the reader's checks */
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++) {
            int st = 0;
            const int nl = ft_slice(b->len[i], &st);
            if (nl >= 0)
                fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride + st, b->qual + (size_t)i * b->stride + st, nl);
        }
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        if (ft_first != 1 || ft_last != 0) fprintf(f, "Trimming: base %d to %d\n", ft_first, ft_last);
        if (ft_trim_last) {
            fprintf(f, "Trimming %d bases from the end of the reads\n", ft_trim_last);
            if (ft_min_len) fprintf(f, "Discarding reads shorter than %d bases\n", ft_min_len);
        }
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastq_masker   */
static int mk_min_quality = 10;
static char mk_char = 'N';

static int mk_args(int oi, int optc, char *oa)
{
    (void)oi;
    switch (optc) {
    case 'q':
        if (oa == NULL) errx(1, "[-q] parameter requires an argument value");
        mk_min_quality = atoi(oa);
        if (mk_min_quality < -40) errx(1, "Invalid minimum length value (-q %s)", oa);
        break;
    case 'r':
        if (oa == NULL) errx(1, "[-r] parameter requires an argument value");
        if (strlen(oa) != 1) errx(1, "[-r] parameter requires a single character as value");
        mk_char = oa[0];
        break;
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

static int main_masker(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "q:r:", mk_args, fxh_usage_fastq_masker);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTQ_ONLY, fxh_q_offset(), 0);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), 1, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    size_t masked_reads = 0, masked_nuc = 0;
    pbuf os = { 0, 0 }, fl = { 0, 0 };
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        uint8_t *oseq = (uint8_t *)pbuf_get(&os, (size_t)b->n * b->stride), *flag = (uint8_t *)pbuf_get(&fl, (size_t)b->n);
        fxg_batch gb = fxh_as_fxg_batch(b, 1);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_mask_host(ctx, &gb, batch_q(b), mk_min_quality, (unsigned char)mk_char, oseq, flag, &rep), "fxg_mask_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++) {
            if (flag[i]) masked_reads += (size_t)b->weight[i];
            fxh_write_record(wr, b, i, oseq + (size_t)i * b->stride, b->qual + (size_t)i * b->stride, b->len[i]);
        }
        masked_nuc += (size_t)rep.aux[0];
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Minimum Quality Threshold: %d\n", mk_min_quality);
        fprintf(f, "Low-quality nucleotides replaced with '%c'\n", mk_char);
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
        fprintf(f, "Masked reads: %zu\n", masked_reads);
        fprintf(f, "Masked nucleotides: %zu\n", masked_nuc);
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastx_artifacts_filter   */
static int main_artifacts(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "", NULL, fxh_usage_fastx_artifacts_filter);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    pbuf kp = { 0, 0 };
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        uint8_t *keep = (uint8_t *)pbuf_get(&kp, (size_t)b->n);
        fxg_batch gb = fxh_as_fxg_batch(b, fastq);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_artifacts_host(ctx, &gb, batch_q(b), keep, &rep), "fxg_artifacts_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++)
            if (keep[i])
                fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride, b->qual + (size_t)i * b->stride, b->len[i]);
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
        size_t discarded = fxh_num_input_reads(rd) - fxh_num_output_reads(wr);
        fprintf(f, "discarded %zu (%zu%%) artifact reads.\n", discarded, (discarded * 100) / fxh_num_input_reads(rd));
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ fastq_to_fasta
 * src/fastq_to_fasta/fastq_to_fasta.c:50-103: FASTQ in, FASTA out; reads with an 'N' are dropped unless -n (the test runs on
 * the GPU, K-HASN, fused with the reader's checks); -r renames the identifiers to the running output count. */
static int f2a_rename = 0, f2a_discard_n = 1;
static int f2a_parse(int optind_, int optc, char *optarg_)
{
    (void)optind_; (void)optarg_;
    switch (optc) {
    case 'n': f2a_discard_n = 0; break;
    case 'r': f2a_rename = 1; break;
    default: errx(1, "Unknown argument (%c)", optc);
    }
    return 1;
}

static int main_fastq_to_fasta(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "rn", f2a_parse, fxh_usage_fastq_to_fasta);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTQ_ONLY, fxh_q_offset(), 0);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), 0 /* FASTA */, fxh_compress_output());
    fxg_ctx *ctx = fxh_gpu_open();
    pbuf fp = { 0, 0 };
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        uint8_t *has_n = (uint8_t *)pbuf_get(&fp, (size_t)b->n);
        fxg_batch gb = fxh_as_fxg_batch(b, 1);
        fxg_report rep;
        fxh_gpu_check(ctx, fxg_has_n_host(ctx, &gb, batch_q(b), has_n, &rep), "fxg_has_n_host");
        const int64_t lim = rep.first_bad_read >= 0 ? rep.first_bad_read : b->n;
        for (int64_t i = 0; i < lim; i++) {
            if (f2a_discard_n && has_n[i]) continue;
            const uint8_t *srow = b->seq + (size_t)i * b->stride;
            if (f2a_rename) {
                char num[32];
                const int nl = snprintf(num, sizeof(num), "%zu", fxh_num_output_reads(wr) + 1);
                fxh_write_record_named(wr, b, i, srow, NULL, b->len[i], num, nl);
            } else {
                fxh_write_record(wr, b, i, srow, NULL, b->len[i]);
            }
        }
        if (rep.first_bad_read >= 0) { fxh_writer_close(wr); fxh_die_bad_record(rd, b, rep.first_bad_read); }
    }
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
        if (f2a_discard_n) {
            size_t discarded = fxh_num_input_reads(rd) - fxh_num_output_reads(wr);
            fprintf(f, "discarded %zu (%zu%%) low-quality reads.\n", discarded, (discarded * 100) / fxh_num_input_reads(rd));
        }
    }
    fxg_destroy(ctx);
    return 0;
}

/* ================================================================================ dispatch */
int main(int argc, char **argv)
{
    char *self = strdup(argv[0]);
    const char *name = basename(self);
    if (!strcmp(name, "fastq_quality_trimmer")) return main_trimmer(argc, argv);
    if (!strcmp(name, "fastq_quality_filter")) return main_filter(argc, argv);
    if (!strcmp(name, "fastx_reverse_complement")) return main_revcomp(argc, argv);
    if (!strcmp(name, "fastx_clipper")) return main_clipper(argc, argv);
    if (!strcmp(name, "fastx_collapser")) return main_collapser(argc, argv);
    if (!strcmp(name, "fastx_quality_stats")) return main_stats(argc, argv);
    if (!strcmp(name, "fastx_trimmer")) return main_fastx_trimmer(argc, argv);
    if (!strcmp(name, "fastq_masker")) return main_masker(argc, argv);
    if (!strcmp(name, "fastx_artifacts_filter")) return main_artifacts(argc, argv);
    if (!strcmp(name, "fastq_to_fasta")) return main_fastq_to_fasta(argc, argv);
    fprintf(stderr, "%s: multi-call binary; invoke it as fastq_quality_trimmer, fastq_quality_filter, fastx_reverse_complement, "
                    "fastx_clipper, fastx_collapser, fastx_quality_stats, fastx_trimmer, fastq_masker, fastx_artifacts_filter or fastq_to_fasta\n", name);
    return 1;
}
