/* CPU-only harness for the streaming engine of the drop-in tools (fastx_toolkit_b200/csrc/host/fxh_stream.c): reader thread,
 * record-boundary splitting and carry-over, workers, in-order writer, fallback to the record path — with NO GPU.  The GPU
 * text path (fxg_text_*) is replaced by a TEST DOUBLE that only indexes the lines of a chunk and copies whole records
 * through (the identity transform), flagging what K-RECS would flag; so the program must behave like the reference's
 * fastx_trimmer with default arguments: same bytes for valid input, same prefix + message + exit status for structurally
 * broken input (tests/test_stream_engine.py).  Test infrastructure only — nothing here is linked into the product. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fxg.h"
#include "fxh.h"
#include "fxh_stream.h"

void *fxg_alloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 1); }
void fxg_free_pinned(void *p) { free(p); }
int fxg_host_register(void *p, size_t bytes) { (void)p; (void)bytes; return FXG_OK; }
int fxg_host_unregister(void *p) { (void)p; return FXG_OK; }
int fxg_init(int device, fxg_ctx **out) { (void)device; *out = (fxg_ctx *)malloc(8); return FXG_OK; }
void fxg_destroy(fxg_ctx *ctx) { free(ctx); }
const char *fxg_strerror(int code) { (void)code; return "stub"; }
const char *fxg_last_error(const fxg_ctx *ctx) { (void)ctx; return "stub"; }

struct fxg_text { int fasta; };
int fxg_text_new(fxg_ctx *ctx, int device, size_t max_chunk_bytes, fxg_text **out)
{
    (void)ctx; (void)device; (void)max_chunk_bytes;
    *out = (fxg_text *)calloc(1, sizeof(fxg_text));
    return FXG_OK;
}
void fxg_text_free(fxg_text *t) { free(t); }
int fxg_text_set_format(fxg_text *t, int fasta) { t->fasta = fasta; return FXG_OK; }
int fxg_text_set_deflate(fxg_text *t, int on) { (void)t; (void)on; return FXG_OK; }      /* the double emits plain text: the writer stores it */
const char *fxg_text_error(const fxg_text *t) { (void)t; return "stub"; }

/* identity "op": whole records of the chunk, copied through; the structural checks of K-RECS */
int fxg_text_run_host(fxg_text *t, int op, const char *text, size_t bytes, int q_offset, int a0, int a1, char *out, fxg_text_report *rep)
{
    (void)op; (void)q_offset; (void)a0; (void)a1;
    memset(rep, 0, sizeof *rep);
    rep->anomaly_record = -1;
    const int lpr = t->fasta ? 2 : 4;
    size_t nlines = 0;
    for (size_t i = 0; i < bytes; i++) nlines += text[i] == '\n';
    const size_t nrec = nlines / (size_t)lpr;
    rep->n_records = (int64_t)nrec;
    if (nrec == 0) return FXG_OK;
    size_t pos = 0, rec = 0;
    int maxlen = 0;
    while (rec < nrec) {
        size_t st[4], ln[4];
        for (int k = 0; k < lpr; k++) {
            const char *nl = (const char *)memchr(text + pos, '\n', bytes - pos);
            st[k] = pos; ln[k] = (size_t)(nl - (text + pos));
            if (ln[k] > 0 && text[pos + ln[k] - 1] == '\r') ln[k]--;
            pos = (size_t)(nl - text) + 1;
        }
        int an = 0;
        if (ln[0] == 0 || text[st[0]] != (lpr == 4 ? '@' : '>')) an = FXG_TEXT_PREFIX;
        else if (ln[1] == 0) an = FXG_TEXT_EMPTY_SEQ;
        else if (lpr == 4 && ln[3] != ln[1]) an = FXG_TEXT_QUAL_LEN;
        else if (memchr(text + st[0], '\r', ln[0]) || (lpr == 4 && memchr(text + st[2], '\r', ln[2]))) an = FXG_TEXT_BAD_RECORD;
        if (an) { rep->anomaly = an; rep->anomaly_record = (int64_t)rec; rep->consumed_bytes = 0; return FXG_OK; }
        if ((int)ln[1] > maxlen) maxlen = (int)ln[1];
        rec++;
    }
    rep->consumed_bytes = (int64_t)pos;
    rep->max_len = rep->min_len = maxlen;
    memcpy(out, text, pos);
    rep->out_bytes = (int64_t)pos;
    rep->n_out_records = (int64_t)nrec;
    rep->n_reads = rep->n_out_reads = (int64_t)nrec;
    usleep(200 * (unsigned)(nrec % 7));          /* workers finish out of order */
    return FXG_OK;
}
int fxg_text_clip_host(fxg_text *t, const char *a, size_t b, int c, const fxg_clip_opts *d, int e, int f, char *g, fxg_text_report *h)
{ (void)t; (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; return FXG_ERR_UNSUPPORTED; }
/* a stateful op (what K-STATS / the collapser are): the records of a clean chunk are counted into *hist; the engine must
 * never let a chunk count whose predecessor is handed back to the record path */
int fxg_text_stats_host(fxg_text *t, const char *text, size_t bytes, int q, uint64_t *hist, int32_t e, fxg_text_report *rep)
{
    (void)e;
    static char sink[1 << 26];
    if (bytes > sizeof sink) return FXG_ERR_ARG;
    int rc = fxg_text_run_host(t, 2, text, bytes, q, 0, 0, sink, rep);
    if (rc == FXG_OK && rep->anomaly == 0) __atomic_fetch_add(hist, (uint64_t)rep->n_records, __ATOMIC_RELAXED);
    return rc;
}
int fxg_text_collapse_host(fxg_text *t, const char *a, size_t b, int c, fxg_collapser *d, int64_t e, fxg_text_report *f)
{ (void)t; (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; return FXG_ERR_UNSUPPORTED; }

static const char *const usage_text = "usage: fxs_harness [-h] [-v] [-z] [-i INFILE] [-o OUTFILE] [-Q N]\n";

int main(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "", NULL, usage_text);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxs_job job;
    memset(&job, 0, sizeof job);
    job.op = FXS_REVCOMP; job.ngpu = 2; job.first_dev = 0;
    uint64_t counted[2] = { 0, 0 }, *hists[2] = { &counted[0], &counted[1] };
    const int stateful = getenv("FXS_HARNESS_STATEFUL") != NULL;
    if (stateful) { job.op = FXS_STATS; job.hist_dev = hists; job.max_cycles = 1; }
    const int fell_back = fxs_run(&job, rd, stateful ? NULL : wr);
    fxh_batch *b;
    uint64_t host_counted = 0;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL) {
        host_counted += (uint64_t)b->n;
        for (int64_t i = 0; i < b->n && !stateful; i++)
            fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride, b->qual ? b->qual + (size_t)i * b->stride : NULL, b->len[i]);
    }
    if (stateful) fprintf(stderr, "[harness] counted=%llu\n", (unsigned long long)(counted[0] + counted[1] + host_counted));
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
    }
    if (getenv("FXS_HARNESS_REPORT")) fprintf(stderr, "[harness] engine records=%lld chunks=%lld fallback=%d\n", (long long)job.records, (long long)job.chunks, fell_back);
    return 0;
}
