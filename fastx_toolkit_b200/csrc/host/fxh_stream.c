/*
 * fxh_stream.c — see fxh_stream.h.  Threads: 1 reader (+ helpers for parallel pread on regular files), ngpu x W workers,
 * the caller as the in-order writer.  Buffers are pinned (DMA source / destination of the GPU text path).
 */
#define _GNU_SOURCE
#include <err.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include "fxh_stream.h"

#define FXS_HEAD ((size_t)1 << 20)     /* room in front of a block for the partial record carried over from the previous one */

typedef struct fxs_chunk {
    int64_t seq;
    char *buf;                  /* pinned: FXS_HEAD + chunk_bytes */
    size_t head;                /* the chunk's text starts at buf + head (FXS_HEAD - carried bytes) */
    size_t len;                 /* bytes of whole records from head */
    size_t block_len;           /* bytes read from the input into buf + FXS_HEAD */
    int last;                   /* end of input reached with this block */
    /* result */
    char *out;
    fxg_text_report rep;
    int rc;
    char errmsg[256];
    char *lastseq; int lastseq_len;     /* CLIP: sequence line of the last record (the aligner's query buffer afterwards) */
    struct fxs_chunk *next;
} fxs_chunk;

typedef struct {
    fxs_job *job;
    int fd, fastq, q_offset;
    size_t chunk_bytes, out_cap;
    /* what the fxh_reader had already read */
    const char *mem; size_t mem_len, mem_pos; int mem_eof;
    int fd_eof;
    int read_threads;
    off_t file_pos; int regular;

    pthread_mutex_t mu;
    pthread_cond_t cv;
    fxs_chunk *free_in;
    char **free_out; int n_free_out;
    fxs_chunk *ready_head, *ready_tail;
    fxs_chunk **done; int done_cap;
    int stop, reader_done;
    int64_t n_chunks;           /* valid once reader_done */
    int expect_len;             /* CLIP: read length of chunk 0 (0 = not known yet) */
    int reader_failed; char reader_err[256];
    int deflate;                /* -z: the chunks leave the GPU as DEFLATE blocks */
    int ordered_apply;          /* STATS / COLLAPSE: a chunk changes state on the GPU, so it runs only once every earlier chunk is
                                   known to be clean — a chunk that is handed back to the record path must not have successors
                                   that were already counted */
    int64_t next_apply;
    pthread_mutex_t col_mu[64];
} fxs_state;

typedef struct { fxs_state *s; int index, dev; pthread_t th; int started; } fxs_worker;

static size_t env_size(const char *name, size_t dflt, size_t lo)
{
    const char *e = getenv(name);
    long long v = e ? atoll(e) : 0;
    return v >= (long long)lo ? (size_t)v : dflt;
}

/* ---- reader ------------------------------------------------------------------------------------------------------- */
typedef struct { int fd; char *dst; size_t n; off_t off; ssize_t got; int err_no; } pread_job;
static void *pread_main(void *arg)
{
    pread_job *j = (pread_job *)arg;
    size_t got = 0;
    while (got < j->n) {
        ssize_t k = pread(j->fd, j->dst + got, j->n - got, j->off + (off_t)got);
        if (k < 0) { if (errno == EINTR) continue; j->err_no = errno; break; }
        if (k == 0) break;
        got += (size_t)k;
    }
    j->got = (ssize_t)got;
    return NULL;
}

/* up to `want` bytes of input into dst: first what the fxh_reader had buffered, then the file descriptor */
static size_t fill(fxs_state *s, char *dst, size_t want)
{
    size_t got = 0;
    if (s->mem_pos < s->mem_len) {
        size_t k = s->mem_len - s->mem_pos;
        if (k > want) k = want;
        memcpy(dst, s->mem + s->mem_pos, k);
        s->mem_pos += k; got = k;
    }
    if (s->mem_pos >= s->mem_len && s->mem_eof) s->fd_eof = 1;
    if (got < want && !s->fd_eof && s->regular && s->read_threads > 1 && want - got >= ((size_t)8 << 20)) {
        /* regular file: the block is read by several threads at once (one pread stream per thread) */
        const int T = s->read_threads;
        pread_job jobs[16]; pthread_t th[16];
        const size_t total = want - got, part = (total / (size_t)T + 4095) & ~(size_t)4095;
        int nt = 0;
        for (size_t o = 0; o < total && nt < 16; o += part, nt++) {
            jobs[nt].fd = s->fd; jobs[nt].dst = dst + got + o; jobs[nt].n = (total - o < part) ? total - o : part;
            jobs[nt].off = s->file_pos + (off_t)o; jobs[nt].got = 0; jobs[nt].err_no = 0;
            if (nt > 0 && pthread_create(&th[nt], NULL, pread_main, &jobs[nt]) != 0) err(1, "pthread_create");
        }
        pread_main(&jobs[0]);
        for (int i = 1; i < nt; i++) pthread_join(th[i], NULL);
        size_t sum = 0; int short_read = 0;
        for (int i = 0; i < nt; i++) {
            if (jobs[i].err_no) { errno = jobs[i].err_no; return (size_t)-1; }
            if (!short_read) sum += (size_t)jobs[i].got;
            if ((size_t)jobs[i].got < jobs[i].n) short_read = 1;
        }
        got += sum; s->file_pos += (off_t)sum;
        if (short_read) s->fd_eof = 1;
        return got;
    }
    while (got < want && !s->fd_eof) {
        ssize_t k = s->regular ? pread(s->fd, dst + got, want - got, s->file_pos) : read(s->fd, dst + got, want - got);
        if (k < 0) { if (errno == EINTR) continue; return (size_t)-1; }
        if (k == 0) { s->fd_eof = 1; break; }
        got += (size_t)k;
        if (s->regular) s->file_pos += k;
    }
    return got;
}

/* Offset of the start of the last record that may be incomplete: everything before it is whole records, provided the
 * input is well formed (the GPU checks that).  FASTQ: the last line that starts with '@' and whose second-next line starts
 * with '+' — a quality line may start with '@', but then the line two further on is a sequence line, never a '+' line.
 * FASTA: the last line that starts with '>'.  0 = no such line in the data. */
static size_t find_split(const char *d, size_t n, int fastq)
{
    size_t ls[3] = { 0, 0, 0 };      /* starts of the two lines after the candidate (ls[0] = next, ls[1] = second next) */
    int have = 0;
    size_t p = n;
    while (p > 0) {
        /* start of the line that contains byte p-1 */
        const char *nl = (const char *)memrchr(d, '\n', p - 1);
        const size_t start = nl ? (size_t)(nl - d) + 1 : 0;
        if (start < n) {
            if (!fastq) { if (d[start] == '>') return start; }
            else if (have >= 2 && d[start] == '@' && ls[1] < n && d[ls[1]] == '+') return start;
        }
        ls[1] = ls[0]; ls[0] = start; if (have < 2) have++;
        if (start == 0) break;
        p = start;               /* the newline that ends the previous line sits at start-1 */
    }
    return 0;
}

static fxs_chunk *get_free_in(fxs_state *s)
{
    pthread_mutex_lock(&s->mu);
    while (!s->free_in && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
    fxs_chunk *c = NULL;
    if (!s->stop) { c = s->free_in; s->free_in = c->next; c->next = NULL; }
    pthread_mutex_unlock(&s->mu);
    return c;
}

static void *reader_main(void *arg)
{
    fxs_state *s = (fxs_state *)arg;
    char *carry = (char *)malloc(FXS_HEAD);
    size_t carry_len = 0;
    int64_t seq = 0;
    if (!carry) err(1, "out of memory");
    for (;;) {
        fxs_chunk *c = get_free_in(s);
        if (!c) break;
        memcpy(c->buf + FXS_HEAD - carry_len, carry, carry_len);
        c->head = FXS_HEAD - carry_len;
        const size_t n = fill(s, c->buf + FXS_HEAD, s->chunk_bytes);
        if (n == (size_t)-1) {
            pthread_mutex_lock(&s->mu);
            s->reader_failed = 1; snprintf(s->reader_err, sizeof s->reader_err, "%s", strerror(errno));
            c->next = s->free_in; s->free_in = c;
            pthread_mutex_unlock(&s->mu);
            break;
        }
        c->block_len = n;
        c->last = s->fd_eof && s->mem_pos >= s->mem_len;
        const size_t total = carry_len + n;
        c->seq = seq; c->out = NULL; c->rc = 0; c->lastseq_len = 0;
        if (c->last) { c->len = total; carry_len = 0; }
        else {
            const size_t sp = find_split(c->buf + c->head, total, s->fastq);
            if (sp == 0 || total - sp > FXS_HEAD) c->len = 0;      /* no record boundary in a whole block: the host path's problem */
            else { c->len = sp; carry_len = total - sp; memcpy(carry, c->buf + c->head + sp, carry_len); }
        }
        const int give_up = !c->last && c->len == 0;
        if (total == 0 && c->last && seq > 0) {                    /* the input ended exactly at the previous block */
            pthread_mutex_lock(&s->mu);
            c->next = s->free_in; s->free_in = c;
            pthread_mutex_unlock(&s->mu);
            break;
        }
        pthread_mutex_lock(&s->mu);
        c->next = NULL;
        if (s->ready_tail) s->ready_tail->next = c; else s->ready_head = c;
        s->ready_tail = c;
        seq++;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mu);
        if (c->last || give_up) break;
    }
    free(carry);
    pthread_mutex_lock(&s->mu);
    s->reader_done = 1; s->n_chunks = seq;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
    return NULL;
}

/* ---- workers ------------------------------------------------------------------------------------------------------- */
static void *worker_main(void *arg)
{
    fxs_worker *w = (fxs_worker *)arg;
    fxs_state *s = w->s;
    fxs_job *job = s->job;
    fxg_ctx *ctx = fxh_gpu_open_dev(w->dev);
    fxg_text *tx = NULL;
    int rc = fxg_text_new(ctx, w->dev, FXS_HEAD + s->chunk_bytes, &tx);
    if (rc != FXG_OK) errx(1, "fxg_text_new failed on GPU %d: %s", w->dev, fxg_strerror(rc));
    fxg_text_set_format(tx, !s->fastq);
    if (s->deflate && fxg_text_set_deflate(tx, 1) != FXG_OK) errx(1, "fxg_text_set_deflate failed on GPU %d: %s", w->dev, fxg_text_error(tx));
    for (;;) {
        pthread_mutex_lock(&s->mu);
        while (s->n_free_out == 0 && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
        if (s->stop) { pthread_mutex_unlock(&s->mu); break; }
        char *out = s->free_out[--s->n_free_out];
        while (!s->ready_head && !s->reader_done && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
        if (!s->ready_head || s->stop) { s->free_out[s->n_free_out++] = out; pthread_cond_broadcast(&s->cv); pthread_mutex_unlock(&s->mu); break; }
        fxs_chunk *c = s->ready_head;
        s->ready_head = c->next;
        if (!s->ready_head) s->ready_tail = NULL;
        c->next = NULL;
        if (job->op == FXS_CLIP && c->seq > 0)
            while (s->expect_len == 0 && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
        const int expect = s->expect_len;
        int skip = 0;
        if (s->ordered_apply) {
            while (s->next_apply != c->seq && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
            skip = s->stop;
        }
        pthread_mutex_unlock(&s->mu);
        if (skip) {                                    /* the engine is winding down: the chunk stays unprocessed (its bytes go back to the reader) */
            pthread_mutex_lock(&s->mu);
            s->free_out[s->n_free_out++] = out;
            pthread_mutex_unlock(&s->mu);
            break;
        }

        c->out = out;
        const char *text = c->buf + c->head;
        memset(&c->rep, 0, sizeof c->rep);
        if (c->len == 0) { c->rc = FXG_OK; c->rep.anomaly = FXG_TEXT_LONG_LINE; }       /* reader found no record boundary */
        else if (expect < 0 && job->op == FXS_CLIP) { c->rc = FXG_OK; c->rep.anomaly = FXG_TEXT_MIXED_LEN; }
        else switch (job->op) {
        case FXS_TRIM: case FXS_FILTER: case FXS_REVCOMP:
            c->rc = fxg_text_run_host(tx, job->op, text, c->len, s->q_offset, job->a0, job->a1, out, &c->rep);
            break;
        case FXS_CLIP:
            c->rc = fxg_text_clip_host(tx, text, c->len, s->q_offset, job->clip, job->a0, expect, out, &c->rep);
            break;
        case FXS_STATS:
            c->rc = fxg_text_stats_host(tx, text, c->len, s->q_offset, job->hist_dev[w->dev - job->first_dev], job->max_cycles, &c->rep);
            break;
        case FXS_COLLAPSE:
            pthread_mutex_lock(&s->col_mu[w->dev - job->first_dev]);
            c->rc = fxg_text_collapse_host(tx, text, c->len, s->q_offset, job->collapsers[w->dev - job->first_dev], c->seq << 32, &c->rep);
            pthread_mutex_unlock(&s->col_mu[w->dev - job->first_dev]);
            break;
        default: c->rc = FXG_ERR_ARG;
        }
        if (c->rc != FXG_OK) snprintf(c->errmsg, sizeof c->errmsg, "%s (%s)", fxg_strerror(c->rc), fxg_text_error(tx));
        if (job->op == FXS_CLIP && c->rc == FXG_OK && c->rep.anomaly == 0 && c->rep.n_records > 0) {
            /* sequence line of the last consumed record: 2 newlines back from the end of the record */
            size_t e = (size_t)c->rep.consumed_bytes - 1;
            int nl = 0;
            while (e > 0 && nl < (s->fastq ? 2 : 0)) { e--; if (text[e] == '\n') nl++; }
            size_t s2 = e;
            while (s2 > 0 && text[s2 - 1] != '\n') s2--;
            int L = (int)(e - s2);
            if (L > 0 && text[s2 + (size_t)L - 1] == '\r') L--;
            if (L > FXH_MAX_LINE) L = FXH_MAX_LINE;
            memcpy(c->lastseq, text + s2, (size_t)L);
            c->lastseq_len = L;
        }
        pthread_mutex_lock(&s->mu);
        if (s->ordered_apply && c->rc == FXG_OK && c->rep.anomaly == 0 && c->rep.n_records > 0 && (size_t)c->rep.consumed_bytes == c->len)
            s->next_apply = c->seq + 1;                /* clean and complete: the next chunk may run */
        if (job->op == FXS_CLIP && c->seq == 0 && s->expect_len == 0)
            s->expect_len = (c->rc == FXG_OK && c->rep.anomaly == 0 && c->rep.max_len > 0) ? c->rep.max_len : -1;
        s->done[c->seq % s->done_cap] = c;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mu);
    }
    fxg_text_free(tx);
    fxg_destroy(ctx);
    return NULL;
}

/* ---- the caller's thread: in-order writer ---------------------------------------------------------------------------- */
int fxs_run(fxs_job *job, fxh_reader *rd, fxh_writer *wr)
{
    if (!fxh_text_path_enabled()) return 1;
    const double t0 = fxh_now();
    fxs_state S;
    memset(&S, 0, sizeof S);
    fxs_state *s = &S;
    s->job = job;
    s->fastq = fxh_reader_is_fastq(rd);
    if (!s->fastq && (job->op == FXS_TRIM || job->op == FXS_FILTER)) return 1;
    s->q_offset = fxh_q_offset();
    s->fd = fxh_reader_fd(rd);
    char *mem; size_t mem_len; int mem_eof;
    fxh_reader_detach(rd, &mem, &mem_len, &mem_eof);
    if (mem_len == 0) return 1;
    s->mem = mem; s->mem_len = mem_len; s->mem_eof = mem_eof;
    {
        struct stat sb;
        s->regular = (fstat(s->fd, &sb) == 0 && S_ISREG(sb.st_mode));
        if (s->regular) { s->file_pos = lseek(s->fd, 0, SEEK_CUR); if (s->file_pos < 0) s->regular = 0; }
    }
    s->chunk_bytes = env_size("FASTX_CHUNK_BYTES", (size_t)32 << 20, 16384);
    if (mem_eof && mem_len + 4096 < s->chunk_bytes) s->chunk_bytes = (mem_len + 4096 + 4095) & ~(size_t)4095;   /* small input: one chunk */
    s->read_threads = (int)env_size("FASTX_READ_THREADS", 4, 1);
    if (s->read_threads > 16) s->read_threads = 16;
    int ngpu = job->ngpu > 0 ? job->ngpu : 1;
    if (ngpu > 1 && s->regular && !getenv("FASTX_GPUS_FORCE")) {
        /* every extra GPU costs a CUDA context (~0.5 s): worth it from about 4 GB of input per GPU on */
        struct stat sb;
        if (fstat(s->fd, &sb) == 0) {
            const long long per_gpu = 4ll << 30;
            const int useful = (int)((sb.st_size + per_gpu - 1) / per_gpu);
            if (useful < ngpu) ngpu = useful < 1 ? 1 : useful;
        }
    }
    if ((job->op == FXS_STATS || job->op == FXS_COLLAPSE) && ngpu != job->ngpu) ngpu = job->ngpu;      /* the caller merges ngpu partial results */
    int W = (int)env_size("FASTX_WORKERS", 3, 1);
    if (W > 8) W = 8;
    int nworkers = ngpu * W;
    if (mem_eof && mem_len <= s->chunk_bytes) nworkers = 1;                 /* one chunk in all */
    if (job->op == FXS_COLLAPSE && nworkers > 2 * ngpu) nworkers = 2 * ngpu;  /* a table is one object: adds to it are serialised */
    const int has_out = job->op == FXS_TRIM || job->op == FXS_FILTER || job->op == FXS_REVCOMP || job->op == FXS_CLIP;
    s->deflate = has_out && wr && fxh_writer_frames_gzip(wr);
    s->ordered_apply = (job->op == FXS_STATS || job->op == FXS_COLLAPSE);
    s->out_cap = has_out ? (FXS_HEAD + s->chunk_bytes) + (FXS_HEAD + s->chunk_bytes) / 4 + 64 : 64;
    const int n_in = 2 * nworkers + 2, n_out = nworkers + 1;
    pthread_mutex_init(&s->mu, NULL); pthread_cond_init(&s->cv, NULL);
    for (int g = 0; g < 64; g++) pthread_mutex_init(&s->col_mu[g], NULL);
    s->done_cap = n_in + 2;
    s->done = (fxs_chunk **)calloc((size_t)s->done_cap, sizeof(fxs_chunk *));
    s->free_out = (char **)calloc((size_t)n_out, sizeof(char *));
    fxs_chunk *chunks = (fxs_chunk *)calloc((size_t)n_in, sizeof(fxs_chunk));
    if (!s->done || !s->free_out || !chunks) err(1, "out of memory");
    for (int i = 0; i < n_in; i++) {
        chunks[i].buf = (char *)fxg_alloc_pinned(FXS_HEAD + s->chunk_bytes + 64);
        chunks[i].lastseq = (char *)malloc(FXH_MAX_LINE + 16);
        if (!chunks[i].buf || !chunks[i].lastseq) errx(1, "cannot allocate the pinned input buffers (%d x %zu bytes)", n_in, FXS_HEAD + s->chunk_bytes);
        chunks[i].next = s->free_in; s->free_in = &chunks[i];
    }
    char **outs = (char **)calloc((size_t)n_out, sizeof(char *));
    if (!outs) err(1, "out of memory");
    for (int i = 0; i < n_out; i++) {
        outs[i] = (char *)fxg_alloc_pinned(s->out_cap);
        if (!outs[i]) errx(1, "cannot allocate the pinned output buffers (%d x %zu bytes)", n_out, s->out_cap);
        s->free_out[s->n_free_out++] = outs[i];
    }
    pthread_t reader;
    if (pthread_create(&reader, NULL, reader_main, s) != 0) err(1, "pthread_create");
    fxs_worker *workers = (fxs_worker *)calloc((size_t)nworkers, sizeof(fxs_worker));
    if (!workers) err(1, "out of memory");
    for (int i = 0; i < nworkers; i++) {
        workers[i].s = s; workers[i].index = i; workers[i].dev = job->first_dev + (i % ngpu);
        if (pthread_create(&workers[i].th, NULL, worker_main, &workers[i]) != 0) err(1, "pthread_create");
        workers[i].started = 1;
    }

    int64_t next = 0;
    int fallback = 0;
    size_t fallback_at = 0;
    fxs_chunk *fb_chunk = NULL;
    const int lpr = s->fastq ? 4 : 2;
    for (;;) {
        pthread_mutex_lock(&s->mu);
        const double tw = fxh_now();
        for (;;) {
            fxs_chunk *d = s->done[next % s->done_cap];
            if (d && d->seq == next) break;
            if (s->reader_failed) { pthread_mutex_unlock(&s->mu); errx(1, "failed to read input file '%s': %s", fxh_input_filename(), s->reader_err); }
            if (s->reader_done && next >= s->n_chunks) break;
            pthread_cond_wait(&s->cv, &s->mu);
        }
        job->t_wait_gpu += fxh_now() - tw;
        fxs_chunk *c = s->done[next % s->done_cap];
        if (!c || c->seq != next) { pthread_mutex_unlock(&s->mu); break; }      /* all chunks done */
        s->done[next % s->done_cap] = NULL;
        pthread_mutex_unlock(&s->mu);

        if (c->rc != FXG_OK) errx(1, "GPU text path failed: %s", c->errmsg);
        if (c->rep.anomaly != 0 || c->rep.n_records == 0) { fallback = 1; fallback_at = c->head; fb_chunk = c; break; }
        if (wr && has_out) {
            const double t1 = fxh_now();
            if (c->rep.deflated)
                fxh_writer_write_deflated(wr, c->out, (size_t)c->rep.out_bytes, (uint64_t)c->rep.raw_out_bytes, c->rep.out_crc32_pure, c->rep.n_out_records,
                                          s->fastq ? c->rep.n_out_records : c->rep.n_out_reads);
            else
                fxh_writer_write_now(wr, c->out, (size_t)c->rep.out_bytes, c->rep.n_out_records, s->fastq ? c->rep.n_out_records : c->rep.n_out_reads);
            job->t_write += fxh_now() - t1;
        }
        fxh_reader_account(rd, c->rep.n_records, s->fastq ? c->rep.n_records : c->rep.n_reads, lpr);
        job->records += c->rep.n_records;
        job->reads += s->fastq ? c->rep.n_records : c->rep.n_reads;
        job->chunks++;
        if (c->rep.max_len > job->max_len) job->max_len = c->rep.max_len;
        if (job->op == FXS_CLIP) {
            for (int k = 0; k < 6; k++) job->clip_class[k] += (unsigned int)c->rep.clip_class[k];
            if (c->lastseq_len > 0) fxh_reader_seed_shadow(rd, c->lastseq, c->lastseq_len);
        }
        const int leftover = (size_t)c->rep.consumed_bytes < c->len;
        const int was_last = c->last;
        if (leftover) { fallback = 1; fallback_at = c->head + (size_t)c->rep.consumed_bytes; fb_chunk = c; break; }
        /* recycle the buffers */
        pthread_mutex_lock(&s->mu);
        s->free_out[s->n_free_out++] = c->out; c->out = NULL;
        c->next = s->free_in; s->free_in = c;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mu);
        next++;
        if (was_last) break;
    }

    /* wind down */
    pthread_mutex_lock(&s->mu);
    s->stop = 1;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
    pthread_join(reader, NULL);
    for (int i = 0; i < nworkers; i++) if (workers[i].started) pthread_join(workers[i].th, NULL);

    if (fallback) {
        /* give the reader everything from the first unprocessed byte on: the rest of this chunk, then the blocks of every
         * later chunk that had been read (in order), then what the fxh_reader had buffered and was not used yet */
        struct iovec *iov = (struct iovec *)calloc((size_t)n_in + 2, sizeof(struct iovec));
        if (!iov) err(1, "out of memory");
        int niov = 0;
        iov[niov].iov_base = fb_chunk->buf + fallback_at;
        iov[niov].iov_len = FXS_HEAD + fb_chunk->block_len - fallback_at;
        niov++;
        for (int64_t q = fb_chunk->seq + 1; q < fb_chunk->seq + 1 + n_in; q++) {
            fxs_chunk *f = NULL;
            for (int i = 0; i < n_in; i++) {
                fxs_chunk *c = &chunks[i];
                if (c == fb_chunk || c->seq != q) continue;
                int in_free = 0;
                for (fxs_chunk *x = s->free_in; x; x = x->next) if (x == c) in_free = 1;
                if (!in_free) f = c;
            }
            if (!f) break;
            iov[niov].iov_base = f->buf + FXS_HEAD; iov[niov].iov_len = f->block_len; niov++;
        }
        if (s->mem_pos < s->mem_len) { iov[niov].iov_base = (void *)(s->mem + s->mem_pos); iov[niov].iov_len = s->mem_len - s->mem_pos; niov++; }
        if (s->regular) lseek(s->fd, s->file_pos, SEEK_SET);
        fxh_reader_restart(rd, iov, niov, s->fd_eof && s->mem_pos >= s->mem_len);
        free(iov);
    } else {
        fxh_reader_restart(rd, NULL, 0, 1);
    }
    for (int i = 0; i < n_in; i++) { fxg_free_pinned(chunks[i].buf); free(chunks[i].lastseq); }
    for (int i = 0; i < n_out; i++) fxg_free_pinned(outs[i]);
    free(chunks); free(outs); free(workers); free(s->done); free(s->free_out);
    pthread_mutex_destroy(&s->mu); pthread_cond_destroy(&s->cv);
    for (int g = 0; g < 64; g++) pthread_mutex_destroy(&s->col_mu[g]);
    job->t_total = fxh_now() - t0;
    if (getenv("FASTX_TIMING"))
        fprintf(stderr, "[timing] stream engine: %lld chunks, %lld records, %d workers on %d GPU(s), total %.3f s (waiting for the GPU path %.3f s, write(2) %.3f s)%s\n",
                (long long)job->chunks, (long long)job->records, nworkers, ngpu, job->t_total, job->t_wait_gpu, job->t_write, fallback ? " -> host path" : "");
    if (getenv("FASTX_PATH_REPORT"))
        fprintf(stderr, "[path] gpu_text_records=%lld fallback=%d\n", (long long)job->records, fallback);
    return fallback;
}
