/*
 * fxh.c — host side of the drop-in FASTX tools (see fxh.h).  Block I/O only: the input is read with read(2)
 * into one large buffer, record boundaries are found with memchr, sequence/quality lines are memcpy'd into
 * pinned SoA slabs, and output is assembled in a large buffer and written with write(2).  The per-byte
 * validation and the per-read transform run on the GPU behind include/fxg.h.
 */
#define _GNU_SOURCE
#include "fxh.h"

#include <err.h>
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * command line — src/libfastx/fastx_args.c:39-143
 * ---------------------------------------------------------------------------------------------- */
static const char *g_input = "-";
static const char *g_output = "-";
static int g_verbose = 0, g_compress = 0, g_q_offset = 33;
static FILE *g_report = NULL;

const char *fxh_input_filename(void) { return g_input; }
const char *fxh_output_filename(void) { return g_output; }
int fxh_verbose(void) { return g_verbose; }
int fxh_compress_output(void) { return g_compress; }
int fxh_q_offset(void) { return g_q_offset; }
FILE *fxh_report_file(void) { return g_report ? g_report : stderr; }

int fxh_parse_cmdline(int argc, char *argv[], const char *program_options, fxh_parse_arg_fn fn, const char *usage)
{
    char opts[128];
    int opt;
    snprintf(opts, sizeof opts, "Q:zhvi:o:%s", program_options);
    g_report = stderr;          /* output defaults to STDOUT, so the report goes to STDERR */
    while ((opt = getopt(argc, argv, opts)) != -1) {
        if (opt != ':' && strchr(program_options, opt) != NULL) {
            if (!fn(optind, opt, optarg)) return 0;
            continue;
        }
        switch (opt) {
        case 'h':
            printf("%s", usage);
            exit(1);
        case 'v': g_verbose = 1; break;
        case 'z': g_compress = 1; break;
        case 'i':
            if (optarg == NULL) errx(1, "[-i] option requires FILENAME argument");
            g_input = optarg;
            break;
        case 'o':
            if (optarg == NULL) errx(1, "[-o] option requires FILENAME argument");
            g_output = optarg;
            g_report = stdout;  /* an output file was named, so the report can use STDOUT */
            break;
        case 'Q':
            if (optarg == NULL) errx(1, "[-Q] option requires VALUE argument");
            g_q_offset = atoi(optarg);
            break;
        default:
            printf("use '-h' for usage information.\n");
            exit(1);
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * GPU helpers
 * ---------------------------------------------------------------------------------------------- */
int fxh_gpu_count(void)
{
    const char *e = getenv("FASTX_GPUS");
    int n = e ? atoi(e) : 1;
    return n > 0 ? (n > 64 ? 64 : n) : 1;
}

/* cuInit() enumerates and initialises every visible GPU (about 0.2 s each on an 8-GPU box): before the first CUDA call,
 * narrow the visible set to the GPUs this process will use.  Device numbers are then relative to FASTX_GPU. */
static int g_dev_shift = 0, g_narrowed = 0;
static void narrow_visible_devices(void)
{
    if (g_narrowed) return;
    g_narrowed = 1;
    if (getenv("CUDA_VISIBLE_DEVICES") || getenv("FASTX_KEEP_VISIBLE")) return;
    const int first = getenv("FASTX_GPU") ? atoi(getenv("FASTX_GPU")) : 0, n = fxh_gpu_count();
    if (first < 0) return;
    char list[512]; size_t o = 0;
    for (int g = 0; g < n && o + 8 < sizeof list; g++) o += (size_t)snprintf(list + o, sizeof list - o, "%s%d", g ? "," : "", first + g);
    setenv("CUDA_VISIBLE_DEVICES", list, 1);
    g_dev_shift = first;
}

/* CUDA ordinal of the first GPU this process uses (FASTX_GPU, or 0 once the visible set has been narrowed to start there) */
int fxh_first_device(void)
{
    narrow_visible_devices();
    return (getenv("FASTX_GPU") ? atoi(getenv("FASTX_GPU")) : 0) - g_dev_shift;
}

fxg_ctx *fxh_gpu_open_dev(int dev)       /* dev: CUDA ordinal, fxh_first_device() + k */
{
    narrow_visible_devices();
    fxg_ctx *ctx = NULL;
    int rc = fxg_init(dev, &ctx);
    if (rc != FXG_OK) errx(1, "GPU %d unavailable: %s (%s). This build has no CPU fallback.", dev, fxg_strerror(rc), fxg_last_error(NULL));
    return ctx;
}

fxg_ctx *fxh_gpu_open(void)
{
    int dev = fxh_first_device();
    fxg_ctx *ctx = NULL;
    int rc = fxg_init(dev, &ctx);
    if (rc != FXG_OK) errx(1, "GPU %d unavailable: %s (%s). This build has no CPU fallback.", dev, fxg_strerror(rc), fxg_last_error(NULL));
    return ctx;
}

void fxh_gpu_check(fxg_ctx *ctx, int rc, const char *what)
{
    if (rc != FXG_OK) errx(1, "%s failed: %s (%s)", what, fxg_strerror(rc), fxg_last_error(ctx));
}

int64_t fxh_batch_reads(void)
{
    const char *e = getenv("FASTX_BATCH_READS");
    long long v = e ? atoll(e) : 0;
    return v > 0 ? v : 2000000;
}

/* ------------------------------------------------------------------------------------------------
 * reader
 * ---------------------------------------------------------------------------------------------- */
struct fxh_reader {
    int fd;
    char filename[4096];
    int fastq, q_offset, stale_rows;
    char *buf;
    size_t cap, len, pos;
    int eof;
    uint64_t line_no;                 /* lines consumed so far */
    fxh_batch b;
    size_t slab_bytes;                /* bytes allocated for each of seq / qual */
    int64_t meta_cap;
    char pending[1024];
    int has_pending;
    size_t n_seq, n_reads;
    uint8_t *shadow;                  /* clipper: the aligner's query buffer as the reference leaves it */
    int wmax;
    int64_t next_index;
};

static int is_base(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }

static int all_bases(const char *s, size_t n)
{
    for (size_t i = 0; i < n; i++) if (!is_base((unsigned char)s[i])) return 0;
    return 1;
}

static void refill(fxh_reader *r, size_t keep_from)
{
    /* move the unparsed tail to the front, then read more */
    size_t tail = r->len - keep_from;
    if (keep_from > 0) { memmove(r->buf, r->buf + keep_from, tail); r->len = tail; r->pos -= keep_from; }
    if (r->len + 4096 > r->cap) {               /* one record larger than the window: grow it */
        size_t ncap = r->cap * 2;
        char *nb = (char *)realloc(r->buf, ncap + 1);
        if (!nb) err(1, "out of memory (input buffer)");
        r->buf = nb; r->cap = ncap;
    }
    while (!r->eof && r->len < r->cap) {
        ssize_t k = read(r->fd, r->buf + r->len, r->cap - r->len);
        if (k < 0) { if (errno == EINTR) continue; err(1, "failed to read input file '%s'", r->filename); }
        if (k == 0) { r->eof = 1; break; }
        r->len += (size_t)k;
        if (r->len >= (r->cap >> 1)) break;    /* enough for now */
    }
}

fxh_reader *fxh_reader_open(const char *filename, int allowed, int q_offset, int stale_rows)
{
    fxh_reader *r = (fxh_reader *)calloc(1, sizeof *r);
    if (!r) err(1, "out of memory");
    if (strncmp(filename, "-", 1) == 0) r->fd = STDIN_FILENO;     /* fastx.c:180-181 */
    else {
        r->fd = open(filename, O_RDONLY);
        if (r->fd < 0) err(1, "failed to open input file '%s'", filename);
    }
    strncpy(r->filename, filename, sizeof(r->filename) - 1);
    r->q_offset = q_offset;
    r->stale_rows = stale_rows;
    r->cap = (size_t)256 << 20;
    if (getenv("FASTX_WINDOW_BYTES")) {         /* testing knob: small windows exercise the refill / carry-over logic */
        long long v = atoll(getenv("FASTX_WINDOW_BYTES"));
        if (v >= 65536) r->cap = (size_t)v;
    }
    {   /* a regular file smaller than the default window needs no more than its own size */
        struct stat sb;
        if (r->fd != STDIN_FILENO && fstat(r->fd, &sb) == 0 && S_ISREG(sb.st_mode) && (size_t)sb.st_size + (2u << 20) < r->cap)
            r->cap = (size_t)sb.st_size + (2u << 20);
    }
    r->buf = (char *)malloc(r->cap + 1);
    if (!r->buf) err(1, "out of memory (input buffer)");
    refill(r, 0);
    /* detect_input_format, fastx.c:86-116 */
    if (r->len == 0) errx(1, "Premature End-Of-File (filename ='%s')", r->filename);
    int c = (unsigned char)r->buf[0];
    if (c == '>') {
        if (allowed == FXH_FASTQ_ONLY) errx(1, "input file (%s) is FASTA, but only FASTQ input is allowed.", r->filename);
        r->fastq = 0;
    } else if (c == '@') {
        if (allowed == FXH_FASTA_ONLY) errx(1, "input file (%s) is FASTQ, but only FASTA input is allowed.", r->filename);
        r->fastq = 1;
    } else
        errx(1, "input file (%s) has unknown file format (not FASTA or FASTQ), first character = %c (%d)", r->filename, c, c);
    return r;
}

int fxh_reader_is_fastq(const fxh_reader *r) { return r->fastq; }
size_t fxh_num_input_sequences(const fxh_reader *r) { return r->n_seq; }
size_t fxh_num_input_reads(const fxh_reader *r) { return r->n_reads; }

/* One line starting at r->pos: returns 1 and sets start/len (chomped at CR/LF), 0 if no more data at all,
 * -1 if the line is incomplete and more input may come. */
static int next_line(fxh_reader *r, char **start, size_t *n)
{
    if (r->pos >= r->len) return r->eof ? 0 : -1;
    char *s = r->buf + r->pos;
    char *nl = (char *)memchr(s, '\n', r->len - r->pos);
    size_t raw;
    if (nl) raw = (size_t)(nl - s) + 1;
    else if (r->eof) raw = r->len - r->pos;        /* last line without a newline */
    else return -1;
    size_t l = nl ? raw - 1 : raw;
    {   /* chomp(): the line ends at its FIRST CR or LF (src/libfastx/chomp.c:34-44), an interior CR included */
        const char *cr = (const char *)memchr(s, '\r', l);
        if (cr) l = (size_t)(cr - s);
    }
    if (l >= FXH_MAX_LINE - 1)
        errx(1, "line %llu is longer than %d characters (the reference's fgets() buffer); not supported",
             (unsigned long long)(r->line_no + 1), FXH_MAX_LINE - 2);
    *start = s; *n = l;
    r->pos += raw;
    r->line_no++;
    return 1;
}

static void grow_meta(fxh_reader *r, int64_t cap)
{
    if (cap <= r->meta_cap) return;
    fxh_batch *b = &r->b;
    b->len = (int32_t *)realloc(b->len, (size_t)cap * sizeof(int32_t));
    b->width = (int32_t *)realloc(b->width, (size_t)cap * sizeof(int32_t));
    b->weight = (int32_t *)realloc(b->weight, (size_t)cap * sizeof(int32_t));
    b->name = (const char **)realloc((void *)b->name, (size_t)cap * sizeof(char *));
    b->name2 = (const char **)realloc((void *)b->name2, (size_t)cap * sizeof(char *));
    b->name_len = (int32_t *)realloc(b->name_len, (size_t)cap * sizeof(int32_t));
    b->name2_len = (int32_t *)realloc(b->name2_len, (size_t)cap * sizeof(int32_t));
    b->line_no = (uint64_t *)realloc(b->line_no, (size_t)cap * sizeof(uint64_t));
    if (!b->len || !b->width || !b->weight || !b->name || !b->name2 || !b->name_len || !b->name2_len || !b->line_no)
        err(1, "out of memory (batch tables)");
    r->meta_cap = cap;
}

static void grow_slabs(fxh_reader *r, size_t bytes)
{
    if (bytes <= r->slab_bytes) return;
    fxh_batch *b = &r->b;
    uint8_t *ns = (uint8_t *)fxg_alloc_pinned(bytes), *nq = (uint8_t *)fxg_alloc_pinned(bytes);
    if (!ns || !nq) errx(1, "cannot allocate %zu bytes of pinned host memory", bytes);
    if (b->n > 0) { memcpy(ns, b->seq, (size_t)b->n * b->stride); memcpy(nq, b->qual, (size_t)b->n * b->stride); }
    fxg_free_pinned(b->seq); fxg_free_pinned(b->qual);
    b->seq = ns; b->qual = nq;
    r->slab_bytes = bytes;
}

/* get_reads_count, fastx.c:475-497 (FASTA only) */
static int reads_count(const fxh_reader *r, const char *name, size_t n)
{
    if (r->fastq) return 1;
    const char *dash = (const char *)memchr(name, '-', n);
    if (!dash) return 1;
    char tmp[32];
    size_t m = n - (size_t)(dash + 1 - name);
    if (m >= sizeof tmp) m = sizeof tmp - 1;
    memcpy(tmp, dash + 1, m); tmp[m] = 0;
    int c = atoi(tmp);
    return c > 0 ? c : 1;
}

#define PEND(r, ...) do { snprintf((r)->pending, sizeof((r)->pending), __VA_ARGS__); (r)->has_pending = 1; } while (0)

/* numeric quality line -> bytes (value + 33); mirrors convert_numeric_quality_score_line, fastx.c:137-167 */
static int parse_numeric_qual(fxh_reader *r, char *line, size_t n, size_t nbases, uint8_t *dst, uint64_t line_no)
{
    char saved = line[n];
    line[n] = 0;
    size_t index = 0;
    const char *tok = line;
    char *endp;
    int ok = 1;
    do {
        long v = strtol(tok, &endp, 10);
        if (endp == tok) { PEND(r, "Error: invalid quality score data on line %llu (quality_tok = \"%s\"", (unsigned long long)line_no, tok); ok = 0; break; }
        if (v > 93 || v < -15) { PEND(r, "invalid quality score value (%d) in line %llu.", (int)v, (unsigned long long)line_no); ok = 0; break; }
        if (index < nbases) dst[index] = (uint8_t)(v + 33);
        index++;
        tok = endp;
    } while (*tok != '\0');
    line[n] = saved;
    if (ok && index != nbases) {
        PEND(r, "number of quality values (%zu) doesn't match number of nucleotides (%zu) on line %llu", index, nbases, (unsigned long long)line_no);
        ok = 0;
    }
    return ok;
}

fxh_batch *fxh_reader_next(fxh_reader *r, int64_t max_reads)
{
    fxh_batch *b = &r->b;
    if (r->has_pending) errx(1, "%s", r->pending);
    b->n = 0;
    b->first_index = r->next_index;
    b->numeric_qual = -1;          /* undecided until the first FASTQ record */
    if (max_reads < 1) max_reads = 1;
    grow_meta(r, max_reads);
    if (b->stride == 0) b->stride = 16;
    b->cap = 0;                    /* rows the slabs can hold; fixed when the first record is seen */

    while (b->n < max_reads && (b->n == 0 || b->n < b->cap)) {
        const size_t rec_pos = r->pos;
        const uint64_t rec_line = r->line_no;
        char *l1, *l2, *l3 = NULL, *l4 = NULL;
        size_t n1, n2, n3 = 0, n4 = 0;
        int rc = next_line(r, &l1, &n1);
        if (rc == 0) break;                                     /* end of input */
        int need_more = (rc < 0);
        int got2 = 0, got3 = 0, got4 = 0;
        if (!need_more) { rc = next_line(r, &l2, &n2); if (rc < 0) need_more = 1; else got2 = rc; }
        if (!need_more && got2 && r->fastq) {
            rc = next_line(r, &l3, &n3); if (rc < 0) need_more = 1; else got3 = rc;
            if (!need_more && got3) { rc = next_line(r, &l4, &n4); if (rc < 0) need_more = 1; else got4 = rc; }
        }
        if (need_more) {                                        /* record straddles the buffer end */
            r->pos = rec_pos; r->line_no = rec_line;
            if (b->n > 0) break;                                /* hand out what we have; names point into buf */
            refill(r, rec_pos);
            continue;
        }
        const uint64_t ln1 = rec_line + 1;
        /* line 1: prefix check, fastx.c:331-347 */
        if (r->fastq && (n1 == 0 || l1[0] != '@')) {
            PEND(r, "Invalid input: expecting FASTQ prefix character '@' on line %llu. Is this a valid FASTQ file?\n", (unsigned long long)ln1);
            break;
        }
        if (!r->fastq && (n1 == 0 || l1[0] != '>')) {
            if (n1 > 0 && all_bases(l1, n1))     /* a blank line is no nucleotide string (fastx.c:338-346 tests the first character) */
                PEND(r, "Invalid input: This looks like a multi-line FASTA file.\nLine %llu contains a nucleotides string instead of a '>' prefix.\n"
                        "FASTX-Toolkit can't handle multi-line FASTA files.\nPlease use the FASTA-Formatter tool to convert this file into a single-line FASTA.\n",
                     (unsigned long long)ln1);
            else
                PEND(r, "Invalid input: expecting FASTA prefix character '>' on line %llu. Is this a valid FASTA file?\n", (unsigned long long)ln1);
            break;
        }
        if (!got2) { PEND(r, "Failed to read complete record, missing 2nd line (nucleotides), on line %llu\n", (unsigned long long)(ln1 + 1)); break; }
        if (n2 == 0) { PEND(r, "found empty nucleotide sequence on line %llu\n", (unsigned long long)(ln1 + 1)); break; }

        int numeric = 0;
        if (r->fastq) {
            /* later lines broken: the reference has already validated the bases of THIS record by then */
            if (!got3 || !got4) {
                if (!all_bases(l2, n2)) PEND(r, "found invalid nucleotide sequence (%.*s) on line %llu\n", (int)n2, l2, (unsigned long long)(ln1 + 1));
                else if (!got3) PEND(r, "Failed to read complete record, missing 3rd line (name-2), on line %llu\n", (unsigned long long)(ln1 + 2));
                else PEND(r, "Failed to read complete record, missing 4th line (quality), on line %llu\n", (unsigned long long)(ln1 + 3));
                break;
            }
            numeric = (n4 != n2);                               /* fastx.c:382-390 */
            if (b->numeric_qual < 0) b->numeric_qual = numeric;
            else if (b->numeric_qual != numeric) {              /* keep one quality encoding per batch */
                r->pos = rec_pos; r->line_no = rec_line;
                break;
            }
        }
        /* stride: grow (only on an empty batch) so that every row fits */
        int need = (int)((n2 + 15) & ~(size_t)15);
        if (r->stale_rows && r->wmax > (int)n2) need = (r->wmax + 15) & ~15;
        if (need > b->stride) {
            if (b->n > 0) { r->pos = rec_pos; r->line_no = rec_line; break; }
            b->stride = need;
        }
        if (b->n == 0) {
            /* size the slabs for what the text buffer can still hold (small inputs stay small) */
            int64_t est = (int64_t)((r->len - rec_pos) / (n2 + 2)) + 16;
            if (est > max_reads) est = max_reads;
            grow_slabs(r, (size_t)est * (size_t)b->stride);
            b->cap = (int64_t)(r->slab_bytes / (size_t)b->stride);
            if (b->cap > max_reads) b->cap = max_reads;
        }
        const int64_t i = b->n;
        uint8_t *srow = b->seq + (size_t)i * b->stride, *qrow = b->qual + (size_t)i * b->stride;
        if (r->stale_rows) {
            /* the aligner's std::string keeps its old bytes beyond the new contents */
            if (!r->shadow) { r->shadow = (uint8_t *)calloc(1, FXH_MAX_LINE + 16); if (!r->shadow) err(1, "out of memory"); }
            memcpy(r->shadow, l2, n2);
            r->shadow[n2] = 0;
            if ((int)n2 > r->wmax) r->wmax = (int)n2;
            memcpy(srow, r->shadow, (size_t)r->wmax);
            b->width[i] = r->wmax;
        } else {
            memcpy(srow, l2, n2);
            b->width[i] = (int32_t)n2;
        }
        if (r->fastq) {
            if (!numeric) memcpy(qrow, l4, n4);
            else if (!parse_numeric_qual(r, l4, n4, n2, qrow, ln1 + 3)) {
                if (!all_bases(l2, n2)) PEND(r, "found invalid nucleotide sequence (%.*s) on line %llu\n", (int)n2, l2, (unsigned long long)(ln1 + 1));
                break;
            }
            b->name2[i] = l3 + (n3 > 0 ? 1 : 0);
            b->name2_len[i] = (int32_t)(n3 > 0 ? n3 - 1 : 0);
        } else { b->name2[i] = NULL; b->name2_len[i] = 0; }
        b->len[i] = (int32_t)n2;
        b->name[i] = l1 + 1;
        b->name_len[i] = (int32_t)(n1 - 1);
        b->line_no[i] = ln1;
        b->weight[i] = reads_count(r, l1 + 1, n1 - 1);
        r->n_seq++;
        r->n_reads += (size_t)b->weight[i];
        b->n++;
    }
    if (b->numeric_qual < 0) b->numeric_qual = 0;
    r->next_index += b->n;
    if (b->n == 0) {
        if (r->has_pending) errx(1, "%s", r->pending);
        return NULL;
    }
    return b;
}

size_t fxh_reader_raw(fxh_reader *r, char **p)
{
    if (!r->eof && (r->len - r->pos) < (r->cap >> 1)) refill(r, r->pos);
    *p = r->buf + r->pos;
    return r->len - r->pos;
}

void fxh_reader_consume(fxh_reader *r, size_t bytes, int64_t records)
{
    r->pos += bytes;
    r->line_no += 4ull * (uint64_t)records;
    r->n_seq += (size_t)records;
    r->n_reads += (size_t)records;
    r->next_index += records;
}

/* ---- hooks for the streaming engine (fxh_stream.c) ---- */
int fxh_reader_fd(const fxh_reader *r) { return r->fd; }

/* the unparsed bytes the reader holds (they stay where they are until fxh_reader_restart) and whether the input ended there */
void fxh_reader_detach(fxh_reader *r, char **p, size_t *len, int *eof)
{
    *p = r->buf + r->pos;
    *len = r->len - r->pos;
    *eof = r->eof;
}

/* records the engine consumed on the reader's behalf */
void fxh_reader_account(fxh_reader *r, int64_t records, int64_t reads, int lines_per_record)
{
    r->line_no += (uint64_t)lines_per_record * (uint64_t)records;
    r->n_seq += (size_t)records;
    r->n_reads += (size_t)reads;
    r->next_index += records;
}

/* the reader continues with these bytes (in order), then with whatever its file descriptor still delivers */
void fxh_reader_restart(fxh_reader *r, const struct iovec *iov, int niov, int eof)
{
    size_t total = 0;
    for (int i = 0; i < niov; i++) total += iov[i].iov_len;
    size_t cap = r->cap;
    while (cap < 2 * total + 4096) cap *= 2;
    char *nb = (char *)malloc(cap + 1);
    if (!nb) err(1, "out of memory (input buffer)");
    size_t off = 0;
    for (int i = 0; i < niov; i++) { memcpy(nb + off, iov[i].iov_base, iov[i].iov_len); off += iov[i].iov_len; }
    free(r->buf);
    r->buf = nb; r->cap = cap; r->len = total; r->pos = 0; r->eof = eof;
}

void fxh_reader_pin(fxh_reader *r)
{
    if (r->cap >= ((size_t)32 << 20)) (void)fxg_host_register(r->buf, r->cap + 1);   /* small inputs: not worth page-locking */
}

double fxh_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void fxh_reader_seed_shadow(fxh_reader *r, const char *last_seq, int len)
{
    if (!r->stale_rows || len <= 0 || len > FXH_MAX_LINE) return;
    if (!r->shadow) { r->shadow = (uint8_t *)calloc(1, FXH_MAX_LINE + 16); if (!r->shadow) err(1, "out of memory"); }
    memcpy(r->shadow, last_seq, (size_t)len);
    r->shadow[len] = 0;
    if (len > r->wmax) r->wmax = len;
}

int fxh_reader_at_eof(const fxh_reader *r) { return r->eof; }

size_t fxh_text_chunk_bytes(void)
{
    const char *e = getenv("FASTX_CHUNK_BYTES");       /* testing knob; default 64 MB per GPU text chunk */
    long long v = e ? atoll(e) : 0;
    return v >= 16384 ? (size_t)v : ((size_t)64 << 20);
}

int fxh_text_path_enabled(void)
{
    const char *e = getenv("FASTX_TEXT_PATH");
    return !(e && e[0] == '0');
}

fxg_batch fxh_as_fxg_batch(const fxh_batch *b, int with_qual)
{
    fxg_batch g;
    g.seq = b->seq;
    g.qual = with_qual ? b->qual : NULL;
    g.len = b->len;
    g.uniform_len = 0;
    g.stride = b->stride;
    g.n = b->n;
    return g;
}

/* The GPU says record idx is illegal: word it as the reference does (bases first, then quality values). */
void fxh_die_bad_record(const fxh_reader *r, const fxh_batch *b, int64_t idx)
{
    const uint8_t *s = b->seq + (size_t)idx * b->stride, *q = b->qual + (size_t)idx * b->stride;
    const int L = b->len[idx];
    for (int i = 0; i < L; i++)
        if (!is_base(s[i]))
            errx(1, "found invalid nucleotide sequence (%.*s) on line %llu\n", L, (const char *)s, (unsigned long long)(b->line_no[idx] + 1));
    if (r->fastq && !b->numeric_qual) {
        for (int i = 0; i < L; i++) {
            const int c = (int)(signed char)q[i], v = c - r->q_offset;
            if (v < -15 || v > 93)
                errx(1, "Invalid quality score value (char '%c' ord %d quality value %d) on line %llu", c, c, v,
                     (unsigned long long)(b->line_no[idx] + 3));
        }
    }
    errx(1, "internal error: GPU flagged record %lld (line %llu) but the host finds it valid", (long long)(b->first_index + idx),
         (unsigned long long)b->line_no[idx]);
}

/* ------------------------------------------------------------------------------------------------
 * writer — fastx.c:193-312 (open / gzip child), :406-473 (record)
 * ---------------------------------------------------------------------------------------------- */
struct fxh_writer {
    int fd;
    int closed;
    pid_t gzip_pid;
    /* background write(2) of one already-formatted block at a time (the GPU text path alternates two blocks) */
    pthread_t aw_thread;
    pthread_mutex_t aw_mu;
    pthread_cond_t aw_cv;
    int aw_started, aw_busy, aw_quit;
    const char *aw_buf;
    size_t aw_bytes;
    int fastq;
    int regular, store_threads;       /* output is a regular file (no gzip child): large blocks may be stored through a mapping */
    int gz;                           /* -z without a gzip child: this writer frames the gzip stream itself */
    uint32_t gz_crc; uint64_t gz_len; /* pure CRC-32 register and length of the uncompressed stream so far */
    char *buf;
    size_t cap, len;
    size_t n_seq, n_reads;
};

/* ---- -z: gzip framing (RFC 1952) around DEFLATE blocks that are either made by the GPU (fxg_text_set_deflate) or, for text
 * formatted on the host, plain stored blocks.  The reference pipes its text through a forked gzip (fastx.c:214-248); what
 * matters to a reader of the file is the uncompressed stream, which is the same. ---- */
#define GZ_POLY 0xEDB88320u
static uint32_t gz_tab[256];
static void gz_tab_init(void)
{
    if (gz_tab[1]) return;
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ GZ_POLY : c >> 1; gz_tab[i] = c; }
}
static uint32_t gz_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0u) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ GZ_POLY : b >> 1;
    }
    return p;
}
static uint32_t gz_shift(uint32_t crc, uint64_t nbytes)      /* crc * x^(8 nbytes) mod P: the register after nbytes more zero bytes */
{
    uint32_t sq = 1u << 30, p = 1u << 31;                    /* x^1, x^0 */
    for (int k = 0; k < 3; k++) sq = gz_multmodp(sq, sq);    /* x^8 */
    while (nbytes) { if (nbytes & 1) p = gz_multmodp(sq, p); sq = gz_multmodp(sq, sq); nbytes >>= 1; }
    return gz_multmodp(p, crc);
}

static int open_output_file(const char *filename)
{
    if (strncmp(filename, "-", 6) == 0) return STDOUT_FILENO;
    int fd = open(filename, O_CREAT | O_WRONLY | O_TRUNC, 0666);
    if (fd == -1) err(1, "Failed to create output file (%s)", filename);
    return fd;
}

static int open_output_compressor(const char *filename, pid_t *pid)
{
    int p[2];
    if (pipe(p) != 0) err(1, "pipe (for gzip) failed");
    pid_t child = fork();
    if (child > 0) { close(p[0]); *pid = child; return p[1]; }
    dup2(p[0], STDIN_FILENO);
    close(p[1]);
    int fd = open_output_file(filename);
    dup2(fd, STDOUT_FILENO);
    execlp("gzip", "gzip", (char *)NULL);
    err(1, "execlp(gzip) failed");
    return 0;
}

/* errx()/exit() must not lose buffered records: the reference's stdio buffers are flushed at exit too */
static fxh_writer *g_open_writer = NULL;
static void flush_at_exit(void)
{
    if (g_open_writer) fxh_writer_close(g_open_writer);
}

fxh_writer *fxh_writer_open(const char *filename, int fastq, int compress)
{
    fxh_writer *w = (fxh_writer *)calloc(1, sizeof *w);
    if (!w) err(1, "out of memory");
    if (!g_open_writer) atexit(flush_at_exit);
    g_open_writer = w;
    if (compress && !getenv("FASTX_GZIP_CHILD")) {
        static const unsigned char hdr[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3 };
        w->fd = open_output_file(filename);
        w->gz = 1;
        gz_tab_init();
        if (write(w->fd, hdr, sizeof hdr) != (ssize_t)sizeof hdr) err(1, "writing nucleotides failed");
    } else
        w->fd = compress ? open_output_compressor(filename, &w->gzip_pid) : open_output_file(filename);
    w->fastq = fastq;
    {
        struct stat sb;
        w->regular = !compress && fstat(w->fd, &sb) == 0 && S_ISREG(sb.st_mode);
        const char *e = getenv("FASTX_WRITE_THREADS");
        w->store_threads = e ? atoi(e) : 4;
    }
    w->cap = (size_t)16 << 20;
    w->buf = (char *)malloc(w->cap);
    if (!w->buf) err(1, "out of memory (output buffer)");
    return w;
}

static void aw_wait(fxh_writer *w);
static void write_all(int fd, const char *text, size_t bytes);

/* -z framing: host-formatted text goes out as stored DEFLATE blocks (at most 65 535 bytes each), byte aligned */
static void gz_write_stored(fxh_writer *w, const char *text, size_t bytes)
{
    uint32_t reg = 0;
    for (size_t i = 0; i < bytes; i++) reg = gz_tab[(reg ^ (unsigned char)text[i]) & 0xFFu] ^ (reg >> 8);
    w->gz_crc = gz_shift(w->gz_crc, bytes) ^ reg;
    w->gz_len += bytes;
    for (size_t off = 0; off < bytes; off += 65535) {
        const size_t n = bytes - off < 65535 ? bytes - off : 65535;
        const unsigned char h[5] = { 0, (unsigned char)(n & 0xFF), (unsigned char)(n >> 8), (unsigned char)(~n & 0xFF), (unsigned char)((~n >> 8) & 0xFF) };
        write_all(w->fd, (const char *)h, 5);
        write_all(w->fd, text + off, n);
    }
}

static void writer_flush(fxh_writer *w)
{
    aw_wait(w);          /* keep the byte order: a background block goes out before anything buffered later */
    if (w->gz) { if (w->len) gz_write_stored(w, w->buf, w->len); w->len = 0; return; }
    size_t off = 0;
    while (off < w->len) {
        ssize_t k = write(w->fd, w->buf + off, w->len - off);
        if (k < 0) { if (errno == EINTR) continue; err(1, "writing nucleotides failed"); }
        off += (size_t)k;
    }
    w->len = 0;
}

static inline void writer_room(fxh_writer *w, size_t n)
{
    if (w->len + n > w->cap) {
        writer_flush(w);
        if (n > w->cap) { w->cap = n * 2; w->buf = (char *)realloc(w->buf, w->cap); if (!w->buf) err(1, "out of memory (output buffer)"); }
    }
}

void fxh_write_record(fxh_writer *w, const fxh_batch *b, int64_t i, const uint8_t *seq_row, const uint8_t *qual_row, int32_t out_len)
{
    fxh_write_record_named(w, b, i, seq_row, qual_row, out_len, b->name[i], b->name_len[i]);
}

void fxh_write_record_named(fxh_writer *w, const fxh_batch *b, int64_t i, const uint8_t *seq_row, const uint8_t *qual_row, int32_t out_len,
                            const char *name, int32_t name_len)
{
    const size_t nl = (size_t)name_len, n2l = (size_t)b->name2_len[i], L = (size_t)out_len;
    writer_room(w, nl + n2l + L * (b->numeric_qual ? 5 : 1) + L + 16);
    char *p = w->buf + w->len;
    *p++ = w->fastq ? '@' : '>';
    memcpy(p, name, nl); p += nl; *p++ = '\n';
    memcpy(p, seq_row, L); p += L; *p++ = '\n';
    if (w->fastq) {
        *p++ = '+';
        memcpy(p, b->name2[i], n2l); p += n2l; *p++ = '\n';
        if (!b->numeric_qual) { memcpy(p, qual_row, L); p += L; }
        else {
            for (size_t k = 0; k < L; k++) {                 /* write_numeric_qual_string, fastx.c:421-438 */
                p += sprintf(p, "%d", (int)qual_row[k] - 33);
                if (k + 1 < L) *p++ = ' ';
            }
        }
        *p++ = '\n';
    }
    w->len = (size_t)(p - w->buf);
    w->n_seq++;
    w->n_reads += (size_t)b->weight[i];
}

static void write_all(int fd, const char *text, size_t bytes)
{
    size_t off = 0;
    while (off < bytes) {
        ssize_t k = write(fd, text + off, bytes - off);
        if (k < 0) { if (errno == EINTR) continue; err(1, "writing nucleotides failed"); }
        off += (size_t)k;
    }
}

static void *aw_main(void *arg)
{
    fxh_writer *w = (fxh_writer *)arg;
    pthread_mutex_lock(&w->aw_mu);
    for (;;) {
        while (!w->aw_busy && !w->aw_quit) pthread_cond_wait(&w->aw_cv, &w->aw_mu);
        if (w->aw_busy) {
            const char *b = w->aw_buf; size_t n = w->aw_bytes;
            pthread_mutex_unlock(&w->aw_mu);
            write_all(w->fd, b, n);
            pthread_mutex_lock(&w->aw_mu);
            w->aw_busy = 0;
            pthread_cond_broadcast(&w->aw_cv);
        } else if (w->aw_quit) break;
    }
    pthread_mutex_unlock(&w->aw_mu);
    return NULL;
}

/* wait until the background block (if any) is on its way to the kernel */
static void aw_wait(fxh_writer *w)
{
    if (!w->aw_started) return;
    pthread_mutex_lock(&w->aw_mu);
    while (w->aw_busy) pthread_cond_wait(&w->aw_cv, &w->aw_mu);
    pthread_mutex_unlock(&w->aw_mu);
}

/* Formatted text produced elsewhere (the GPU): written by a background thread so that the next chunk's read + GPU
 * work overlaps this write.  `text` must stay untouched until the NEXT fxh_write_raw() call returns. */
void fxh_write_raw(fxh_writer *w, const char *text, size_t bytes, int64_t records)
{
    writer_flush(w);
    if (w->gz) { if (bytes) gz_write_stored(w, text, bytes); w->n_seq += (size_t)records; w->n_reads += (size_t)records; return; }
    if (!w->aw_started) {
        pthread_mutex_init(&w->aw_mu, NULL);
        pthread_cond_init(&w->aw_cv, NULL);
        if (pthread_create(&w->aw_thread, NULL, aw_main, w) != 0) err(1, "pthread_create");
        w->aw_started = 1;
    }
    aw_wait(w);
    pthread_mutex_lock(&w->aw_mu);
    w->aw_buf = text; w->aw_bytes = bytes; w->aw_busy = 1;
    pthread_cond_broadcast(&w->aw_cv);
    pthread_mutex_unlock(&w->aw_mu);
    w->n_seq += (size_t)records;
    w->n_reads += (size_t)records;
}

/* Large blocks of formatted text into a REGULAR output file: write(2) copies with one thread and holds the inode lock, so the
 * block is stored through a shared mapping by several threads at once instead (page faults and copies run in parallel); the
 * file is grown to exactly the bytes stored and the descriptor's offset moved behind them, so write(2) may follow. */
typedef struct { char *dst; const char *src; size_t n; } store_job;
static void *store_main(void *arg) { store_job *j = (store_job *)arg; memcpy(j->dst, j->src, j->n); return NULL; }

static int store_parallel(fxh_writer *w, const char *text, size_t bytes)
{
    if (!w->regular || w->store_threads < 2 || bytes < ((size_t)4 << 20)) return 0;
    const off_t off = lseek(w->fd, 0, SEEK_CUR);
    if (off < 0) return 0;
    const long pg = sysconf(_SC_PAGESIZE);
    const off_t aoff = off & ~((off_t)pg - 1);
    const size_t lead = (size_t)(off - aoff);
    if (ftruncate(w->fd, off + (off_t)bytes) != 0) return 0;
    char *map = (char *)mmap(NULL, lead + bytes, PROT_READ | PROT_WRITE, MAP_SHARED, w->fd, aoff);
    if (map == MAP_FAILED) return 0;
    const int T = w->store_threads > 16 ? 16 : w->store_threads;
    store_job jobs[16]; pthread_t th[16];
    const size_t part = ((bytes / (size_t)T) + 4095) & ~(size_t)4095;
    int nt = 0;
    for (size_t o = 0; o < bytes && nt < T; o += part, nt++) {
        jobs[nt].dst = map + lead + o; jobs[nt].src = text + o; jobs[nt].n = (bytes - o < part) ? bytes - o : part;
        if (nt > 0 && pthread_create(&th[nt], NULL, store_main, &jobs[nt]) != 0) err(1, "pthread_create");
    }
    store_main(&jobs[0]);
    for (int i = 1; i < nt; i++) pthread_join(th[i], NULL);
    munmap(map, lead + bytes);
    if (lseek(w->fd, off + (off_t)bytes, SEEK_SET) < 0) err(1, "lseek on the output file failed");
    return 1;
}

/* formatted text, written before the call returns (the streaming engine's own thread is the background here) */
void fxh_writer_write_now(fxh_writer *w, const char *text, size_t bytes, int64_t records, int64_t reads)
{
    writer_flush(w);
    if (w->gz) gz_write_stored(w, text, bytes);
    else if (!store_parallel(w, text, bytes)) write_all(w->fd, text, bytes);
    w->n_seq += (size_t)records;
    w->n_reads += (size_t)reads;
}

int fxh_writer_frames_gzip(const fxh_writer *w) { return w->gz; }

/* -z: DEFLATE blocks made by the GPU for `raw_len` bytes of text whose pure CRC-32 register is `crc_pure` */
void fxh_writer_write_deflated(fxh_writer *w, const char *blocks, size_t bytes, uint64_t raw_len, uint32_t crc_pure, int64_t records, int64_t reads)
{
    writer_flush(w);
    write_all(w->fd, blocks, bytes);
    w->gz_crc = gz_shift(w->gz_crc, raw_len) ^ crc_pure;
    w->gz_len += raw_len;
    w->n_seq += (size_t)records;
    w->n_reads += (size_t)reads;
}

void fxh_writer_close(fxh_writer *w)
{
    if (w->closed) return;
    w->closed = 1;
    if (g_open_writer == w) g_open_writer = NULL;
    writer_flush(w);
    if (w->aw_started) {
        pthread_mutex_lock(&w->aw_mu);
        w->aw_quit = 1;
        pthread_cond_broadcast(&w->aw_cv);
        pthread_mutex_unlock(&w->aw_mu);
        pthread_join(w->aw_thread, NULL);
        w->aw_started = 0;
    }
    if (w->gz) {          /* final empty stored block, CRC-32 and ISIZE */
        const uint32_t crc = (gz_shift(0xFFFFFFFFu, w->gz_len) ^ w->gz_crc) ^ 0xFFFFFFFFu, isz = (uint32_t)(w->gz_len & 0xFFFFFFFFu);
        const unsigned char t[13] = { 1, 0, 0, 0xFF, 0xFF, (unsigned char)crc, (unsigned char)(crc >> 8), (unsigned char)(crc >> 16), (unsigned char)(crc >> 24),
                                      (unsigned char)isz, (unsigned char)(isz >> 8), (unsigned char)(isz >> 16), (unsigned char)(isz >> 24) };
        write_all(w->fd, (const char *)t, sizeof t);
    }
    if (w->fd != STDOUT_FILENO) close(w->fd);
    if (w->gzip_pid > 0) { int st; waitpid(w->gzip_pid, &st, 0); }   /* let gzip finish before we exit */
}

size_t fxh_num_output_sequences(const fxh_writer *w) { return w->n_seq; }
size_t fxh_num_output_reads(const fxh_writer *w) { return w->n_reads; }
