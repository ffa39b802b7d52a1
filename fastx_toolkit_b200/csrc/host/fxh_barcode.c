/*
 * fxh_barcode.c — drop-in for the reference's barcode splitter (scripts/fastx_barcode_splitter.pl, a Perl script):
 * same command line (Getopt::Long style), same barcode-file checks and messages, same output files and summary.
 * The matching loop (match_sequences, :208-290) runs on the GPU: the host cuts the fragment each read is compared on
 * (first/last barcode-length characters of the sequence line), K-BARCODE (fxg_barcode_host) returns the winning entry
 * per read, and the host appends the record's lines, verbatim, to that barcode's file.
 *
 * Kept from the script: FASTA/FASTQ auto-detection by the first byte of STDIN (:330-350), two or four LINES per record
 * without any validation, "\n"-only chomp of the sequence line (a CR stays part of the fragment), the order of the
 * entry list (barcode, then its --partial forms, :170-176), first-lowest-count-wins, the summary table sorted by
 * identifier (:300-310), exit status 1 for the usage screen, errno for a barcode file that cannot be opened.
 * Not kept: the wording of --help (this file prints its own summary), Perl's "uninitialized value" warnings, and the
 * exit status of Perl's die() where it leaks an unrelated errno (25 for barcode-file errors): this tool exits 255 there.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include "fxg.h"
#include "fxh.h"

/* ------------------------------------------------------------------------------------------------ options */
static const char *o_bcfile, *o_prefix, *o_suffix = "";
static int o_eol, o_bol, o_exact, o_quiet, o_debug, o_help;
static long o_partial = 0, o_mismatches = 1;

static void die(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fflush(stdout);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    exit(status);
}

static void usage(const char *self)
{
    printf("Barcode Splitter (B200 build of FASTX-Toolkit's fastx_barcode_splitter.pl)\n"
           "\n"
           "Reads FASTA/FASTQ from STDIN (format auto-detected), writes one file per barcode plus an 'unmatched' file,\n"
           "prints a summary to STDOUT.\n"
           "\n"
           "usage: %s --bcfile FILE --prefix PREFIX [--suffix SUFFIX] [--bol|--eol]\n"
           "         [--mismatches N] [--exact] [--partial N] [--help] [--quiet] [--debug]\n"
           "\n"
           "--bcfile FILE    barcode file: one 'identifier<TAB>barcode' per line, '#' starts a comment line\n"
           "--prefix PREFIX  prepended to every output file name (may hold a directory)\n"
           "--suffix SUFFIX  appended to every output file name\n"
           "--bol | --eol    match the barcodes at the beginning / at the end of the sequences (one is required)\n"
           "--mismatches N   mismatches allowed (default 1);  --exact = --mismatches 0\n"
           "--partial N      also try the barcodes shortened by up to N bases (a missing base counts as a mismatch)\n"
           "--quiet          no summary;  --debug  chatter on STDERR;  --help  this screen\n", self);
    exit(1);
}

enum { T_FLAG, T_STR, T_INT };
static const struct { const char *name; int type; void *dst; } OPTS[] = {
    { "bcfile", T_STR, &o_bcfile }, { "eol", T_FLAG, &o_eol }, { "bol", T_FLAG, &o_bol }, { "exact", T_FLAG, &o_exact },
    { "prefix", T_STR, &o_prefix }, { "suffix", T_STR, &o_suffix }, { "quiet", T_FLAG, &o_quiet }, { "partial", T_INT, &o_partial },
    { "debug", T_FLAG, &o_debug }, { "mismatches", T_INT, &o_mismatches }, { "help", T_FLAG, &o_help },
};
#define N_OPTS ((int)(sizeof(OPTS) / sizeof(OPTS[0])))

/* Getopt::Long with its default configuration: "--name" or "-name", case-insensitive, unique abbreviations,
 * "--name=value" or "--name value"; non-option words are skipped, "--" ends the options.  Returns 0 when a word was
 * rejected (the script then runs its own checks and leaves quietly, :140). */
static int get_options(int argc, char **argv)
{
    int ok = 1;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "--")) break;
        if (a[0] != '-' || a[1] == 0) continue;
        const char *name = a + 1 + (a[1] == '-');
        const char *eq = strchr(name, '=');
        const size_t nl = eq ? (size_t)(eq - name) : strlen(name);
        int hit = -1, hits = 0;
        for (int k = 0; k < N_OPTS; k++) {
            if (strlen(OPTS[k].name) == nl && !strncasecmp(OPTS[k].name, name, nl)) { hit = k; hits = 1; break; }
            if (nl > 0 && !strncasecmp(OPTS[k].name, name, nl)) { hit = k; hits++; }
        }
        if (hits == 0 || nl == 0) { fprintf(stderr, "Unknown option: %.*s\n", (int)nl, name); ok = 0; continue; }
        if (hits > 1) {
            const char *cand[N_OPTS];                          /* Getopt::Long lists the candidates sorted */
            int nc = 0;
            for (int k = 0; k < N_OPTS; k++)
                if (!strncasecmp(OPTS[k].name, name, nl)) cand[nc++] = OPTS[k].name;
            for (int x = 0; x < nc; x++)
                for (int y = x + 1; y < nc; y++)
                    if (strcmp(cand[y], cand[x]) < 0) { const char *t = cand[x]; cand[x] = cand[y]; cand[y] = t; }
            fprintf(stderr, "Option %.*s is ambiguous (", (int)nl, name);
            for (int x = 0; x < nc; x++) fprintf(stderr, "%s%s", x ? ", " : "", cand[x]);
            fprintf(stderr, ")\n");
            ok = 0;
            continue;
        }
        if (OPTS[hit].type == T_FLAG) {
            if (eq) { fprintf(stderr, "Option %s does not take an argument\n", OPTS[hit].name); ok = 0; continue; }
            *(int *)OPTS[hit].dst = 1;
            continue;
        }
        const char *val = eq ? eq + 1 : (i + 1 < argc ? argv[++i] : NULL);
        if (!val) { fprintf(stderr, "Option %s requires an argument\n", OPTS[hit].name); ok = 0; continue; }
        if (OPTS[hit].type == T_STR) { *(const char **)OPTS[hit].dst = val; continue; }
        const char *p = val;
        if (*p == '-' || *p == '+') p++;
        int digits = 0;
        while (isdigit((unsigned char)*p)) { p++; digits++; }
        if (!digits || *p) { fprintf(stderr, "Value \"%s\" invalid for option %s (number expected)\n", val, OPTS[hit].name); ok = 0; continue; }
        *(long *)OPTS[hit].dst = strtol(val, NULL, 10);
    }
    return ok;
}

static void parse_command_line(int argc, char **argv)
{
    if (argc <= 1) usage(argv[0]);
    const int ok = get_options(argc, argv);
    if (o_help) usage(argv[0]);
    if (!o_bcfile) die(255, "Error: barcode file not specified (use '--bcfile [FILENAME]')\n");
    if (!o_prefix) die(255, "Error: prefix path/filename not specified (use '--prefix [PATH]')\n");
    if (o_bol == o_eol) {
        if (o_eol) die(255, "Error: can't specify both --eol & --bol\n");
        die(255, "Error: must specify either --eol or --bol\n");
    }
    if (o_partial < 0) die(255, "Error: invalid for value partial matches (valid values are 0 or greater)\n");
    if (o_exact) o_mismatches = 0;
    if (o_mismatches < 0) die(255, "Error: invalid value for mismatches (valid values are 0 or more)\n");
    if (o_partial > o_mismatches)
        die(255, "Error: partial overlap value (%ld) bigger than max. allowed mismatches (%ld)\n", o_partial, o_mismatches);
    if (!ok) exit(0);
}

/* ------------------------------------------------------------------------------------------------ barcode file */
typedef struct { char *name; FILE *f; char *path; uint64_t count; } ident_t;
static ident_t *idents;          /* distinct identifiers, in order of first appearance; "unmatched" included */
static int n_idents, unmatched_ident;
static uint8_t *entries;         /* n_entries x stride, zero padded */
static int32_t *entry_len, *entry_ident;
static int n_entries, cap_entries, barcode_len = -1, stride;

static int ident_index(const char *name)
{
    for (int i = 0; i < n_idents; i++) if (!strcmp(idents[i].name, name)) return i;
    idents = (ident_t *)realloc(idents, (size_t)(n_idents + 1) * sizeof(ident_t));
    memset(&idents[n_idents], 0, sizeof(ident_t));
    idents[n_idents].name = strdup(name);
    return n_idents++;
}

static void push_entry(int ident, const char *bc, int len)
{
    if (n_entries == cap_entries) {
        cap_entries = cap_entries ? cap_entries * 2 : 64;
        entries = (uint8_t *)realloc(entries, (size_t)cap_entries * stride);
        entry_len = (int32_t *)realloc(entry_len, (size_t)cap_entries * sizeof(int32_t));
        entry_ident = (int32_t *)realloc(entry_ident, (size_t)cap_entries * sizeof(int32_t));
    }
    memset(entries + (size_t)n_entries * stride, 0, (size_t)stride);
    memcpy(entries + (size_t)n_entries * stride, bc, (size_t)len);
    entry_len[n_entries] = len;
    entry_ident[n_entries] = ident;
    n_entries++;
}

/* load_barcode_file, :143-190 */
static void load_barcode_file(const char *filename)
{
    FILE *f = fopen(filename, "r");
    if (!f) die(errno ? errno : 255, "Error: failed to open barcode file (%s)\n", filename);
    char *line = NULL;
    size_t cap = 0;
    long lineno = 0;
    while (getline(&line, &cap, f) >= 0) {
        lineno++;
        if (line[0] == '#') continue;
        /* my ($ident, $barcode) = split;   — whitespace-separated words, leading whitespace ignored */
        char *save = NULL;
        char *ident = strtok_r(line, " \t\n\r\f\v", &save);
        char *bc = ident ? strtok_r(NULL, " \t\n\r\f\v", &save) : NULL;
        char empty[1] = "";
        if (!bc) bc = empty;                                   /* uc(undef) is "" */
        if (!ident) ident = empty;
        for (char *p = bc; *p; p++) *p = (char)toupper((unsigned char)*p);
        int good = *bc != 0;
        for (const char *p = bc; *p; p++) if (!strchr("AGCT", *p)) good = 0;
        if (!good) die(255, "Error: bad barcode value (%s) at barcode file (%s) line %ld\n", bc, filename, lineno);
        good = *ident != 0;
        for (const char *p = ident; *p; p++) if (!(isalnum((unsigned char)*p) || *p == '_')) good = 0;
        if (!good) die(255, "Error: bad identifier value (%s) at barcode file (%s) line %ld (must be alphanumeric)\n", ident, filename, lineno);
        int len = (int)strlen(bc);
        if (len <= o_mismatches)
            die(255, "Error: badcode(%s, %s) is shorter or equal to maximum number of mismatches (%ld). This makes no sense. Specify fewer  mismatches.\n",
                ident, bc, o_mismatches);
        if (barcode_len < 0) {
            barcode_len = len;
            stride = (len + 15) & ~15;
            if (stride > 64) die(255, "Error: barcodes longer than 64 characters are not supported by this build\n");
        }
        if (barcode_len != len) die(255, "Error: found barcodes in different lengths. this feature is not supported yet.\n");
        const int id = ident_index(ident);
        push_entry(id, bc, len);
        for (long i = 1; i <= o_partial; i++) {               /* :170-176: drop one more base at the barcode's outer end */
            if (len == 0) { push_entry(id, bc, 0); continue; }
            if (o_bol) memmove(bc, bc + 1, (size_t)len);       /* includes the NUL */
            else bc[len - 1] = 0;
            len--;
            push_entry(id, bc, len);
        }
    }
    free(line);
    fclose(f);
    if (o_debug) {
        fprintf(stderr, "barcode\tsequence\n");
        for (int e = 0; e < n_entries; e++)
            fprintf(stderr, "%s\t%.*s\n", idents[entry_ident[e]].name, entry_len[e], (const char *)entries + (size_t)e * stride);
    }
}

static void create_output_files(void)
{
    unmatched_ident = ident_index("unmatched");
    for (int i = 0; i < n_idents; i++) {
        size_t l = strlen(o_prefix) + strlen(idents[i].name) + strlen(o_suffix) + 1;
        idents[i].path = (char *)malloc(l);
        snprintf(idents[i].path, l, "%s%s%s", o_prefix, idents[i].name, o_suffix);
        idents[i].f = fopen(idents[i].path, "w");
        if (!idents[i].f) die(errno ? errno : 255, "Error: failed to create output file (%s)\n", idents[i].path);
        setvbuf(idents[i].f, NULL, _IOFBF, 1 << 20);
    }
}

static void close_output_files(void)
{
    for (int i = 0; i < n_idents; i++)
        if (idents[i].f) { fclose(idents[i].f); idents[i].f = NULL; }
}

static int cmp_ident(const void *a, const void *b) { return strcmp((*(ident_t *const *)a)->name, (*(ident_t *const *)b)->name); }

static void print_results(void)
{
    printf("Barcode\tCount\tLocation\n");
    ident_t **order = (ident_t **)malloc((size_t)n_idents * sizeof(ident_t *));
    for (int i = 0; i < n_idents; i++) order[i] = &idents[i];
    qsort(order, (size_t)n_idents, sizeof(ident_t *), cmp_ident);
    uint64_t total = 0;
    for (int i = 0; i < n_idents; i++) {
        printf("%s\t%llu\t%s\n", order[i]->name, (unsigned long long)order[i]->count, order[i]->path);
        total += order[i]->count;
    }
    printf("total\t%llu\n", (unsigned long long)total);
    free(order);
}

/* ------------------------------------------------------------------------------------------------ input */
static char *buf;
static size_t buf_cap, buf_len;
static int at_eof;

static void fill(void)
{
    while (!at_eof && buf_len < buf_cap) {
        ssize_t k = read(0, buf + buf_len, buf_cap - buf_len);
        if (k < 0) { if (errno == EINTR) continue; die(255, "Error: reading STDIN failed\n"); }
        if (k == 0) at_eof = 1;
        buf_len += (size_t)k;
    }
}

typedef struct { size_t start, seq_end; size_t end; } rec_t;     /* [start,end) = the record's lines; sequence = line 2 */

int main(int argc, char **argv)
{
    parse_command_line(argc, argv);
    load_barcode_file(o_bcfile);

    /* open_and_detect_input_format, :330-350 */
    const char *cap_env = getenv("FASTX_CHUNK_BYTES");
    buf_cap = cap_env && atoll(cap_env) > 0 ? (size_t)atoll(cap_env) : ((size_t)64 << 20);
    buf = (char *)malloc(buf_cap);
    if (!buf) die(255, "Error: out of memory\n");
    fill();
    if (buf_len == 0) die(255, "Error: unknown file format. First character = '' (expecting > or @)\n");
    int fastq;
    if (buf[0] == '>') { fastq = 0; if (o_debug) fprintf(stderr, "Detected FASTA format\n"); }
    else if (buf[0] == '@') { fastq = 1; if (o_debug) fprintf(stderr, "Detected FASTQ format\n"); }
    else die(255, "Error: unknown file format. First character = '%c' (expecting > or @)\n", buf[0]);
    const int lpr = fastq ? 4 : 2;

    create_output_files();
    if (n_entries == 0) {      /* an empty barcode list: $barcodes_length is undef in the script; every read is unmatched */
        barcode_len = 0; stride = 16;
    }
    fxg_ctx *ctx = fxh_gpu_open();
    fxg_barcode_table table = { entries, entry_len, n_entries, barcode_len, (int32_t)o_mismatches };

    rec_t *recs = NULL;
    size_t cap_recs = 0;
    uint8_t *frag = NULL;
    int32_t *flen = NULL, *best = NULL;
    size_t cap_frag = 0;

    for (;;) {
        /* split the buffer into complete records (at end of input the last line may lack its newline) */
        size_t n = 0, pos = 0, rec_start = 0, seq_end = 0;
        int line_in_rec = 0;
        const char *missing = NULL;
        while (pos < buf_len) {
            const char *nl = (const char *)memchr(buf + pos, '\n', buf_len - pos);
            size_t line_end;
            if (nl) line_end = (size_t)(nl - buf) + 1;
            else if (at_eof) line_end = buf_len;
            else break;
            if (line_in_rec == 0) rec_start = pos;
            if (line_in_rec == 1) seq_end = nl ? line_end - 1 : line_end;      /* chomp: the "\n" only */
            pos = line_end;
            if (++line_in_rec == lpr) {
                if (n == cap_recs) { cap_recs = cap_recs ? cap_recs * 2 : (1u << 16); recs = (rec_t *)realloc(recs, cap_recs * sizeof(rec_t)); }
                recs[n].start = rec_start; recs[n].seq_end = seq_end; recs[n].end = line_end;
                n++;
                line_in_rec = 0;
            }
        }
        const size_t consumed = line_in_rec == 0 ? pos : rec_start;
        if (at_eof && line_in_rec != 0)          /* read_record, :313-328 */
            missing = line_in_rec == 1 ? "Error: bad input file, expecting line with sequences\n"
                    : line_in_rec == 2 ? "Error: bad input file, expecting line with sequence name2\n"
                                       : "Error: bad input file, expecting line with quality scores\n";
        if (n == 0 && !at_eof && buf_len == buf_cap) {          /* one record larger than the buffer: grow */
            buf_cap *= 2;
            buf = (char *)realloc(buf, buf_cap);
            if (!buf) die(255, "Error: out of memory\n");
            fill();
            continue;
        }

        if (n > 0) {
            if (n > cap_frag) {
                if (frag) { fxg_free_pinned(frag); fxg_free_pinned(flen); fxg_free_pinned(best); }
                cap_frag = n + n / 4 + 1024;
                frag = (uint8_t *)fxg_alloc_pinned(cap_frag * (size_t)stride);
                flen = (int32_t *)fxg_alloc_pinned(cap_frag * sizeof(int32_t));
                best = (int32_t *)fxg_alloc_pinned(cap_frag * sizeof(int32_t));
                if (!frag || !flen || !best) die(255, "Error: out of pinned memory\n");
            }
            for (size_t i = 0; i < n; i++) {
                /* the fragment the barcodes are tested against, :244-249 */
                const char *name_end = (const char *)memchr(buf + recs[i].start, '\n', recs[i].end - recs[i].start);
                const size_t seq_start = (size_t)(name_end - buf) + 1;
                const size_t L = recs[i].seq_end - seq_start;
                const size_t fl = L < (size_t)barcode_len ? L : (size_t)barcode_len;
                uint8_t *row = frag + i * (size_t)stride;
                memset(row, 0, (size_t)stride);
                memcpy(row, buf + (o_bol ? seq_start : recs[i].seq_end - fl), fl);
                flen[i] = (int32_t)fl;
            }
            if (n_entries > 0) {
                fxg_batch fb = { frag, NULL, flen, 0, stride, (int64_t)n };
                fxh_gpu_check(ctx, fxg_barcode_host(ctx, &fb, &table, best, NULL), "fxg_barcode_host");
            } else {
                for (size_t i = 0; i < n; i++) best[i] = -1;
            }
            for (size_t i = 0; i < n; i++) {
                const int id = best[i] >= 0 ? entry_ident[best[i]] : unmatched_ident;
                const char *name_end = (const char *)memchr(buf + recs[i].start, '\n', recs[i].end - recs[i].start);
                const size_t seq_start = (size_t)(name_end - buf) + 1;
                if (o_debug) {
                    fprintf(stderr, "sequence %.*s: \n", (int)(recs[i].seq_end - seq_start), buf + seq_start);
                    fprintf(stderr, "sequence %.*s matched barcode: %s\n", (int)(recs[i].seq_end - seq_start), buf + seq_start, idents[id].name);
                }
                idents[id].count++;
                FILE *f = idents[id].f;
                /* write_record, :353-368: name line as read, chomped sequence + "\n", then lines 3 and 4 as read */
                fwrite(buf + recs[i].start, 1, recs[i].seq_end - recs[i].start, f);
                fputc('\n', f);
                const size_t after_seq = recs[i].seq_end + ((recs[i].seq_end < recs[i].end && buf[recs[i].seq_end] == '\n') ? 1 : 0);
                if (fastq) fwrite(buf + after_seq, 1, recs[i].end - after_seq, f);
            }
        }
        if (missing) { close_output_files(); die(255, "%s", missing); }
        if (at_eof) break;
        memmove(buf, buf + consumed, buf_len - consumed);
        buf_len -= consumed;
        fill();
    }
    close_output_files();
    fxg_destroy(ctx);
    if (!o_quiet) print_results();
    return 0;
}
