/*
 * fxh.h — host side (C) of the drop-in FASTX tools: libfastx-compatible command line, block FASTA/FASTQ
 * reader that packs records into pinned SoA slabs for libfxg.so, and block writer.
 *
 * Mirrors, for the six hot-path tools, the behaviour of
 *   src/libfastx/fastx_args.c:76-143   shared options  -h -v -z -i FILE -o FILE -Q N, report stream rule
 *   src/libfastx/fastx.c:86-116        format sniffing (first byte '>' / '@')
 *   src/libfastx/fastx.c:314-404       record reader (4 / 2 lines, chomp at CR/LF, ASCII vs numeric quality)
 *   src/libfastx/fastx.c:440-473       record writer (ASCII in => ASCII out, numeric in => numeric out)
 *   src/libfastx/fastx.c:475-497       get_reads_count ("N-COUNT" collapsed ids, FASTA only)
 * Error texts are the reference's (errx(1, ...)); the per-byte checks themselves (legal bases, legal quality
 * values) run on the GPU — the host only re-reads the one offending record to word the message.
 */
#ifndef FXH_H
#define FXH_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/uio.h>

#include "fxg.h"

#define FXH_MAX_LINE 25000          /* src/libfastx/fastx.h:33-35 MAX_SEQ_LINE_LENGTH */

enum { FXH_FASTA_ONLY = 0, FXH_FASTA_OR_FASTQ = 1, FXH_FASTQ_ONLY = 2 };

/* ---- command line (fastx_args.c) ---- */
typedef int (*fxh_parse_arg_fn)(int optind_, int optc, char *optarg_);
int         fxh_parse_cmdline(int argc, char *argv[], const char *program_options, fxh_parse_arg_fn fn, const char *usage);
const char *fxh_input_filename(void);
const char *fxh_output_filename(void);
int         fxh_verbose(void);
int         fxh_compress_output(void);
int         fxh_q_offset(void);
FILE       *fxh_report_file(void);

/* ---- one batch of records, packed for the GPU ---- */
typedef struct {
    /* SoA slabs (pinned): read i at seq + i*stride / qual + i*stride */
    uint8_t *seq, *qual;
    int32_t *len;              /* per record */
    int32_t *width;            /* clipper: DP matrix width (running max length), else unused */
    int32_t *weight;           /* get_reads_count() per record */
    int32_t stride;
    int64_t n, cap;
    int     numeric_qual;      /* 1: this batch holds numeric-quality records, encoded with offset 33 */
    /* record text kept for re-emission */
    const char **name;   int32_t *name_len;      /* line 1 without the prefix char */
    const char **name2;  int32_t *name2_len;     /* line 3 without its first char (FASTQ) */
    uint64_t *line_no;         /* input line number of line 1 of each record */
    int64_t first_index;       /* index of record 0 within the whole input */
} fxh_batch;

typedef struct fxh_reader fxh_reader;

/* allowed: FXH_FASTA_ONLY / FXH_FASTA_OR_FASTQ / FXH_FASTQ_ONLY.  stale_rows != 0 packs clipper rows the way
 * the reference's aligner sees them (NUL + stale bytes beyond the read). */
fxh_reader *fxh_reader_open(const char *filename, int allowed, int q_offset, int stale_rows);
int         fxh_reader_is_fastq(const fxh_reader *r);
/* Next batch (<= max_reads records), NULL at end of input.  Structural errors are deferred: the batch holding
 * the records before the broken one is returned first, then fxh_reader_next() dies with the reference's text. */
fxh_batch  *fxh_reader_next(fxh_reader *r, int64_t max_reads);
size_t      fxh_num_input_sequences(const fxh_reader *r);
size_t      fxh_num_input_reads(const fxh_reader *r);
/* die with the reference's message for record `idx` of batch b, which the GPU flagged as invalid */
void        fxh_die_bad_record(const fxh_reader *r, const fxh_batch *b, int64_t idx);
fxg_batch   fxh_as_fxg_batch(const fxh_batch *b, int with_qual);

/* ---- raw access for the GPU text path (fxg_text_*): the unparsed bytes the reader currently holds ---- */
size_t      fxh_reader_raw(fxh_reader *r, char **p);                       /* refills first; 0 = nothing left          */
void        fxh_reader_consume(fxh_reader *r, size_t bytes, int64_t records);   /* 4-line FASTQ records             */
void        fxh_reader_pin(fxh_reader *r);                                 /* page-lock the text buffer for DMA        */
int         fxh_reader_at_eof(const fxh_reader *r);
/* clipper: tell the packer what the aligner's query buffer holds after records the text path consumed
 * (the last read, all earlier reads having had the same length) */
void        fxh_reader_seed_shadow(fxh_reader *r, const char *last_seq, int len);
int         fxh_text_path_enabled(void);                                   /* FASTX_TEXT_PATH=0 disables it            */
size_t      fxh_text_chunk_bytes(void);                                    /* FASTX_CHUNK_BYTES (default 64 MB)        */

/* ---- hooks for the streaming engine (fxh_stream.c), which reads the file descriptor itself while it runs ---- */
int         fxh_reader_fd(const fxh_reader *r);
void        fxh_reader_detach(fxh_reader *r, char **p, size_t *len, int *eof);
void        fxh_reader_account(fxh_reader *r, int64_t records, int64_t reads, int lines_per_record);
void        fxh_reader_restart(fxh_reader *r, const struct iovec *iov, int niov, int eof);

/* ---- writer ---- */
typedef struct fxh_writer fxh_writer;
fxh_writer *fxh_writer_open(const char *filename, int fastq, int compress);
/* emit record i of b with the sequence (and quality) cut to out_len; seq/qual rows may come from another slab */
void        fxh_write_record(fxh_writer *w, const fxh_batch *b, int64_t i, const uint8_t *seq_row, const uint8_t *qual_row,
                             int32_t out_len);
/* same, with another identifier line (fastq_to_fasta -r, fastq_to_fasta.c:84-85) */
void        fxh_write_record_named(fxh_writer *w, const fxh_batch *b, int64_t i, const uint8_t *seq_row, const uint8_t *qual_row,
                                   int32_t out_len, const char *name, int32_t name_len);
void        fxh_write_raw(fxh_writer *w, const char *text, size_t bytes, int64_t records);   /* already formatted */
void        fxh_writer_write_now(fxh_writer *w, const char *text, size_t bytes, int64_t records, int64_t reads);
int         fxh_writer_frames_gzip(const fxh_writer *w);      /* -z handled by this writer (GPU DEFLATE blocks / stored blocks) */
void        fxh_writer_write_deflated(fxh_writer *w, const char *blocks, size_t bytes, uint64_t raw_len, uint32_t crc_pure, int64_t records,
                                      int64_t reads);
void        fxh_writer_close(fxh_writer *w);
size_t      fxh_num_output_sequences(const fxh_writer *w);
size_t      fxh_num_output_reads(const fxh_writer *w);

/* ---- GPU context helpers ---- */
fxg_ctx    *fxh_gpu_open(void);                 /* FASTX_GPU=<index> (default 0); dies if no GPU: no CPU fallback */
fxg_ctx    *fxh_gpu_open_dev(int dev);
int         fxh_first_device(void);         /* CUDA ordinal of FASTX_GPU (the visible set is narrowed to the GPUs in use)  */
int         fxh_gpu_count(void);                /* FASTX_GPUS=<n> (default 1): tools that can use several GPUs */
void        fxh_gpu_check(fxg_ctx *ctx, int rc, const char *what);
int64_t     fxh_batch_reads(void);              /* FASTX_BATCH_READS (default 2 M) */
double      fxh_now(void);                      /* monotonic seconds (FASTX_TIMING=1 phase report) */

#endif
