/*
 * fxh_stream.h — the streaming engine of the drop-in tools: read(2) / H2D / GPU / D2H / write(2) overlapped.
 *
 * The reference tools run `while (fastx_read_next_record) { BODY; fastx_write_record }` on one thread
 * (e.g. src/fastq_quality_trimmer/fastq_quality_trimmer.c:91-103).  Here the input is cut into text chunks at record
 * boundaries by a reader thread, W workers per GPU (FASTX_GPUS GPUs, chunks in arrival order to whichever worker is free)
 * push them through the GPU text path (fxg_text_*: parse, pack, BODY, emit), and the calling thread writes the emitted
 * text back IN INPUT ORDER.  Anything the GPU path hands back (a chunk it flags as an anomaly, a trailing partial record)
 * stops the engine and repositions the fxh_reader on the first unprocessed byte, so that the record-by-record host path
 * that follows produces the reference's output prefix, message and exit status.
 */
#ifndef FXH_STREAM_H
#define FXH_STREAM_H

#include "fxh.h"

enum { FXS_TRIM = 0, FXS_FILTER = 1, FXS_REVCOMP = 2, FXS_STATS = 3, FXS_CLIP = 4, FXS_COLLAPSE = 5 };

typedef struct {
    int op;                         /* FXS_*                                                                           */
    int a0, a1;                     /* TRIM: -t, -l   FILTER: -q, -p   CLIP: a0 = -k (adapter-only reads)              */
    const fxg_clip_opts *clip;      /* CLIP                                                                            */
    int ngpu, first_dev;            /* GPUs first_dev .. first_dev + ngpu - 1                                          */
    uint64_t **hist_dev;            /* STATS: one device histogram per GPU (zeroed by the caller), max_cycles cycles   */
    int32_t max_cycles;
    fxg_collapser **collapsers;     /* COLLAPSE: one count map per GPU (several GPUs: merged afterwards, fxg_dcollapse_*)     */
    /* results */
    int64_t records, reads;         /* consumed by the engine                                                          */
    int64_t chunks, numeric_chunks; /* text chunks pushed through the GPU path / of them with numeric qualities        */
    int     max_len;                /* longest read seen                                                               */
    unsigned int clip_class[6];     /* CLIP: reads per FXG_CLIP_* class                                                */
    double  t_total, t_wait_gpu, t_write;
} fxs_job;

/* Runs the GPU text path over the input of `rd` (FASTQ or FASTA, whatever the reader detected), writing through `wr`
 * (may be NULL for STATS / COLLAPSE).  Returns 0 when the engine consumed the whole input, 1 when it stopped early: the
 * reader then holds the unprocessed rest and the caller's record-by-record loop takes over. */
int fxs_run(fxs_job *job, fxh_reader *rd, fxh_writer *wr);

#endif
