/* CPU-only harness for the host layer of the drop-in tools (fastx_toolkit_b200/csrc/host/fxh.c): block reader ->
 * SoA batches -> block writer, with NO GPU.  The few libfxg.so entry points fxh.c needs are stubbed below (pinned memory =
 * malloc), so the program is `fastx_trimmer` with default arguments minus the per-byte checks that run on the GPU:
 * for structurally valid input it must print what the reference's fastx_trimmer prints, and for structurally broken
 * input the same prefix, message and exit status (tests/test_host_layer.py).  Test infrastructure only. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fxg.h"
#include "fxh.h"

/* ---- stubs for libfxg.so ---- */
void *fxg_alloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 1); }
void fxg_free_pinned(void *p) { free(p); }
int fxg_host_register(void *p, size_t bytes) { (void)p; (void)bytes; return FXG_OK; }
int fxg_host_unregister(void *p) { (void)p; return FXG_OK; }
int fxg_init(int device, fxg_ctx **out) { (void)device; *out = NULL; return FXG_ERR_CUDA; }
const char *fxg_strerror(int code) { (void)code; return "stub"; }
const char *fxg_last_error(const fxg_ctx *ctx) { (void)ctx; return "stub"; }

static const char *const usage_text = "usage: fxh_passthrough [-h] [-v] [-z] [-i INFILE] [-o OUTFILE] [-Q N]\n";

int main(int argc, char **argv)
{
    fxh_parse_cmdline(argc, argv, "", NULL, usage_text);
    fxh_reader *rd = fxh_reader_open(fxh_input_filename(), FXH_FASTA_OR_FASTQ, fxh_q_offset(), 0);
    const int fastq = fxh_reader_is_fastq(rd);
    fxh_writer *wr = fxh_writer_open(fxh_output_filename(), fastq, fxh_compress_output());
    fxh_batch *b;
    while ((b = fxh_reader_next(rd, fxh_batch_reads())) != NULL)
        for (int64_t i = 0; i < b->n; i++)
            fxh_write_record(wr, b, i, b->seq + (size_t)i * b->stride, b->qual ? b->qual + (size_t)i * b->stride : NULL, b->len[i]);
    fxh_writer_close(wr);
    if (fxh_verbose()) {
        FILE *f = fxh_report_file();
        fprintf(f, "Input: %zu reads.\n", fxh_num_input_reads(rd));
        fprintf(f, "Output: %zu reads.\n", fxh_num_output_reads(wr));
    }
    return 0;
}
