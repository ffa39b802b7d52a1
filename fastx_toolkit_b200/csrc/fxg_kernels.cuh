// fxg_kernels.cuh — launch-parameter structs shared between the kernels (fxg_kernels.cu) and the
// C-ABI layer (fxg_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fxg_device.cuh"

namespace fxg {

constexpr int THREADS = 256;     // threads per CTA for the tile-pipeline kernels
constexpr int MAX_STAGES = 8;
constexpr int W_WARPS = 4;       // warps per CTA in the warp-private pipeline kernels
constexpr int W_THREADS = W_WARPS * 32;
constexpr int MAX_DYN_SMEM = 227 * 1024 - 1024;   // per-CTA opt-in limit minus the kernels' static smem (barriers)

enum { MODE_TRIM = 0, MODE_FILTER = 1, MODE_VALIDATE = 2, MODE_HASN = 3, MODE_ARTIFACT = 4 };

// counters block in device memory (one per context)
enum { CNT_OUT = 0, CNT_FIRST_BAD = 1, CNT_AUX0 = 2, CNT_WORDS = 16 };

struct TilePlan {
    int32_t g;            // lanes cooperating on one read (power of two, 1..32)
    int32_t tile_reads;   // reads per TMA tile
    int32_t stages;       // smem ring depth
    int32_t grid;         // persistent CTAs
    uint32_t smem_bytes;  // dynamic shared memory per CTA
    int32_t warp_ring;    // 1: warp-private pipeline kernel (tile = 32/g reads per warp), 0: CTA-tile kernel
    int32_t rot_shift;    // warp-ring, g == 1: lane (l&7)>>rot_shift starts that many chunks into its read
};

struct ScanParams {
    const uint8_t *seq;   // may be NULL (HAS_SEQ=false instantiation)
    const uint8_t *qual;
    const int32_t *len;   // may be NULL
    int32_t uniform_len;
    int32_t stride;
    int64_t n;
    const int64_t *n_dev; // when not NULL: the number of reads lives on the device (fused pipelines); n is then an upper bound
    int32_t tile_reads;
    int32_t stages;
    int32_t rot_shift;
    QualK qk;
    int32_t min_len;      // trim: -l
    int32_t pct_keep;     // filter: 100 - p   (keep iff 100*low <= L*pct_keep)
    int32_t force_drop;   // filter: -p 0 with -q > 93 drops everything (reference walks off its table)
    void *out;            // int32[n] (trim) or uint8[n] (filter)
    int64_t index_base;
    unsigned long long *counters;
};

struct RevcompParams {
    const uint8_t *seq;
    const uint8_t *qual;  // may be NULL (FASTA)
    const int32_t *len;
    int32_t uniform_len;
    int32_t stride;
    int64_t n;
    int32_t tile_reads;
    int32_t stages;
    QualK qk;
    uint8_t *out_seq;
    uint8_t *out_qual;
    int64_t index_base;
    unsigned long long *counters;
};

struct SynthParams {
    uint8_t *seq;
    uint8_t *qual;
    int64_t n, first_read, n_total;
    int32_t len, stride;
    uint64_t seed;
    int32_t kind, q_offset;
};

// ---- K-STATS (fxg_stats4.cu; global-atomics fallback in fxg_stats.cu) ---------------------------------------------------
constexpr int ST_QWIN = 64;                       // q' = q+15 in [0,64) lives in shared memory
constexpr int ST_MAXW = 40;                       // words (4 cycles each) per pass: 160 cycles
// k_stats4: lane = read; a bin owns 96 words: 64 words of u16 pairs for cycles 0..127 (word
// 32*(wi>>1) + 8k + c, half wi&1, for window word w = 4c + wi and byte k) + 32 full words for cycles 128..159 (64 + 4(w-32) + k)
constexpr int S4_PITCH = 96 * 4;                  // bytes per bin; 96 = 0 (mod 32): the bank of a counter never depends on the data
constexpr int S4_HIST_BYTES = 4 * ST_QWIN * S4_PITCH;   // 256 bins: 98 304 bytes
constexpr int S4_DUMMY_BYTES = 128;               // one scratch counter per lane (masked-off increments land here)
constexpr int S4_WARPS = 12;                      // 12 x 10 KB tiles (150 bp) beside the histogram
constexpr int S4_TILE_READS = 32;

struct StatsParams {
    const uint8_t *seq;
    const uint8_t *qual;      // NULL: FASTA (simple kernel only)
    const int32_t *len;
    const int32_t *weight;    // NULL: 1 per read (simple kernel only)
    int32_t uniform_len;
    int32_t stride;
    int64_t n;
    int32_t tile_reads;       // reads per warp tile (<= 32)
    int32_t stages;
    QualK qk;
    int32_t w0;               // first 4-cycle word covered by this pass
    int32_t nw;               // words covered (<= ST_MAXW)
    int32_t max_cycles;
    unsigned long long *hist; // u64 [max_cycles][5][109]
    int64_t index_base;
    unsigned long long *counters;
};

// ---- K-CLIP (fxg_clip.cu) ---------------------------------------------------------------------------
struct ClipParams {
    const uint8_t *seq;
    const uint8_t *qual;      // may be NULL
    const int32_t *len;
    const int32_t *width;     // DP matrix width per read (>= len); NULL: width = len
    int32_t uniform_len;
    int32_t stride;
    int64_t n;
    const int64_t *n_dev;     // when not NULL: the number of reads lives on the device (fused pipelines); n is an upper bound
    QualK qk;
    uint8_t adapter[104];
    int32_t alen;
    int32_t min_length, keep_delta, discard_non_clipped, discard_clipped, discard_unknown, min_adapter_len;
    int32_t *out_len;         // emitted length, -1 when discarded
    uint8_t *out_class;       // may be NULL
    int32_t *out_cut;         // may be NULL
    int64_t index_base;
    unsigned long long *counters;
    // integer fast path (k_clip_dpx): reads that need the fp32 kernel are appended here; the second launch
    // (list_pass = 1) runs k_clip_bits over exactly those
    int32_t *list;
    unsigned long long *list_count;
    int32_t list_pass;
};

// ---- K-BARCODE (fxg_barcode.cu) -----------------------------------------------------------------------
constexpr int BC_MAX_STRIDE = 64;        // barcodes of up to 64 characters
constexpr int BC_TILE_ENTRIES = 512;     // entries per launch (32 KB of shared memory at stride 64)

struct BarcodeParams {
    const uint8_t *frag;      // [n][stride]
    const int32_t *flen;      // [n]
    int64_t n;
    int32_t stride;           // 16, 32, 48 or 64
    const uint8_t *entries;   // device, [n_entries][stride], zero padded
    const int32_t *elen;      // device, [n_entries]
    int32_t e0, e1;           // entries handled by this launch
    int32_t barcode_len;      // B
    int32_t allowed;          // --mismatches
    int32_t last;             // 1: final launch, turn (best_mm, best) into the result
    int32_t *best_mm;         // [n] running minimum (first launch initialises it to B)
    int32_t *best;            // [n] running entry index, -1 = none; after the last launch: the answer
};

cudaError_t launch_barcode(const BarcodeParams &p, int sm_count, cudaStream_t st);

// ---- fused pipelines (fxg_pipeline.cu) ----
size_t pipe_scan_tmp_bytes(int64_t n);
// n = the bound the kernels are launched over; n_dev (device, may be NULL) = the live rows among them
cudaError_t launch_pipe_flags_scan(const int32_t *new_len, const uint8_t *keep, int64_t n, const int64_t *n_dev, int32_t *flags, int32_t *pos,
                                   void *tmp, size_t tmp_bytes, int64_t *count_out, int sm_count, cudaStream_t st);
cudaError_t launch_pipe_gather(const uint8_t *src_seq, const uint8_t *src_qual, int stride, int64_t n, const int64_t *n_dev, const int32_t *flags,
                               const int32_t *pos, const int32_t *new_len, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx,
                               uint8_t *dst_seq, uint8_t *dst_qual, int32_t *dst_len, int32_t *dst_idx, int sm_count, cudaStream_t st);
cudaError_t launch_stale_rows(const uint8_t *seq, const int32_t *len, int stride, int64_t n, const int64_t *n_dev, uint8_t *out_seq,
                              int32_t *out_width, void *scratch, size_t scratch_bytes, size_t *need, int sm_count, cudaStream_t st);
cudaError_t launch_pipe_keep_all(int64_t n, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx, int32_t *final_len, int sm_count,
                                 cudaStream_t st);
cudaError_t launch_pipe_scatter(int64_t n, const int64_t *n_dev, const int32_t *flags, const int32_t *new_len, const int32_t *cur_len, int uniform_len,
                                const int32_t *cur_idx, int32_t *final_len, int sm_count, cudaStream_t st);
cudaError_t launch_stats_simple(const StatsParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_clip(const ClipParams &p, int sm_count, int max_width, cudaStream_t st);
cudaError_t launch_extra(int op, const uint8_t *seq, const uint8_t *qual, const int32_t *len, int uniform_len, int stride, int64_t n,
                         int q_offset, int thr_q, int mask_char, uint8_t *out_seq, uint8_t *flags, int64_t index_base,
                         unsigned long long *counters, int sm_count, cudaStream_t st);
cudaError_t launch_stats4(const StatsParams &p, int grid, uint32_t smem_bytes, cudaStream_t st);
cudaError_t launch_hash(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint64_t *out, cudaStream_t st);

cudaError_t launch_scan(int mode, bool has_seq, const TilePlan &plan, const ScanParams &p, cudaStream_t st);
cudaError_t launch_revcomp(bool has_qual, const TilePlan &plan, const RevcompParams &p, cudaStream_t st);
cudaError_t launch_synth(const SynthParams &p, cudaStream_t st);

}  // namespace fxg
