// fxg_api.cu — the C ABI declared in include/fxg.h: context, memory, launch planning and the
// host-buffer pipelines (pinned host slab -> H2D on a side stream -> kernel -> D2H).
// No CPU fallback anywhere: every failure is reported as an error code.
#include <vector>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fxg.h"
#include "fxg_kernels.cuh"

using namespace fxg;

namespace {
constexpr int PIPE_LANES = 3;          // chunks in flight in a *_host call
constexpr unsigned long long NO_BAD = ~0ull;
}

struct fxg_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t hbm_bytes;
    cudaStream_t own_stream;
    cudaStream_t stream;               // own_stream or an adopted one
    unsigned long long *d_counters;    // CNT_WORDS
    unsigned long long *h_counters;    // pinned mirror
    fxg_report report;
    int64_t launches;
    int tune_tile_reads, tune_stages, tune_ctas;
    char err[256];
    // host-pipeline resources (grow-only)
    cudaStream_t lane_stream[PIPE_LANES];
    void *lane_buf[PIPE_LANES][6];     // seq, qual, len, aux, out0, out1
    size_t lane_cap[PIPE_LANES][6];
};

#define CK(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,    \
                     cudaGetErrorString(e_));                                                      \
            return FXG_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static int arg_error(fxg_ctx *ctx, const char *what)
{
    snprintf(ctx->err, sizeof(ctx->err), "bad argument: %s", what);
    return FXG_ERR_ARG;
}

extern "C" const char *fxg_strerror(int code)
{
    switch (code) {
    case FXG_OK: return "ok";
    case FXG_ERR_CUDA: return "CUDA error (no usable GPU, or a launch/copy failed)";
    case FXG_ERR_ARG: return "invalid argument";
    case FXG_ERR_NOMEM: return "out of memory";
    case FXG_ERR_UNSUPPORTED: return "unsupported configuration";
    case FXG_ERR_NCCL: return "NCCL error";
    default: return "unknown error";
    }
}

static char g_init_err[256] = "no context";   // detail of the last fxg_init failure
extern "C" const char *fxg_last_error(const fxg_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

#define INIT_FAIL(what, e)                                                                         \
    do {                                                                                           \
        snprintf(g_init_err, sizeof(g_init_err), "fxg_init: %s: %s", what, cudaGetErrorString(e)); \
        cudaGetLastError();                                                                        \
        return FXG_ERR_CUDA;                                                                       \
    } while (0)

extern "C" int fxg_init(int device, fxg_ctx **out)
{
    if (!out) return FXG_ERR_ARG;
    *out = NULL;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) INIT_FAIL("cudaGetDeviceCount", e);
    if (count <= 0 || device < 0 || device >= count) {
        snprintf(g_init_err, sizeof(g_init_err), "fxg_init: device %d not present (%d CUDA devices)", device, count);
        return FXG_ERR_CUDA;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) INIT_FAIL("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) INIT_FAIL("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        snprintf(g_init_err, sizeof(g_init_err), "fxg_init: device %d is sm_%d%d; this build is sm_100a only", device, prop.major, prop.minor);
        return FXG_ERR_CUDA;
    }
    fxg_ctx *ctx = (fxg_ctx *)calloc(1, sizeof(fxg_ctx));
    if (!ctx) return FXG_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->hbm_bytes = prop.totalGlobalMem;
    // anything that fails from here on releases what was created before it (fxg_destroy copes with the NULL members)
#define INIT_FAIL_CTX(what, e)                                                                     \
    do {                                                                                           \
        snprintf(g_init_err, sizeof(g_init_err), "fxg_init: %s: %s", what, cudaGetErrorString(e)); \
        cudaGetLastError();                                                                        \
        fxg_destroy(ctx);                                                                          \
        return FXG_ERR_CUDA;                                                                       \
    } while (0)
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) INIT_FAIL_CTX("cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    for (int l = 0; l < PIPE_LANES; l++)
        if ((e = cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking)) != cudaSuccess) INIT_FAIL_CTX("cudaStreamCreate", e);
    if ((e = cudaMalloc(&ctx->d_counters, CNT_WORDS * sizeof(unsigned long long))) != cudaSuccess) INIT_FAIL_CTX("cudaMalloc", e);
    {   // stream-ordered scratch (clipper work list): keep freed blocks in the pool instead of returning them at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if ((e = cudaMallocHost(&ctx->h_counters, CNT_WORDS * sizeof(unsigned long long))) != cudaSuccess) INIT_FAIL_CTX("cudaMallocHost", e);
#undef INIT_FAIL_CTX
    *out = ctx;
    return fxg_report_reset(ctx);
}

extern "C" void fxg_destroy(fxg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int l = 0; l < PIPE_LANES; l++) {
        for (int b = 0; b < 6; b++) if (ctx->lane_buf[l][b]) cudaFree(ctx->lane_buf[l][b]);
        if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
    }
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

extern "C" int fxg_device_info(fxg_ctx *ctx, int *sm_count, size_t *hbm_bytes, int *cc_major, int *cc_minor)
{
    if (!ctx) return FXG_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return FXG_OK;
}

extern "C" int fxg_set_stream(fxg_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return FXG_ERR_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return FXG_OK;
}

extern "C" int fxg_use_own_stream(fxg_ctx *ctx)
{
    if (!ctx) return FXG_ERR_ARG;
    ctx->stream = ctx->own_stream;
    return FXG_OK;
}

extern "C" int fxg_set_tuning(fxg_ctx *ctx, int tile_reads, int stages, int ctas_per_sm)
{
    if (!ctx || tile_reads < 0 || stages < 0 || stages > MAX_STAGES || ctas_per_sm < 0) return FXG_ERR_ARG;
    ctx->tune_tile_reads = tile_reads;
    ctx->tune_stages = stages;
    ctx->tune_ctas = ctas_per_sm;
    return FXG_OK;
}

static int refresh_report(fxg_ctx *ctx, cudaStream_t st)
{
    CK(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, CNT_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    ctx->report.n_out = (int64_t)ctx->h_counters[CNT_OUT];
    ctx->report.first_bad_read = ctx->h_counters[CNT_FIRST_BAD] == NO_BAD ? -1 : (int64_t)ctx->h_counters[CNT_FIRST_BAD];
    for (int k = 0; k < 6; k++) ctx->report.aux[k] = (int64_t)ctx->h_counters[CNT_AUX0 + k];
    return FXG_OK;
}

extern "C" int fxg_sync(fxg_ctx *ctx)
{
    if (!ctx) return FXG_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    return refresh_report(ctx, ctx->stream);
}

extern "C" int fxg_get_report(fxg_ctx *ctx, fxg_report *out)
{
    if (!ctx || !out) return FXG_ERR_ARG;
    *out = ctx->report;
    return FXG_OK;
}

extern "C" int fxg_report_reset(fxg_ctx *ctx)
{
    if (!ctx) return FXG_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->report, 0, sizeof(ctx->report));
    ctx->report.first_bad_read = -1;
    for (int k = 0; k < CNT_WORDS; k++) ctx->h_counters[k] = 0;
    ctx->h_counters[CNT_FIRST_BAD] = NO_BAD;
    CK(ctx, cudaMemcpyAsync(ctx->d_counters, ctx->h_counters, CNT_WORDS * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return FXG_OK;
}

extern "C" int64_t fxg_kernel_launches(const fxg_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- memory ------------------------------------------------------------------------------------
extern "C" void *fxg_alloc_pinned(size_t bytes)
{
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
extern "C" void fxg_free_pinned(void *p) { if (p) cudaFreeHost(p); }
extern "C" int fxg_host_register(void *p, size_t bytes)
{
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? FXG_OK : (cudaGetLastError(), FXG_ERR_CUDA);
}
extern "C" int fxg_host_unregister(void *p)
{
    return cudaHostUnregister(p) == cudaSuccess ? FXG_OK : (cudaGetLastError(), FXG_ERR_CUDA);
}
extern "C" void *fxg_alloc_device(fxg_ctx *ctx, size_t bytes)
{
    if (!ctx) return NULL;
    void *p = NULL;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    return p;
}
extern "C" void fxg_free_device(fxg_ctx *ctx, void *p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }
extern "C" int fxg_memcpy_h2d(fxg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return FXG_ERR_ARG;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return FXG_OK;
}
extern "C" int fxg_memcpy_d2h(fxg_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return FXG_ERR_ARG;
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return FXG_OK;
}
extern "C" int fxg_memset_dev(fxg_ctx *ctx, void *dst, int value, size_t bytes)
{
    if (!ctx) return FXG_ERR_ARG;
    CK(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return FXG_OK;
}

// ---- launch planning -----------------------------------------------------------------------------
// G lanes per read: the 8 lanes of a quarter-warp must touch 8 distinct 16-byte bank groups when they
// issue LDS.128 at rows `stride` apart, which holds when G = lowbit(stride/16) (capped at 8); long
// reads get more lanes so a tile still fills the CTA.
static int choose_g(int stride)
{
    const int c = stride >> 4;
    int g = c & -c;
    if (g > 8) g = 8;
    while (g < 32 && c / g > 24) g <<= 1;
    return g;
}

static int gcd8(int c) { return (c & 1) ? 1 : (c & 2) ? 2 : (c & 4) ? 4 : 8; }

// FXG_TUNE="ring,g,stages,ctas" (ring: 1 warp-private kernel, 0 CTA-tile kernel; 0 = default for any
// field) — an experimentation knob for bench/profiling runs; production uses the defaults below.
static void env_tune(int *ring, int *g, int *stages, int *ctas)
{
    *ring = -1; *g = 0; *stages = 0; *ctas = 0;
    const char *e = getenv("FXG_TUNE");
    if (!e) return;
    int r = -1, a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d,%d", &r, &a, &b, &c) >= 1) { *ring = r; *g = a; *stages = b; *ctas = c; }
}

static int make_plan(fxg_ctx *ctx, const fxg_batch *b, int nslabs, int extra_stage_bufs, TilePlan *plan, bool revcomp = false)
{
    const int S = b->stride;
    const size_t per_read = (size_t)S * (size_t)nslabs;
    const size_t smem_max = MAX_DYN_SMEM;
    int t_ring, t_g, t_stages, t_ctas;
    env_tune(&t_ring, &t_g, &t_stages, &t_ctas);
    if (ctx->tune_stages) t_stages = ctx->tune_stages;
    if (ctx->tune_ctas) t_ctas = ctx->tune_ctas;
    memset(plan, 0, sizeof(*plan));

    // ---- warp-private pipeline: a warp's tile (32/g reads) must stay small enough that many warps fit
    if (t_ring != 0 && ctx->tune_tile_reads == 0) {
        int g = 0;
        for (int cand = 1; cand <= 8; cand <<= 1) {
            if (t_g ? cand == t_g : (size_t)(32 / cand) * per_read * (revcomp ? 2 : 1) <= 12 * 1024) { g = cand; break; }
        }
        if (g) {
            const size_t wstage = (size_t)(32 / g) * per_read;
            // one stage per warp: the other warps of the SM (16 with 4 CTAs) cover a warp's wait for HBM;
            // measured on B200 (150 bp): 1 stage x 4 CTAs 6276 GB/s, 2 stages x 2 CTAs 6082 GB/s
            int stages = t_stages ? t_stages : 1;
            if (revcomp) stages = 1;                          // k_revcomp_w: one input + one output tile per warp
            const int bufs = revcomp ? 2 : stages;
            while (!revcomp && stages > 1 && (size_t)stages * wstage * W_WARPS > smem_max) stages--;
            const size_t smem = (size_t)(revcomp ? bufs : stages) * wstage * W_WARPS;
            if (smem <= smem_max) {
                int ctas = t_ctas ? t_ctas : (int)((smem_max + 1024) / (smem + 1024));
                if (ctas < 1) ctas = 1;
                if (!t_ctas && ctas > (revcomp ? 5 : 4)) ctas = revcomp ? 5 : 4;   // measured optima (150 bp)
                if (ctas > 12) ctas = 12;
                const int64_t ntiles = (b->n + (32 / g) - 1) / (32 / g);
                int64_t grid = (int64_t)ctx->sm_count * ctas;
                const int64_t need = (ntiles + W_WARPS - 1) / W_WARPS;
                if (grid > need) grid = need;
                if (grid < 1) grid = 1;
                plan->g = g; plan->tile_reads = 32 / g; plan->stages = stages; plan->grid = (int)grid;
                plan->smem_bytes = (uint32_t)smem; plan->warp_ring = 1;
                const int d = gcd8(S >> 4);
                plan->rot_shift = d == 1 ? 3 : d == 2 ? 2 : d == 4 ? 1 : 0;
                return FXG_OK;
            }
        }
    }

    // ---- CTA-tile kernel (long reads, or forced)
    const int g = t_g ? t_g : choose_g(S);
    const int rpp = THREADS / g;
    int stages = t_stages ? t_stages : 3;
    size_t target = 32 * 1024;                        // bytes per stage
    int tr = ctx->tune_tile_reads ? ctx->tune_tile_reads : (int)(target / per_read);
    if (tr >= rpp) tr -= tr % rpp; else if (tr < 1) tr = 1;
    if ((int64_t)tr > b->n) tr = (int)(b->n > 0 ? b->n : 1);
    while (stages > 1 && (size_t)(stages + extra_stage_bufs) * per_read * tr > smem_max) stages--;
    while (tr > 1 && (size_t)(stages + extra_stage_bufs) * per_read * tr > smem_max) tr--;
    const size_t smem = (size_t)(stages + extra_stage_bufs) * per_read * tr;
    if (smem > smem_max || per_read * tr >= (1u << 20)) return arg_error(ctx, "read stride too large for shared memory");
    int ctas = t_ctas ? t_ctas : (int)(smem_max / (smem + 1024));
    if (ctas < 1) ctas = 1;
    if (ctas > 4) ctas = 4;
    const int64_t ntiles = (b->n + tr - 1) / tr;
    int64_t grid = (int64_t)ctx->sm_count * ctas;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    plan->g = g;
    plan->tile_reads = tr;
    plan->stages = stages;
    plan->grid = (int)grid;
    plan->smem_bytes = (uint32_t)smem;
    return FXG_OK;
}

static int check_batch(fxg_ctx *ctx, const fxg_batch *b, bool need_seq, bool need_qual, int q_offset)
{
    if (!ctx) return FXG_ERR_ARG;
    if (!b) return arg_error(ctx, "batch is NULL");
    if (b->n < 0) return arg_error(ctx, "n < 0");
    if (b->stride <= 0 || (b->stride & 15)) return arg_error(ctx, "stride must be a positive multiple of 16");
    if (need_seq && !b->seq) return arg_error(ctx, "seq is NULL");
    if (need_qual && !b->qual) return arg_error(ctx, "qual is NULL");
    if (((uintptr_t)b->seq & 15) || ((uintptr_t)b->qual & 15)) return arg_error(ctx, "slab base must be 16-byte aligned");
    if (!b->len && (b->uniform_len <= 0 || b->uniform_len > b->stride)) return arg_error(ctx, "uniform_len out of range");
    if (q_offset < 15 || q_offset > 127) {
        snprintf(ctx->err, sizeof(ctx->err), "quality offset %d unsupported (15..127)", q_offset);
        return FXG_ERR_UNSUPPORTED;
    }
    return FXG_OK;
}

// ---- synthetic data --------------------------------------------------------------------------------
extern "C" int fxg_synth_dev(fxg_ctx *ctx, uint8_t *seq, uint8_t *qual, int64_t n, int64_t first_read, int64_t n_total,
                             int32_t len, int32_t stride, uint64_t seed, int kind, int q_offset)
{
    if (!ctx) return FXG_ERR_ARG;
    if (n < 0 || len <= 0 || stride < len || (stride & 15)) return arg_error(ctx, "synth geometry");
    if (n == 0) return FXG_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    SynthParams p;
    p.seq = seq; p.qual = qual; p.n = n; p.first_read = first_read; p.n_total = n_total;
    p.len = len; p.stride = stride; p.seed = seed; p.kind = kind; p.q_offset = q_offset;
    CK(ctx, launch_synth(p, ctx->stream));
    ctx->launches++;
    return FXG_OK;
}

// ---- trim / filter -----------------------------------------------------------------------------------
// n_dev (device memory, fused pipelines only): the live reads among the b->n the launch is sized for
static int scan_enqueue(fxg_ctx *ctx, int mode, const fxg_batch *b, int q_offset, int thr_q, int min_len, int min_percent,
                        void *out, int64_t index_base, cudaStream_t st, const int64_t *n_dev = NULL)
{
    if (b->n == 0) return FXG_OK;
    const bool has_seq = b->seq != NULL;
    TilePlan plan;
    int rc = make_plan(ctx, b, has_seq ? 2 : 1, 0, &plan);
    if (rc) return rc;
    ScanParams p;
    p.seq = b->seq; p.qual = b->qual; p.len = b->len; p.uniform_len = b->uniform_len; p.stride = b->stride; p.n = b->n;
    p.n_dev = n_dev;
    p.tile_reads = plan.tile_reads; p.stages = plan.stages; p.rot_shift = plan.rot_shift;
    p.qk = make_qualk(q_offset, thr_q);
    p.min_len = min_len;
    p.pct_keep = 100 - min_percent;
    p.force_drop = (mode == MODE_FILTER && min_percent == 0 && thr_q > 93) ? 1 : 0;
    p.out = out; p.index_base = index_base; p.counters = ctx->d_counters;
    CK(ctx, launch_scan(mode, has_seq, plan, p, st));
    ctx->launches++;
    ctx->report.n_in += b->n;
    return FXG_OK;
}

extern "C" int fxg_trim_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int threshold, int min_len,
                            int32_t *out_len, int64_t index_base)
{
    int rc = check_batch(ctx, b, false, true, q_offset);
    if (rc) return rc;
    if (!out_len) return arg_error(ctx, "out_len is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    return scan_enqueue(ctx, MODE_TRIM, b, q_offset, threshold, min_len, 0, out_len, index_base, ctx->stream);
}

extern "C" int fxg_filter_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int min_percent,
                              uint8_t *keep, int64_t index_base)
{
    int rc = check_batch(ctx, b, false, true, q_offset);
    if (rc) return rc;
    if (!keep) return arg_error(ctx, "keep is NULL");
    if (min_percent < 0 || min_percent > 100) return arg_error(ctx, "min_percent must be 0..100");
    CK(ctx, cudaSetDevice(ctx->device));
    return scan_enqueue(ctx, MODE_FILTER, b, q_offset, min_quality, 0, min_percent, keep, index_base, ctx->stream);
}

// ---- revcomp ---------------------------------------------------------------------------------------
static int revcomp_enqueue(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *out_seq, uint8_t *out_qual,
                           int64_t index_base, cudaStream_t st)
{
    if (b->n == 0) return FXG_OK;
    const bool has_qual = b->qual != NULL;
    TilePlan plan;
    int rc = make_plan(ctx, b, has_qual ? 2 : 1, 2, &plan, true);
    if (rc) return rc;
    RevcompParams p;
    p.seq = b->seq; p.qual = b->qual; p.len = b->len; p.uniform_len = b->uniform_len; p.stride = b->stride; p.n = b->n;
    p.tile_reads = plan.tile_reads; p.stages = plan.stages;
    p.qk = make_qualk(q_offset, 0);
    p.out_seq = out_seq; p.out_qual = out_qual; p.index_base = index_base; p.counters = ctx->d_counters;
    {
        cudaError_t e = launch_revcomp(has_qual, plan, p, st);
        if (e != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "launch_revcomp: %s (n %lld stride %d ring %d g %d tile %d grid %d smem %u)", cudaGetErrorString(e),
                     (long long)b->n, b->stride, plan.warp_ring, plan.g, plan.tile_reads, plan.grid, plan.smem_bytes);
            return FXG_ERR_CUDA;
        }
    }
    ctx->launches++;
    ctx->report.n_in += b->n;
    return FXG_OK;
}

extern "C" int fxg_revcomp_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *out_seq, uint8_t *out_qual,
                               int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!out_seq || (b->qual && !out_qual)) return arg_error(ctx, "output slab is NULL");
    if (((uintptr_t)out_seq & 15) || ((uintptr_t)out_qual & 15)) return arg_error(ctx, "output slab must be 16-byte aligned");
    CK(ctx, cudaSetDevice(ctx->device));
    return revcomp_enqueue(ctx, b, q_offset, out_seq, out_qual, index_base, ctx->stream);
}

// ---- quality stats ---------------------------------------------------------------------------------
static int stats_enqueue(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist, int32_t max_cycles,
                         const int32_t *weight, int64_t index_base, cudaStream_t st)
{
    if (b->n == 0) return FXG_OK;
    StatsParams p;
    memset(&p, 0, sizeof(p));
    p.seq = b->seq; p.qual = b->qual; p.len = b->len; p.weight = weight; p.uniform_len = b->uniform_len;
    p.stride = b->stride; p.n = b->n; p.qk = make_qualk(q_offset, 0); p.max_cycles = max_cycles;
    p.hist = (unsigned long long *)hist; p.index_base = index_base; p.counters = ctx->d_counters;
    const int lmax = b->len ? b->stride : b->uniform_len;
    const int words = (lmax + 3) / 4;
    int t_ring, t_g, t_stages, t_ctas;
    env_tune(&t_ring, &t_g, &t_stages, &t_ctas);
    // k_stats4 (shared histogram, lane = read, warp pairs per tile buffer) needs qualities, unit weights, a tile of at least 8
    // reads beside the histogram and Q - 15 <= 64 (so that "0 <= q' < 64" implies the reader's range check); everything else
    // (FASTA, weights, -Q > 79, very long rows) goes to the global-atomics kernel
    long rfit4 = ((long)MAX_DYN_SMEM - (long)S4_HIST_BYTES - S4_DUMMY_BYTES) / S4_WARPS / (2L * b->stride);
    if (rfit4 > S4_TILE_READS) rfit4 = S4_TILE_READS;
    const bool fast4 = b->qual && !weight && rfit4 >= 8 && t_ring != 0 && q_offset - 15 <= 64 && b->n / rfit4 < (1ll << 31);
    if (fast4) {
        p.tile_reads = (int)rfit4; p.stages = 2;
        const uint32_t smem = (uint32_t)((size_t)S4_HIST_BYTES + S4_DUMMY_BYTES + (size_t)S4_WARPS * 2 * b->stride * rfit4);
        const int64_t ntiles = (b->n + rfit4 - 1) / rfit4;
        int64_t grid = ctx->sm_count;
        const int64_t need = (ntiles + S4_WARPS - 1) / S4_WARPS;
        if (grid > need) grid = need;
        for (int w0 = 0; w0 < words; w0 += ST_MAXW) {     // reads longer than 160 bases: one pass per 160 cycles
            p.w0 = w0; p.nw = (words - w0 < ST_MAXW) ? (words - w0) : ST_MAXW;
            CK(ctx, launch_stats4(p, (int)grid, smem, st));
            ctx->launches++;
        }
    } else {
        CK(ctx, launch_stats_simple(p, ctx->sm_count, st));
        ctx->launches++;
    }
    ctx->report.n_in += b->n;
    return FXG_OK;
}

extern "C" int fxg_stats_accum_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist, int32_t max_cycles,
                                   const int32_t *weight, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!hist || max_cycles <= 0) return arg_error(ctx, "hist is NULL or max_cycles <= 0");
    CK(ctx, cudaSetDevice(ctx->device));
    return stats_enqueue(ctx, b, q_offset, hist, max_cycles, weight, index_base, ctx->stream);
}

// ---- clipper ---------------------------------------------------------------------------------------
static int clip_enqueue(fxg_ctx *ctx, const fxg_batch *b, const int32_t *width, int q_offset, const fxg_clip_opts *o,
                        int32_t *out_len, uint8_t *out_class, int32_t *out_cut, int64_t index_base, cudaStream_t st,
                        const int64_t *n_dev = NULL)
{
    if (b->n == 0) return FXG_OK;
    ClipParams p;
    memset(&p, 0, sizeof(p));
    p.n_dev = n_dev;
    p.seq = b->seq; p.qual = b->qual; p.len = b->len; p.width = width; p.uniform_len = b->uniform_len;
    p.stride = b->stride; p.n = b->n; p.qk = make_qualk(q_offset, 0);
    p.alen = (int)strlen(o->adapter);
    memcpy(p.adapter, o->adapter, (size_t)p.alen);
    p.min_length = o->min_length; p.keep_delta = o->keep_delta; p.discard_non_clipped = o->discard_non_clipped;
    p.discard_clipped = o->discard_clipped; p.discard_unknown = o->discard_unknown; p.min_adapter_len = o->min_adapter_len;
    p.out_len = out_len; p.out_class = out_class; p.out_cut = out_cut; p.index_base = index_base; p.counters = ctx->d_counters;
    // Integer fast path (k_clip_dpx, two reads per thread in packed s16x2): batches of one length, adapters of
    // A/C/G/T only, up to 16 characters.  Reads with an 'N' go on a list and through the fp32 kernel in a second
    // launch.  FXG_CLIP_V=1 selects the fp32 kernel for everything.
    const char *cv = getenv("FXG_CLIP_V");
    bool dpx = !(cv && cv[0] == '1') && !width && !b->len && !n_dev && p.alen >= 1 && p.alen <= 16 && b->uniform_len <= 256 &&
               b->n < (1ll << 31);
    for (int i = 0; i < p.alen && dpx; i++) {
        const char c = o->adapter[i];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) dpx = false;
    }
    if (dpx) {
        void *scratch = NULL;
        CK(ctx, cudaMallocAsync(&scratch, 16 + (size_t)b->n * sizeof(int32_t), st));
        CK(ctx, cudaMemsetAsync(scratch, 0, 16, st));
        p.list_count = (unsigned long long *)scratch;
        p.list = (int32_t *)((char *)scratch + 16);
        p.list_pass = 0;
        CK(ctx, launch_clip(p, ctx->sm_count, b->uniform_len, st));
        p.list_pass = 1;
        CK(ctx, launch_clip(p, ctx->sm_count, b->uniform_len, st));
        CK(ctx, cudaFreeAsync(scratch, st));
        ctx->launches += 2;
    } else {
        CK(ctx, launch_clip(p, ctx->sm_count, (width || b->len) ? b->stride : b->uniform_len, st));
        ctx->launches++;
    }
    ctx->report.n_in += b->n;
    return FXG_OK;
}

static int check_clip_opts(fxg_ctx *ctx, const fxg_clip_opts *o)
{
    if (!o || !o->adapter) return arg_error(ctx, "clip options / adapter is NULL");
    const size_t al = strlen(o->adapter);
    if (al == 0 || al >= FXG_MAX_ADAPTER) return arg_error(ctx, "adapter length must be 1..99");
    if (o->keep_delta < 0) return arg_error(ctx, "keep_delta < 0");
    return FXG_OK;
}

extern "C" int fxg_clip_dev(fxg_ctx *ctx, const fxg_batch *b, const int32_t *width, int q_offset, const fxg_clip_opts *o,
                            int32_t *out_len, uint8_t *out_class, int32_t *out_cut, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if ((rc = check_clip_opts(ctx, o))) return rc;
    if (!out_len) return arg_error(ctx, "out_len is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    return clip_enqueue(ctx, b, width, q_offset, o, out_len, out_class, out_cut, index_base, ctx->stream);
}

// internal hooks for fxg_text.cu (not part of the public header): run K-TRIM / K-FILTER on a caller-chosen stream
extern "C" int fxg_internal_scan_on_stream(fxg_ctx *ctx, int mode, const fxg_batch *b, int q_offset, int thr_q, int min_len,
                                           int min_percent, void *out, void *stream)
{
    int rc = check_batch(ctx, b, false, true, q_offset);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    return scan_enqueue(ctx, mode == 0 ? MODE_TRIM : MODE_FILTER, b, q_offset, thr_q, min_len, min_percent, out, 0, (cudaStream_t)stream);
}
extern "C" void *fxg_internal_counters(fxg_ctx *ctx) { return ctx ? (void *)ctx->d_counters : NULL; }
extern "C" int fxg_internal_revcomp_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *oseq, uint8_t *oqual, void *stream)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    return revcomp_enqueue(ctx, b, q_offset, oseq, oqual, 0, (cudaStream_t)stream);
}
extern "C" int fxg_internal_clip_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, const fxg_clip_opts *o, int32_t *out_len,
                                           uint8_t *out_class, void *stream)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if ((rc = check_clip_opts(ctx, o))) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    return clip_enqueue(ctx, b, NULL, q_offset, o, out_len, out_class, NULL, 0, (cudaStream_t)stream);
}
extern "C" int fxg_internal_stats_on_stream(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist, int32_t max_cycles,
                                            const int32_t *weight_dev, void *stream)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    return stats_enqueue(ctx, b, q_offset, hist, max_cycles, weight_dev, 0, (cudaStream_t)stream);
}

// ---- K-HASH (std::hash<std::string> of every read; collapser routing key) ----------------------------------
extern "C" int fxg_hash_dev(fxg_ctx *ctx, const fxg_batch *b, uint64_t *hash_dev)
{
    int rc = check_batch(ctx, b, true, false, 33);
    if (rc) return rc;
    if (!hash_dev) return arg_error(ctx, "hash_dev is NULL");
    if (b->n == 0) return FXG_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, launch_hash(b->seq, b->len, b->uniform_len, b->stride, b->n, hash_dev, ctx->stream));
    ctx->launches++;
    return FXG_OK;
}

// ---- (f-2) validate / fastq_masker / fastx_artifacts_filter ----------------------------------------------------
static int extra_enqueue(fxg_ctx *ctx, int op, const fxg_batch *b, int q_offset, int thr_q, int mask_char, uint8_t *out_seq,
                         uint8_t *flags, int64_t index_base, cudaStream_t st)
{
    if (b->n == 0) return FXG_OK;
    if (b->qual && (op == 0 || op == 2 || op == 3)) {
        // FASTQ batches of short reads: the same warp-private TMA tile ring as K-TRIM (fxg_kernels.cu k_scan_w), with the op's
        // per-word work in place of the quality compare
        TilePlan plan;
        int rc = make_plan(ctx, b, 2, 0, &plan);
        if (rc) return rc;
        if (plan.warp_ring) {
            ScanParams p;
            memset(&p, 0, sizeof(p));
            p.seq = b->seq; p.qual = b->qual; p.len = b->len; p.uniform_len = b->uniform_len; p.stride = b->stride; p.n = b->n;
            p.tile_reads = plan.tile_reads; p.stages = plan.stages; p.rot_shift = plan.rot_shift;
            p.qk = make_qualk(q_offset, 0);
            p.out = flags; p.index_base = index_base; p.counters = ctx->d_counters;
            CK(ctx, launch_scan(op == 0 ? MODE_VALIDATE : op == 3 ? MODE_HASN : MODE_ARTIFACT, true, plan, p, st));
            ctx->launches++;
            ctx->report.n_in += b->n;
            return FXG_OK;
        }
    }
    CK(ctx, launch_extra(op, b->seq, b->qual, b->len, b->uniform_len, b->stride, b->n, q_offset, thr_q, mask_char, out_seq, flags,
                         index_base, ctx->d_counters, ctx->sm_count, st));
    ctx->launches += (op == 1 || op == 3) ? 2 : 1;
    ctx->report.n_in += b->n;
    return FXG_OK;
}

extern "C" int fxg_validate_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    return extra_enqueue(ctx, 0, b, q_offset, 0, 0, NULL, NULL, index_base, ctx->stream);
}

extern "C" int fxg_mask_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int mask_char, uint8_t *out_seq,
                            uint8_t *masked_flag, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, true, q_offset);
    if (rc) return rc;
    if (!out_seq || !masked_flag || ((uintptr_t)out_seq & 15)) return arg_error(ctx, "out_seq / masked_flag");
    CK(ctx, cudaSetDevice(ctx->device));
    return extra_enqueue(ctx, 1, b, q_offset, min_quality, mask_char, out_seq, masked_flag, index_base, ctx->stream);
}

extern "C" int fxg_artifacts_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *keep, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!keep) return arg_error(ctx, "keep is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    return extra_enqueue(ctx, 2, b, q_offset, 0, 0, NULL, keep, index_base, ctx->stream);
}

extern "C" int fxg_has_n_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *has_n, int64_t index_base)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!has_n) return arg_error(ctx, "has_n is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    return extra_enqueue(ctx, 3, b, q_offset, 0, 0, NULL, has_n, index_base, ctx->stream);
}

// ---- (f-4) barcode splitter ------------------------------------------------------------------------------
static int barcode_check(fxg_ctx *ctx, const fxg_batch *b, const fxg_barcode_table *t)
{
    if (!ctx) return FXG_ERR_ARG;
    if (!b || !t || !b->seq || !b->len) return arg_error(ctx, "barcode: fragments need seq and len");
    if (b->n < 0 || b->stride <= 0 || (b->stride & 15) || b->stride > BC_MAX_STRIDE) {
        snprintf(ctx->err, sizeof(ctx->err), "barcode: stride %d unsupported (16..%d, multiple of 16)", b->stride, BC_MAX_STRIDE);
        return FXG_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)b->seq & 15)) return arg_error(ctx, "slab base must be 16-byte aligned");
    if (!t->entries || !t->entry_len || t->n_entries <= 0 || t->barcode_len <= 0 || t->barcode_len > b->stride || t->allowed_mismatches < 0)
        return arg_error(ctx, "barcode table");
    return FXG_OK;
}

// fragments already in device memory; runs on `st`
static int barcode_enqueue(fxg_ctx *ctx, const uint8_t *dfrag, const int32_t *dlen, int64_t n, int stride, const fxg_barcode_table *t,
                           int32_t *dbest, cudaStream_t st)
{
    if (n == 0) return FXG_OK;
    const size_t tb = (size_t)t->n_entries * stride, lb = (size_t)t->n_entries * sizeof(int32_t);
    const size_t off_len = (tb + 15) & ~(size_t)15, off_mm = (off_len + lb + 15) & ~(size_t)15;
    void *scratch = NULL;
    CK(ctx, cudaMallocAsync(&scratch, off_mm + (size_t)n * sizeof(int32_t), st));
    CK(ctx, cudaMemcpyAsync(scratch, t->entries, tb, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync((char *)scratch + off_len, t->entry_len, lb, cudaMemcpyHostToDevice, st));
    BarcodeParams p;
    memset(&p, 0, sizeof(p));
    p.frag = dfrag; p.flen = dlen; p.n = n; p.stride = stride;
    p.entries = (const uint8_t *)scratch; p.elen = (const int32_t *)((char *)scratch + off_len);
    p.barcode_len = t->barcode_len; p.allowed = t->allowed_mismatches;
    p.best_mm = (int32_t *)((char *)scratch + off_mm); p.best = dbest;
    for (int e0 = 0; e0 < t->n_entries; e0 += BC_TILE_ENTRIES) {       // long lists: several launches, running minimum
        p.e0 = e0; p.e1 = e0 + BC_TILE_ENTRIES < t->n_entries ? e0 + BC_TILE_ENTRIES : t->n_entries;
        p.last = p.e1 == t->n_entries;
        CK(ctx, launch_barcode(p, ctx->sm_count, st));
        ctx->launches++;
    }
    CK(ctx, cudaFreeAsync(scratch, st));
    ctx->report.n_in += n;
    return FXG_OK;
}

extern "C" int fxg_barcode_dev(fxg_ctx *ctx, const fxg_batch *b, const fxg_barcode_table *t, int32_t *best)
{
    int rc = barcode_check(ctx, b, t);
    if (rc) return rc;
    if (!best) return arg_error(ctx, "best is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    return barcode_enqueue(ctx, b->seq, b->len, b->n, b->stride, t, best, ctx->stream);
}

extern "C" int fxg_barcode_host(fxg_ctx *ctx, const fxg_batch *b, const fxg_barcode_table *t, int32_t *best, fxg_report *report)
{
    int rc = barcode_check(ctx, b, t);
    if (rc) return rc;
    if (!best) return arg_error(ctx, "best is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    if ((rc = fxg_report_reset(ctx))) return rc;
    cudaStream_t st = ctx->lane_stream[0];
    if (b->n > 0) {
        const size_t fb = (size_t)b->n * b->stride, lb = (size_t)b->n * sizeof(int32_t);
        void *d = NULL;
        CK(ctx, cudaMallocAsync(&d, fb + 2 * lb, st));
        uint8_t *dfrag = (uint8_t *)d;
        int32_t *dlen = (int32_t *)((char *)d + fb), *dbest = dlen + b->n;
        CK(ctx, cudaMemcpyAsync(dfrag, b->seq, fb, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(dlen, b->len, lb, cudaMemcpyHostToDevice, st));
        if ((rc = barcode_enqueue(ctx, dfrag, dlen, b->n, b->stride, t, dbest, st))) return rc;
        CK(ctx, cudaMemcpyAsync(best, dbest, lb, cudaMemcpyDeviceToHost, st));
        CK(ctx, cudaFreeAsync(d, st));
    }
    if ((rc = refresh_report(ctx, st))) return rc;
    if (report) *report = ctx->report;
    return FXG_OK;
}

// ---- (f-3) fused pipelines, first version ---------------------------------------------------------------------
extern "C" int fxg_pipeline_dev(fxg_ctx *ctx, const fxg_batch *b, int q_offset, const fxg_stage *stages, int n_stages, int32_t *final_len,
                                int64_t *n_survivors)
{
    int rc = check_batch(ctx, b, true, true, q_offset);
    if (rc) return rc;
    if (!stages || n_stages <= 0 || !final_len) return arg_error(ctx, "pipeline: stages / final_len");
    if (b->n >= (1ll << 31)) return arg_error(ctx, "pipeline: more than 2^31 reads in one batch");
    for (int k = 0; k < n_stages; k++) {
        if (stages[k].op == FXG_STAGE_CLIP) {
            // after another stage (or on a ragged batch) the aligner sees mixed lengths: its stale-buffer rows come from one
            // scan over the survivors (fxg_pipeline.cu launch_stale_rows), whose row operator holds at most 160 bases
            if ((k != 0 || b->len) && b->stride > 160) {
                snprintf(ctx->err, sizeof(ctx->err), "pipeline: the clipper on mixed read lengths supports strides up to 160");
                return FXG_ERR_UNSUPPORTED;
            }
            if ((rc = check_clip_opts(ctx, stages[k].clip))) return rc;
        } else if (stages[k].op == FXG_STAGE_COLLAPSE) {
            if (k != n_stages - 1 || !stages[k].collapser) return arg_error(ctx, "pipeline: the collapser must be the last stage and needs a fxg_collapser");
        } else if (stages[k].op == FXG_STAGE_FILTER) {
            if (stages[k].a1 < 0 || stages[k].a1 > 100) return arg_error(ctx, "min_percent must be 0..100");
        } else if (stages[k].op != FXG_STAGE_TRIM) return arg_error(ctx, "pipeline: unknown stage");
    }
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int S = b->stride;
    const int64_t n = b->n;
    if (n_survivors) *n_survivors = 0;
    if (n == 0) return FXG_OK;
    CK(ctx, cudaMemsetAsync(final_len, 0xFF, (size_t)n * sizeof(int32_t), st));

    // per-stage scratch, sized by the input: decision (int32 or bytes), flags, positions, scan workspace, and the survivor
    // count after each stage.  The counts STAY on the device: every later kernel is launched over the bound n and reads its
    // true row count from d_cnt, so the chain is enqueued back to back and the host reads the counts once, at the end.
    const size_t tmp_bytes = pipe_scan_tmp_bytes(n);
    const size_t ib = (((size_t)n * sizeof(int32_t)) + 255) & ~(size_t)255;
    const size_t cb = (((size_t)(n_stages + 1) * sizeof(int64_t)) + 255) & ~(size_t)255;
    void *scr = NULL;
    CK(ctx, cudaMallocAsync(&scr, cb + 3 * ib + ((tmp_bytes + 255) & ~(size_t)255), st));
    int64_t *d_cnt = (int64_t *)scr;                              // d_cnt[k] = reads entering stage k (k >= 1)
    int32_t *d_dec = (int32_t *)((char *)scr + cb), *d_flags = (int32_t *)((char *)scr + cb + ib), *d_pos = (int32_t *)((char *)scr + cb + 2 * ib);
    void *d_tmp = (char *)scr + cb + 3 * ib;
    // working slabs of the survivors (ping-pong), sized by the bound
    void *work[2] = { NULL, NULL };
    uint8_t *wseq[2] = { NULL, NULL }, *wqual[2] = { NULL, NULL };
    int32_t *wlen[2] = { NULL, NULL }, *widx[2] = { NULL, NULL };
    std::vector<int64_t> h_cnt((size_t)n_stages + 1, 0);
    h_cnt[0] = n;

    fxg_batch cur = *b;
    const int32_t *cur_idx = NULL;
    const int64_t *cur_cnt = NULL;       // device count of the reads entering the stage (NULL: all n)
    const int64_t n_in0 = ctx->report.n_in;
    int which = 0, stages_run = 0;
    rc = FXG_OK;
    for (int k = 0; k < n_stages; k++) {
        const fxg_stage &sg = stages[k];
        cur.n = n;
        if (sg.op == FXG_STAGE_COLLAPSE) {
            // the survivors (already compacted, in input order) enter the count map exactly as the next process of the pipe
            // would read them (fastx_collapser.cpp:112-114); their lengths also go back to final_len.  The count map is
            // sized on the host, so this (last) stage is where the survivor count is read back.
            int64_t alive = n;
            cudaError_t e = cudaSuccess;
            if (cur_cnt) {
                e = cudaMemcpyAsync(h_cnt.data() + 1, d_cnt + 1, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                alive = h_cnt[k];
            }
            if (e == cudaSuccess && alive > 0) {
                e = launch_pipe_keep_all(alive, cur.len, cur.uniform_len, cur_idx, final_len, ctx->sm_count, st);
                ctx->launches++;
            }
            if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline collapse stage: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break; }
            if (alive > 0) {
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) {
                    snprintf(ctx->err, sizeof(ctx->err), "pipeline collapse stage: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break;
                }
                fxg_batch kb = cur;
                kb.n = alive;
                kb.qual = NULL;
                rc = fxg_collapse_add_next(sg.collapser, &kb);
                if (rc) snprintf(ctx->err, sizeof(ctx->err), "pipeline collapse stage: %s", fxg_collapse_error(sg.collapser));
            }
            break;
        }
        const bool bytes = sg.op == FXG_STAGE_FILTER;
        if (sg.op == FXG_STAGE_TRIM) rc = scan_enqueue(ctx, MODE_TRIM, &cur, q_offset, sg.a0, sg.a1, 0, d_dec, 0, st, cur_cnt);
        else if (sg.op == FXG_STAGE_FILTER) rc = scan_enqueue(ctx, MODE_FILTER, &cur, q_offset, sg.a0, 0, sg.a1, d_dec, 0, st, cur_cnt);
        else if (!cur.len) rc = clip_enqueue(ctx, &cur, NULL, q_offset, sg.clip, d_dec, NULL, NULL, 0, st, cur_cnt);
        else {
            // mixed lengths: the clipper works on the reference's stale-buffer rows (one scan over the survivors, in order)
            size_t need = 0;
            cudaError_t e2 = launch_stale_rows(cur.seq, cur.len, S, n, cur_cnt, NULL, NULL, NULL, 0, &need, ctx->sm_count, st);
            void *sscr = NULL, *srows = NULL;
            const size_t rb = (((size_t)n * S) + 255) & ~(size_t)255;
            if (e2 == cudaSuccess) e2 = cudaMallocAsync(&sscr, need, st);
            if (e2 == cudaSuccess) e2 = cudaMallocAsync(&srows, rb + (size_t)n * sizeof(int32_t), st);
            if (e2 == cudaSuccess) e2 = launch_stale_rows(cur.seq, cur.len, S, n, cur_cnt, (uint8_t *)srows, (int32_t *)((char *)srows + rb), sscr, need,
                                                          NULL, ctx->sm_count, st);
            if (e2 != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline stale rows: %s", cudaGetErrorString(e2)); rc = FXG_ERR_CUDA; }
            else {
                fxg_batch sb = cur;
                sb.seq = (const uint8_t *)srows;
                rc = clip_enqueue(ctx, &sb, (const int32_t *)((char *)srows + rb), q_offset, sg.clip, d_dec, NULL, NULL, 0, st, cur_cnt);
                ctx->launches += 2;
            }
            if (sscr) cudaFreeAsync(sscr, st);
            if (srows) cudaFreeAsync(srows, st);
        }
        if (rc) break;
        stages_run = k + 1;
        cudaError_t e = launch_pipe_flags_scan(bytes ? NULL : d_dec, bytes ? (const uint8_t *)d_dec : NULL, n, cur_cnt, d_flags, d_pos, d_tmp, tmp_bytes,
                                              d_cnt + (k + 1), ctx->sm_count, st);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline stage %d: %s", k, cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break; }
        ctx->launches += 3;
        if (k == n_stages - 1) {
            e = launch_pipe_scatter(n, cur_cnt, d_flags, bytes ? NULL : d_dec, cur.len, cur.uniform_len, cur_idx, final_len, ctx->sm_count, st);
            if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline scatter: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break; }
            ctx->launches++;
            break;
        }
        // compact the survivors into the other working slab pair
        const int dst = which;
        if (!work[dst]) {
            const size_t sb = (((size_t)n * S) + 255) & ~(size_t)255, lb = (((size_t)n * sizeof(int32_t)) + 255) & ~(size_t)255;
            e = cudaMallocAsync(&work[dst], 2 * sb + 2 * lb, st);
            if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline: cudaMallocAsync: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break; }
            wseq[dst] = (uint8_t *)work[dst]; wqual[dst] = wseq[dst] + sb;
            wlen[dst] = (int32_t *)(wqual[dst] + sb); widx[dst] = (int32_t *)((char *)wlen[dst] + lb);
        }
        e = launch_pipe_gather(cur.seq, cur.qual, S, n, cur_cnt, d_flags, d_pos, bytes ? NULL : d_dec, cur.len, cur.uniform_len, cur_idx, wseq[dst],
                               wqual[dst], wlen[dst], widx[dst], ctx->sm_count, st);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline gather: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; break; }
        ctx->launches++;
        cur.seq = wseq[dst]; cur.qual = wqual[dst]; cur.len = wlen[dst]; cur.uniform_len = 0;
        cur_idx = widx[dst];
        cur_cnt = d_cnt + (k + 1);
        which ^= 1;
    }
    // the one read-back of the call: the survivor counts of all stages
    if (!rc && stages_run > 0) {
        cudaError_t e = cudaMemcpyAsync(h_cnt.data() + 1, d_cnt + 1, (size_t)stages_run * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "pipeline: %s", cudaGetErrorString(e)); rc = FXG_ERR_CUDA; }
    } else cudaStreamSynchronize(st);
    for (int w = 0; w < 2; w++) if (work[w]) cudaFreeAsync(work[w], st);
    cudaFreeAsync(scr, st);
    if (rc) return rc;
    // the kernels counted the bound as their input; the report counts the reads each stage really saw
    ctx->report.n_in = n_in0;
    for (int k = 0; k < stages_run; k++) ctx->report.n_in += h_cnt[k];
    if (n_survivors) *n_survivors = h_cnt[stages_run];
    return refresh_report(ctx, st);
}

// ---- host-buffer pipelines ---------------------------------------------------------------------------
// Chunks of the host slab travel H2D -> kernel -> D2H on PIPE_LANES side streams, so the copy of one
// chunk overlaps the kernel and the result copy of its neighbours (both copy engines busy).
enum { SLOT_SEQ = 0, SLOT_QUAL, SLOT_LEN, SLOT_AUX, SLOT_OUT0, SLOT_OUT1 };

static int lane_reserve(fxg_ctx *ctx, int lane, int slot, size_t bytes)
{
    if (bytes == 0 || ctx->lane_cap[lane][slot] >= bytes) return FXG_OK;
    if (ctx->lane_buf[lane][slot]) cudaFree(ctx->lane_buf[lane][slot]);
    ctx->lane_buf[lane][slot] = NULL;
    ctx->lane_cap[lane][slot] = 0;
    CK(ctx, cudaMalloc(&ctx->lane_buf[lane][slot], bytes));
    ctx->lane_cap[lane][slot] = bytes;
    return FXG_OK;
}

static int64_t chunk_reads(const fxg_batch *b)
{
    // ~48 MB per slab per chunk: large enough to run PCIe at full rate, small enough to pipeline
    int64_t cr = (48ll << 20) / b->stride;
    if (cr < 1) cr = 1;
    const int64_t third = (b->n + PIPE_LANES - 1) / PIPE_LANES;
    if (cr > third && third > 0) cr = third;
    return cr;
}

enum { HOST_TRIM, HOST_FILTER, HOST_REVCOMP, HOST_STATS, HOST_CLIP, HOST_VALIDATE, HOST_MASK, HOST_ARTIFACT, HOST_HASN };

struct HostOp {
    int op;
    int q_offset, a0, a1;
    void *out0, *out1;                 // host result arrays (op specific)
    const int32_t *aux_host;           // stats: weights, clip: matrix widths (may be NULL)
    uint64_t *hist_dev; int32_t max_cycles;
    const fxg_clip_opts *clip;
};

static int host_pipeline(fxg_ctx *ctx, const fxg_batch *b, const HostOp &h, fxg_report *report)
{
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = fxg_report_reset(ctx);
    if (rc) return rc;
    const int64_t cr = chunk_reads(b);
    const size_t S = (size_t)b->stride;
    const bool has_seq = b->seq != NULL, has_qual = b->qual != NULL;
    int lane = 0;
    for (int64_t r0 = 0; r0 < b->n; r0 += cr, lane = (lane + 1) % PIPE_LANES) {
        const int64_t nr = (b->n - r0 < cr) ? (b->n - r0) : cr;
        cudaStream_t st = ctx->lane_stream[lane];   // one lane's copies and kernels are stream-ordered
        size_t o0 = 0, o1 = 0;
        if (h.op == HOST_TRIM) o0 = (size_t)cr * sizeof(int32_t);
        else if (h.op == HOST_FILTER) o0 = (size_t)cr;
        else if (h.op == HOST_REVCOMP) { o0 = (size_t)cr * S; o1 = has_qual ? (size_t)cr * S : 0; }
        else if (h.op == HOST_CLIP) { o0 = (size_t)cr * sizeof(int32_t); o1 = h.out1 ? (size_t)cr : 0; }
        else if (h.op == HOST_MASK) { o0 = (size_t)cr * S; o1 = (size_t)cr; }
        else if (h.op == HOST_ARTIFACT || h.op == HOST_HASN) o0 = (size_t)cr;
        if (has_seq && (rc = lane_reserve(ctx, lane, SLOT_SEQ, (size_t)cr * S))) return rc;
        if (has_qual && (rc = lane_reserve(ctx, lane, SLOT_QUAL, (size_t)cr * S))) return rc;
        if (b->len && (rc = lane_reserve(ctx, lane, SLOT_LEN, (size_t)cr * sizeof(int32_t)))) return rc;
        if (h.aux_host && (rc = lane_reserve(ctx, lane, SLOT_AUX, (size_t)cr * sizeof(int32_t)))) return rc;
        if ((rc = lane_reserve(ctx, lane, SLOT_OUT0, o0))) return rc;
        if ((rc = lane_reserve(ctx, lane, SLOT_OUT1, o1))) return rc;

        uint8_t *dseq = (uint8_t *)ctx->lane_buf[lane][SLOT_SEQ], *dqual = (uint8_t *)ctx->lane_buf[lane][SLOT_QUAL];
        int32_t *dlen = (int32_t *)ctx->lane_buf[lane][SLOT_LEN], *daux = (int32_t *)ctx->lane_buf[lane][SLOT_AUX];
        void *d0 = ctx->lane_buf[lane][SLOT_OUT0], *d1 = ctx->lane_buf[lane][SLOT_OUT1];
        if (has_seq) CK(ctx, cudaMemcpyAsync(dseq, b->seq + (size_t)r0 * S, (size_t)nr * S, cudaMemcpyHostToDevice, st));
        if (has_qual) CK(ctx, cudaMemcpyAsync(dqual, b->qual + (size_t)r0 * S, (size_t)nr * S, cudaMemcpyHostToDevice, st));
        if (b->len) CK(ctx, cudaMemcpyAsync(dlen, b->len + r0, (size_t)nr * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (h.aux_host) CK(ctx, cudaMemcpyAsync(daux, h.aux_host + r0, (size_t)nr * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        fxg_batch db = *b;
        db.seq = has_seq ? dseq : NULL;
        db.qual = has_qual ? dqual : NULL;
        db.len = b->len ? dlen : NULL;
        db.n = nr;
        switch (h.op) {
        case HOST_TRIM:
            if ((rc = scan_enqueue(ctx, MODE_TRIM, &db, h.q_offset, h.a0, h.a1, 0, d0, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((int32_t *)h.out0 + r0, d0, (size_t)nr * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            break;
        case HOST_FILTER:
            if ((rc = scan_enqueue(ctx, MODE_FILTER, &db, h.q_offset, h.a0, 0, h.a1, d0, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((uint8_t *)h.out0 + r0, d0, (size_t)nr, cudaMemcpyDeviceToHost, st));
            break;
        case HOST_REVCOMP:
            if ((rc = revcomp_enqueue(ctx, &db, h.q_offset, (uint8_t *)d0, has_qual ? (uint8_t *)d1 : NULL, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((uint8_t *)h.out0 + (size_t)r0 * S, d0, (size_t)nr * S, cudaMemcpyDeviceToHost, st));
            if (has_qual) CK(ctx, cudaMemcpyAsync((uint8_t *)h.out1 + (size_t)r0 * S, d1, (size_t)nr * S, cudaMemcpyDeviceToHost, st));
            break;
        case HOST_STATS:
            if ((rc = stats_enqueue(ctx, &db, h.q_offset, h.hist_dev, h.max_cycles, h.aux_host ? daux : NULL, r0, st))) return rc;
            break;
        case HOST_CLIP:
            if ((rc = clip_enqueue(ctx, &db, h.aux_host ? daux : NULL, h.q_offset, h.clip, (int32_t *)d0, h.out1 ? (uint8_t *)d1 : NULL,
                                   NULL, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((int32_t *)h.out0 + r0, d0, (size_t)nr * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            if (h.out1) CK(ctx, cudaMemcpyAsync((uint8_t *)h.out1 + r0, d1, (size_t)nr, cudaMemcpyDeviceToHost, st));
            break;
        case HOST_VALIDATE:
            if ((rc = extra_enqueue(ctx, 0, &db, h.q_offset, 0, 0, NULL, NULL, r0, st))) return rc;
            break;
        case HOST_MASK:
            if ((rc = extra_enqueue(ctx, 1, &db, h.q_offset, h.a0, h.a1, (uint8_t *)d0, (uint8_t *)d1, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((uint8_t *)h.out0 + (size_t)r0 * S, d0, (size_t)nr * S, cudaMemcpyDeviceToHost, st));
            if (h.out1) CK(ctx, cudaMemcpyAsync((uint8_t *)h.out1 + r0, d1, (size_t)nr, cudaMemcpyDeviceToHost, st));
            break;
        case HOST_ARTIFACT:
        case HOST_HASN:
            if ((rc = extra_enqueue(ctx, h.op == HOST_ARTIFACT ? 2 : 3, &db, h.q_offset, 0, 0, NULL, (uint8_t *)d0, r0, st))) return rc;
            CK(ctx, cudaMemcpyAsync((uint8_t *)h.out0 + r0, d0, (size_t)nr, cudaMemcpyDeviceToHost, st));
            break;
        }
    }
    for (int l = 0; l < PIPE_LANES; l++) CK(ctx, cudaStreamSynchronize(ctx->lane_stream[l]));
    rc = refresh_report(ctx, ctx->lane_stream[0]);
    if (rc) return rc;
    if (report) *report = ctx->report;
    return FXG_OK;
}

extern "C" int fxg_trim_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int threshold, int min_len,
                             int32_t *out_len, fxg_report *report)
{
    int rc = check_batch(ctx, b, false, true, q_offset);
    if (rc) return rc;
    if (!out_len) return arg_error(ctx, "out_len is NULL");
    HostOp h = {};
    h.op = HOST_TRIM; h.q_offset = q_offset; h.a0 = threshold; h.a1 = min_len; h.out0 = out_len;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_filter_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int min_percent,
                               uint8_t *keep, fxg_report *report)
{
    int rc = check_batch(ctx, b, false, true, q_offset);
    if (rc) return rc;
    if (!keep) return arg_error(ctx, "keep is NULL");
    if (min_percent < 0 || min_percent > 100) return arg_error(ctx, "min_percent must be 0..100");
    HostOp h = {};
    h.op = HOST_FILTER; h.q_offset = q_offset; h.a0 = min_quality; h.a1 = min_percent; h.out0 = keep;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_revcomp_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *out_seq, uint8_t *out_qual,
                                fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!out_seq || (b->qual && !out_qual)) return arg_error(ctx, "output slab is NULL");
    HostOp h = {};
    h.op = HOST_REVCOMP; h.q_offset = q_offset; h.out0 = out_seq; h.out1 = out_qual;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_stats_accum_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint64_t *hist_dev, int32_t max_cycles,
                                    const int32_t *weight_host, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!hist_dev || max_cycles <= 0) return arg_error(ctx, "hist is NULL or max_cycles <= 0");
    HostOp h = {};
    h.op = HOST_STATS; h.q_offset = q_offset; h.hist_dev = hist_dev; h.max_cycles = max_cycles; h.aux_host = weight_host;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_clip_host(fxg_ctx *ctx, const fxg_batch *b, const int32_t *width_host, int q_offset, const fxg_clip_opts *o,
                             int32_t *out_len, uint8_t *out_class, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if ((rc = check_clip_opts(ctx, o))) return rc;
    if (!out_len) return arg_error(ctx, "out_len is NULL");
    HostOp h = {};
    h.op = HOST_CLIP; h.q_offset = q_offset; h.clip = o; h.aux_host = width_host; h.out0 = out_len; h.out1 = out_class;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_validate_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    HostOp h = {};
    h.op = HOST_VALIDATE; h.q_offset = q_offset;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_mask_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, int min_quality, int mask_char, uint8_t *out_seq,
                             uint8_t *masked_flag, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, true, q_offset);
    if (rc) return rc;
    if (!out_seq) return arg_error(ctx, "out_seq is NULL");
    HostOp h = {};
    h.op = HOST_MASK; h.q_offset = q_offset; h.a0 = min_quality; h.a1 = mask_char; h.out0 = out_seq; h.out1 = masked_flag;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_has_n_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *has_n, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!has_n) return arg_error(ctx, "has_n is NULL");
    HostOp h = {};
    h.op = HOST_HASN; h.q_offset = q_offset; h.out0 = has_n;
    return host_pipeline(ctx, b, h, report);
}

extern "C" int fxg_artifacts_host(fxg_ctx *ctx, const fxg_batch *b, int q_offset, uint8_t *keep, fxg_report *report)
{
    int rc = check_batch(ctx, b, true, false, q_offset);
    if (rc) return rc;
    if (!keep) return arg_error(ctx, "keep is NULL");
    HostOp h = {};
    h.op = HOST_ARTIFACT; h.q_offset = q_offset; h.out0 = keep;
    return host_pipeline(ctx, b, h, report);
}
