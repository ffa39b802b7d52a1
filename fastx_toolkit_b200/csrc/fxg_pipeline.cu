// fxg_pipeline.cu — SURVEY.md §8(f-3), first version: the map-type tools chained on the device, the reads
// never leaving HBM between stages.  What a user runs today as
//     fastx_clipper ... | fastq_quality_trimmer ... | fastq_quality_filter ...
// becomes one call (fxg_pipeline_dev, fxg_api.cu): each stage is the tool's own kernel (K-CLIP, K-TRIM, K-FILTER,
// unchanged), followed by a compaction of the survivors (flags -> exclusive scan -> row gather) so that the next stage
// sees exactly the records the next process of the pipe would read (fastx.c:440-473 writes, fastx.c:314-404 re-reads).
// All three tools only ever shorten a read from its 3' end, so a record is fully described by (original index, length).
//
// The clipper after a trimming stage (or on a ragged batch) works on mixed lengths with the reference aligner's
// stale-buffer semantics: its rows come from one scan over the survivors (launch_stale_rows below).
// A fastx_collapser stage may end the chain: the survivors are added to a fxg_collapser in input order.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "fxg_kernels.cuh"

namespace fxg {

// In a fused pipeline the number of survivors of a stage stays on the device (n_dev; NULL for the first stage, whose
// count the caller knows): every kernel is launched over the upper bound n, the size of the input batch, and reads the
// true count itself, so that the whole chain is enqueued without a host round trip per stage.
__device__ __forceinline__ int64_t live_rows(const int64_t *n_dev, int64_t n) { return n_dev ? *n_dev : n; }

// survivor flags of one stage: lengths (>= 0 keeps) or keep bytes; rows past the live ones (up to the bound n) get 0
__global__ void __launch_bounds__(256) k_pipe_flags(const int32_t *new_len, const uint8_t *keep, int64_t n, const int64_t *n_dev, int32_t *flags)
{
    const int64_t N = live_rows(n_dev, n);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = i < N ? (new_len ? (new_len[i] >= 0 ? 1 : 0) : (keep[i] ? 1 : 0)) : 0;
}

// the survivors of the stage: exclusive position of the last row + its flag
__global__ void k_pipe_count(const int32_t *flags, const int32_t *pos, int64_t n, int64_t *count_out)
{
    *count_out = (int64_t)pos[n - 1] + flags[n - 1];
}

// survivor i moves to row pos[i] of the destination slabs: one thread per 16-byte chunk of a row
__global__ void __launch_bounds__(256) k_pipe_gather(const uint8_t *src_seq, const uint8_t *src_qual, int stride, int64_t n, const int64_t *n_dev,
                                                     const int32_t *flags, const int32_t *pos, const int32_t *new_len, const int32_t *cur_len,
                                                     int uniform_len, const int32_t *cur_idx, uint8_t *dst_seq, uint8_t *dst_qual, int32_t *dst_len,
                                                     int32_t *dst_idx)
{
    const int chunks = stride >> 4;
    const int64_t total = live_rows(n_dev, n) * chunks;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        if (!flags[i]) continue;
        const int64_t d = pos[i];
        const size_t so = (size_t)i * stride + (size_t)c * 16, dof = (size_t)d * stride + (size_t)c * 16;
        *reinterpret_cast<uint4 *>(dst_seq + dof) = __ldg(reinterpret_cast<const uint4 *>(src_seq + so));
        *reinterpret_cast<uint4 *>(dst_qual + dof) = __ldg(reinterpret_cast<const uint4 *>(src_qual + so));
        if (c == 0) {
            dst_len[d] = new_len ? new_len[i] : (cur_len ? cur_len[i] : uniform_len);
            dst_idx[d] = cur_idx ? cur_idx[i] : (int32_t)i;
        }
    }
}

// last stage: the survivors' lengths go back to their original positions (final_len was preset to -1)
__global__ void __launch_bounds__(256) k_pipe_scatter(int64_t n, const int64_t *n_dev, const int32_t *flags, const int32_t *new_len,
                                                      const int32_t *cur_len, int uniform_len, const int32_t *cur_idx, int32_t *final_len)
{
    const int64_t N = live_rows(n_dev, n);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) final_len[cur_idx ? cur_idx[i] : i] = new_len ? new_len[i] : (cur_len ? cur_len[i] : uniform_len);
}

// a collapser stage keeps every read it is given: final_len = current length
__global__ void __launch_bounds__(256) k_pipe_keep_all(int64_t n, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx, int32_t *final_len)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        final_len[cur_idx ? cur_idx[i] : i] = cur_len ? cur_len[i] : uniform_len;
}

static unsigned pgrid(int64_t n, int sm) { int64_t b = (n + 255) / 256; if (b > (int64_t)sm * 16) b = (int64_t)sm * 16; if (b < 1) b = 1; return (unsigned)b; }

size_t pipe_scan_tmp_bytes(int64_t n)
{
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(NULL, need, (const int32_t *)NULL, (int32_t *)NULL, (int)n);
    return need;
}

// flags + exclusive scan over the bound n; the survivor count lands in *count_out (device memory)
cudaError_t launch_pipe_flags_scan(const int32_t *new_len, const uint8_t *keep, int64_t n, const int64_t *n_dev, int32_t *flags, int32_t *pos,
                                   void *tmp, size_t tmp_bytes, int64_t *count_out, int sm_count, cudaStream_t st)
{
    k_pipe_flags<<<pgrid(n, sm_count), 256, 0, st>>>(new_len, keep, n, n_dev, flags);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, pos, (int)n, st);
    if (e != cudaSuccess) return e;
    k_pipe_count<<<1, 1, 0, st>>>(flags, pos, n, count_out);
    return cudaGetLastError();
}

cudaError_t launch_pipe_gather(const uint8_t *src_seq, const uint8_t *src_qual, int stride, int64_t n, const int64_t *n_dev, const int32_t *flags,
                               const int32_t *pos, const int32_t *new_len, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx,
                               uint8_t *dst_seq, uint8_t *dst_qual, int32_t *dst_len, int32_t *dst_idx, int sm_count, cudaStream_t st)
{
    k_pipe_gather<<<pgrid(n * (stride >> 4), sm_count), 256, 0, st>>>(src_seq, src_qual, stride, n, n_dev, flags, pos, new_len, cur_len, uniform_len,
                                                                     cur_idx, dst_seq, dst_qual, dst_len, dst_idx);
    return cudaGetLastError();
}

cudaError_t launch_pipe_keep_all(int64_t n, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx, int32_t *final_len, int sm_count,
                                 cudaStream_t st)
{
    k_pipe_keep_all<<<pgrid(n, sm_count), 256, 0, st>>>(n, cur_len, uniform_len, cur_idx, final_len);
    return cudaGetLastError();
}

cudaError_t launch_pipe_scatter(int64_t n, const int64_t *n_dev, const int32_t *flags, const int32_t *new_len, const int32_t *cur_len, int uniform_len,
                                const int32_t *cur_idx, int32_t *final_len, int sm_count, cudaStream_t st)
{
    k_pipe_scatter<<<pgrid(n, sm_count), 256, 0, st>>>(n, n_dev, flags, new_len, cur_len, uniform_len, cur_idx, final_len);
    return cudaGetLastError();
}

}  // namespace fxg

// ------------------------------------------------------------------------------------------------------------------
// The rows the reference's aligner sees for MIXED-length reads
// (SURVEY Appendix D.1), needed when the clipper is not the first stage.  Its query buffer only grows and keeps stale
// bytes, so row i = the buffer after reads 0..i: read i's bases, a NUL, then whatever longer earlier reads left behind,
// up to the running maximum length.  "Overwrite a prefix" is closed under composition and associative
// (tests/test_stale_scan_model.py), so all rows come from ONE inclusive scan over the reads.
// ------------------------------------------------------------------------------------------------------------------
namespace fxg {

template <int Q16>                 // a row of Q16 * 16 bytes + the length of the valid prefix
struct __align__(16) StaleRow {
    uint4 d[Q16];
    int32_t p;
    int32_t pad[3];
};

template <int Q16>
struct StaleCompose {              // (g after f): g's prefix wins, f shows through behind it
    __device__ __forceinline__ StaleRow<Q16> operator()(const StaleRow<Q16> &f, const StaleRow<Q16> &g) const
    {
        StaleRow<Q16> r;
        r.p = f.p > g.p ? f.p : g.p;
        r.pad[0] = r.pad[1] = r.pad[2] = 0;
        const uint8_t *fb = reinterpret_cast<const uint8_t *>(f.d), *gb = reinterpret_cast<const uint8_t *>(g.d);
#pragma unroll
        for (int q = 0; q < Q16; q++) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = 16 * q + 4 * k;                      // first byte of this 32-bit word
                const uint32_t fw = reinterpret_cast<const uint32_t *>(fb)[x >> 2], gw = reinterpret_cast<const uint32_t *>(gb)[x >> 2];
                const uint32_t m = head_mask(g.p - x);             // bytes x .. x+3 that lie inside g's prefix
                w[k] = (gw & m) | (fw & ~m);
            }
            r.d[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return r;
    }
};

template <int Q16>
struct StaleLoad {                 // read i as an operator: (bases + NUL, len + 1)
    const uint8_t *seq;
    const int32_t *len;
    int stride;
    const int64_t *n_dev;          // rows past the live ones (the scan runs over the bound) are the identity: prefix 0
    __device__ __forceinline__ StaleRow<Q16> operator()(int64_t i) const
    {
        StaleRow<Q16> r;
        const bool live = !n_dev || i < *n_dev;
        const int L = live ? len[i] : -1;
        r.p = L + 1;
        r.pad[0] = r.pad[1] = r.pad[2] = 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(seq + (size_t)i * stride);
#pragma unroll
        for (int q = 0; q < Q16; q++) {
            uint4 v = (live && 16 * q < stride) ? __ldg(src + q) : make_uint4(0, 0, 0, 0);
            uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] &= head_mask(L - (16 * q + 4 * k));      // NUL at L and nothing behind it
            r.d[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return r;
    }
};

template <int Q16>
__global__ void __launch_bounds__(256) k_stale_split(const StaleRow<Q16> *rows, int64_t n, const int64_t *n_dev, int stride, uint8_t *out_seq,
                                                     int32_t *out_width)
{
    const int chunks = stride >> 4;
    const int64_t total = (n_dev ? *n_dev : n) * chunks;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        *reinterpret_cast<uint4 *>(out_seq + (size_t)i * stride + 16 * c) = rows[i].d[c];
        if (c == 0) out_width[i] = rows[i].p - 1;              // running maximum read length = the aligner's matrix width
    }
}

template <int Q16>
static cudaError_t stale_rows_t(const uint8_t *seq, const int32_t *len, int stride, int64_t n, const int64_t *n_dev, uint8_t *out_seq,
                                int32_t *out_width, void *scratch, size_t scratch_bytes, size_t *need, int sm_count, cudaStream_t st)
{
    typedef StaleRow<Q16> Row;
    StaleLoad<Q16> ld = { seq, len, stride, n_dev };
    thrust::transform_iterator<StaleLoad<Q16>, thrust::counting_iterator<int64_t>, Row, Row> in(thrust::counting_iterator<int64_t>(0), ld);
    const size_t rows_bytes = (((size_t)n * sizeof(Row)) + 255) & ~(size_t)255;
    size_t tmp = 0;
    cudaError_t e = cub::DeviceScan::InclusiveScan(NULL, tmp, in, (Row *)NULL, StaleCompose<Q16>(), (int)n, st);
    if (e != cudaSuccess) return e;
    if (need) { *need = rows_bytes + tmp; return cudaSuccess; }
    if (scratch_bytes < rows_bytes + tmp) return cudaErrorInvalidValue;
    Row *rows = (Row *)scratch;
    e = cub::DeviceScan::InclusiveScan((char *)scratch + rows_bytes, tmp, in, rows, StaleCompose<Q16>(), (int)n, st);
    if (e != cudaSuccess) return e;
    k_stale_split<Q16><<<pgrid(n * (stride >> 4), sm_count), 256, 0, st>>>(rows, n, n_dev, stride, out_seq, out_width);
    return cudaGetLastError();
}

// need != NULL: only report the scratch size.  Strides of 16..160 bytes (reads up to 159 bases + the NUL).
cudaError_t launch_stale_rows(const uint8_t *seq, const int32_t *len, int stride, int64_t n, const int64_t *n_dev, uint8_t *out_seq,
                              int32_t *out_width, void *scratch, size_t scratch_bytes, size_t *need, int sm_count, cudaStream_t st)
{
    switch (stride >> 4) {
    case 1: case 2: return stale_rows_t<2>(seq, len, stride, n, n_dev, out_seq, out_width, scratch, scratch_bytes, need, sm_count, st);
    case 3: case 4: return stale_rows_t<4>(seq, len, stride, n, n_dev, out_seq, out_width, scratch, scratch_bytes, need, sm_count, st);
    case 5: case 6: case 7: return stale_rows_t<7>(seq, len, stride, n, n_dev, out_seq, out_width, scratch, scratch_bytes, need, sm_count, st);
    case 8: case 9: case 10: return stale_rows_t<10>(seq, len, stride, n, n_dev, out_seq, out_width, scratch, scratch_bytes, need, sm_count, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fxg
