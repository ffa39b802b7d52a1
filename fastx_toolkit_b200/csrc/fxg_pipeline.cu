// fxg_pipeline.cu — SURVEY.md §8(f-3), first version: the map-type tools chained on the device, the reads
// never leaving HBM between stages.  What a user runs today as
//     fastx_clipper ... | fastq_quality_trimmer ... | fastq_quality_filter ...
// becomes one call (fxg_pipeline_dev, fxg_api.cu): each stage is the tool's own kernel (K-CLIP, K-TRIM, K-FILTER,
// unchanged), followed by a compaction of the survivors (flags -> exclusive scan -> row gather) so that the next stage
// sees exactly the records the next process of the pipe would read (fastx.c:440-473 writes, fastx.c:314-404 re-reads).
// All three tools only ever shorten a read from its 3' end, so a record is fully described by (original index, length).
//
// Limits of this version: the clipper may only be the FIRST stage and only on batches of one read length — after a
// trimming stage the reference's aligner works on mixed lengths with its stale-buffer semantics,
// which needs a scan over the reads that is not written yet (the target is pinned: tests/test_pipeline_oracle.py).
#include <cub/cub.cuh>

#include "fxg_kernels.cuh"

namespace fxg {

// survivor flags of one stage: lengths (>= 0 keeps) or keep bytes
__global__ void __launch_bounds__(256) k_pipe_flags(const int32_t *new_len, const uint8_t *keep, int64_t n, int32_t *flags)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = new_len ? (new_len[i] >= 0 ? 1 : 0) : (keep[i] ? 1 : 0);
}

// survivor i moves to row pos[i] of the destination slabs: one thread per 16-byte chunk of a row
__global__ void __launch_bounds__(256) k_pipe_gather(const uint8_t *src_seq, const uint8_t *src_qual, int stride, int64_t n, const int32_t *flags,
                                                     const int32_t *pos, const int32_t *new_len, const int32_t *cur_len, int uniform_len,
                                                     const int32_t *cur_idx, uint8_t *dst_seq, uint8_t *dst_qual, int32_t *dst_len, int32_t *dst_idx)
{
    const int chunks = stride >> 4;
    const int64_t total = n * chunks;
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
__global__ void __launch_bounds__(256) k_pipe_scatter(int64_t n, const int32_t *flags, const int32_t *new_len, const int32_t *cur_len, int uniform_len,
                                                      const int32_t *cur_idx, int32_t *final_len)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) final_len[cur_idx ? cur_idx[i] : i] = new_len ? new_len[i] : (cur_len ? cur_len[i] : uniform_len);
}

static unsigned pgrid(int64_t n, int sm) { int64_t b = (n + 255) / 256; if (b > (int64_t)sm * 16) b = (int64_t)sm * 16; if (b < 1) b = 1; return (unsigned)b; }

size_t pipe_scan_tmp_bytes(int64_t n)
{
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(NULL, need, (const int32_t *)NULL, (int32_t *)NULL, (int)n);
    return need;
}

cudaError_t launch_pipe_flags_scan(const int32_t *new_len, const uint8_t *keep, int64_t n, int32_t *flags, int32_t *pos, void *tmp, size_t tmp_bytes,
                                   int sm_count, cudaStream_t st)
{
    k_pipe_flags<<<pgrid(n, sm_count), 256, 0, st>>>(new_len, keep, n, flags);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, pos, (int)n, st);
}

cudaError_t launch_pipe_gather(const uint8_t *src_seq, const uint8_t *src_qual, int stride, int64_t n, const int32_t *flags, const int32_t *pos,
                               const int32_t *new_len, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx, uint8_t *dst_seq,
                               uint8_t *dst_qual, int32_t *dst_len, int32_t *dst_idx, int sm_count, cudaStream_t st)
{
    k_pipe_gather<<<pgrid(n * (stride >> 4), sm_count), 256, 0, st>>>(src_seq, src_qual, stride, n, flags, pos, new_len, cur_len, uniform_len, cur_idx,
                                                                     dst_seq, dst_qual, dst_len, dst_idx);
    return cudaGetLastError();
}

cudaError_t launch_pipe_scatter(int64_t n, const int32_t *flags, const int32_t *new_len, const int32_t *cur_len, int uniform_len, const int32_t *cur_idx,
                                int32_t *final_len, int sm_count, cudaStream_t st)
{
    k_pipe_scatter<<<pgrid(n, sm_count), 256, 0, st>>>(n, flags, new_len, cur_len, uniform_len, cur_idx, final_len);
    return cudaGetLastError();
}

}  // namespace fxg
