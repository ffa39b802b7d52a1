// fxg_collapse.cuh — pieces of the collapser shared by the one-GPU object (fxg_collapse.cu) and the multi-GPU one
// (fxg_dcollapse.cu): the dedup-table kernel's parameters, its launchers and the ordering pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxg {

// what travels with a key row through the owner exchange (16 bytes)
struct RowMeta {
    int64_t  first;     // global first-occurrence index candidate
    uint32_t weight;    // count carried by the row (1 for a plain read)
    int32_t  len;
};

struct DedupParams {
    const uint8_t *keys;        // slab: row r at keys + r*stride
    const int32_t *len;         // per-row length (ignored when meta != NULL)
    const RowMeta *meta;        // per-row {first, weight, len} (rows that arrived through the exchange), or NULL
    int32_t stride;
    int64_t row0, n;            // rows [row0, row0+n) are inserted by this launch
    const int32_t *weight;      // per-row (relative to row0) weight, NULL = 1          (meta == NULL)
    const int64_t *first;       // per-row explicit first-occurrence index, NULL = index_base + row   (meta == NULL)
    int64_t index_base;
    uint64_t *hash;             // per-row hash (out)
    unsigned long long *slots;  // table: (tag32 << 32) | (rep_row + 1), 0 = empty
    uint64_t mask;              // table size - 1
    unsigned long long *count;  // per slot
    unsigned long long *firsts; // per slot (min)
    unsigned long long *counters;
};

cudaError_t launch_hash_dedup(const DedupParams &p, cudaStream_t st);
// compact the occupied slots: n_out (device counter, zeroed by the caller) and the four dense arrays
cudaError_t launch_compact(const unsigned long long *slots, const unsigned long long *count, const unsigned long long *firsts,
                           const uint64_t *hash, int64_t nslots, unsigned long long *n_out, uint32_t *u_rep, uint64_t *u_hash,
                           uint64_t *u_first, uint64_t *u_count, cudaStream_t st);
cudaError_t launch_gather_rows(const uint8_t *keys, const int32_t *len, const RowMeta *meta, const uint32_t *rep, const uint32_t *perm,
                               int stride, uint32_t n, uint8_t *out_rows, int32_t *out_len, cudaStream_t st);

}  // namespace fxg

// K-ORDER: perm[k] = index of the unique printed at rank k, from (hash, first, count) of U uniques (device arrays).
// max_first: an upper bound of every `first` value (0 = unknown) — it only sizes the radix passes.
int fxg_order_impl(const uint64_t *d_hash, const uint64_t *d_first, const uint64_t *d_count, uint32_t U, uint32_t *d_perm,
                   uint64_t max_first, cudaStream_t st, char *errbuf, size_t errlen, int64_t *launches);
