// fxg_collapse.cu — fastx_collapser on the GPU:
//   K-HASH   std::hash<std::string> = libstdc++ _Hash_bytes (Murmur-style, seed 0xc70f6907) per read
//   K-DEDUP  exact dedup: open-addressing table of (tag | representative row), full-key compare on tag hit,
//            per-key count (sum of weights) and first-occurrence index (min)
//   K-ORDER  the output order of src/fastx_collapser/fastx_collapser.cpp:116-122 — count descending, ties in
//            REVERSE iteration order of libstdc++'s std::unordered_map<std::string,size_t> — reproduced from
//            (hash, first index, count) alone by replaying the table's rehash epochs as sorts
//            (SURVEY.md Appendix B).  The sorts are CUB radix sorts (library primitive); hashing, dedup and
//            the epoch logic are hand-written.
//
// Reference lines: collapsed_sequences[string(seq)] += get_reads_count()   fastx_collapser.cpp:112-114
//                  copy -> list::sort(by count) -> reverse print           fastx_collapser.cpp:116-122, 80-91
#include <cub/cub.cuh>
#include <stdio.h>
#include <string.h>

#include "fxg.h"
#include "fxg_kernels.cuh"

namespace fxg {

// ---------------------------------------------------------------------------------------------------
// K-HASH: libstdc++ 64-bit _Hash_bytes (libsupc++/hash_bytes.cc), one thread per read
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }

__device__ __forceinline__ uint64_t hash_row(const uint8_t *row, int len)
{
    const uint64_t mul = (0xc6a4a793ull << 32) + 0x5bd1e995ull;
    uint64_t h = 0xc70f6907ull ^ ((uint64_t)len * mul);
    const int n8 = len >> 3;
    const uint2 *p = reinterpret_cast<const uint2 *>(row);      // rows are 16-byte aligned
    for (int k = 0; k < n8; k++) {
        const uint2 v = __ldg(p + k);
        uint64_t d = ((uint64_t)v.y << 32) | v.x;
        d = shift_mix(d * mul) * mul;
        h ^= d;
        h *= mul;
    }
    const int rem = len & 7;
    if (rem) {
        const uint2 v = __ldg(p + n8);
        uint64_t d = ((uint64_t)v.y << 32) | v.x;
        d &= (rem == 8) ? ~0ull : ((1ull << (8 * rem)) - 1ull);
        h ^= d;
        h *= mul;
    }
    h = shift_mix(h) * mul;
    h = shift_mix(h);
    return h;
}

__device__ __forceinline__ bool rows_equal(const uint8_t *a, const uint8_t *b, int len)
{
    const uint4 *pa = reinterpret_cast<const uint4 *>(a), *pb = reinterpret_cast<const uint4 *>(b);
    const int nfull = len >> 4, rem = len & 15;
    for (int c = 0; c < nfull; c++) {
        const uint4 x = __ldg(pa + c), y = __ldg(pb + c);
        if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) return false;
    }
    if (rem) {
        const uint4 x = __ldg(pa + nfull), y = __ldg(pb + nfull);
        const uint32_t xs[4] = { x.x, x.y, x.z, x.w }, ys[4] = { y.x, y.y, y.z, y.w };
#pragma unroll
        for (int w = 0; w < 4; w++)
            if ((xs[w] ^ ys[w]) & head_mask(rem - 4 * w)) return false;
    }
    return true;
}

struct DedupParams {
    const uint8_t *keys;        // owned slab: row r at keys + r*stride
    const int32_t *len;         // per-row length
    int32_t stride;
    int64_t row0, n;            // rows [row0, row0+n) are inserted by this launch
    const int32_t *weight;      // per-row (relative to row0) weight, NULL = 1
    const int64_t *first;       // per-row explicit first-occurrence index, NULL = index_base + row
    int64_t index_base;
    uint64_t *hash;             // per-row hash (out)
    unsigned long long *slots;  // table: (tag32 << 32) | (rep_row + 1), 0 = empty
    uint64_t mask;              // table size - 1
    unsigned long long *count;  // per slot
    unsigned long long *firsts; // per slot (min)
    unsigned long long *counters;
};

__global__ void __launch_bounds__(256) k_hash_dedup(const DedupParams P)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = P.row0 + t;
        const int L = __ldg(P.len + r);
        const uint8_t *row = P.keys + (size_t)r * P.stride;
        // validate bases (the reader would have rejected the record: fastx.c:45-54, 361-364)
        bool bad = (L <= 0 || L > P.stride);
        if (!bad) {
            for (int c = 0; c * 16 < L; c++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row) + c);
                const uint32_t ws[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int w = 0; w < 4; w++)
                    if (seq_bad_bits(ws[w]) & head_mask(L - 16 * c - 4 * w)) bad = true;
            }
        }
        if (bad) { atomicMin(&P.counters[CNT_FIRST_BAD], (unsigned long long)(P.index_base + t)); continue; }

        const uint64_t h = hash_row(row, L);
        P.hash[r] = h;
        const unsigned long long mine = ((h >> 32) << 32) | (unsigned long long)(r + 1);
        uint64_t slot = h & P.mask;
        for (;;) {
            unsigned long long cur = P.slots[slot];
            if (cur == 0ull) {
                cur = atomicCAS(&P.slots[slot], 0ull, mine);
                if (cur == 0ull) break;                               // claimed: this row represents the key
            }
            if ((cur >> 32) == (h >> 32)) {
                const int64_t rep = (int64_t)(cur & 0xFFFFFFFFull) - 1;
                if (__ldg(P.len + rep) == L && rows_equal(P.keys + (size_t)rep * P.stride, row, L)) break;
            }
            slot = (slot + 1) & P.mask;
        }
        const unsigned long long w = P.weight ? (unsigned long long)__ldg(P.weight + t) : 1ull;
        const unsigned long long f = P.first ? (unsigned long long)__ldg(P.first + t) : (unsigned long long)(P.index_base + t);
        atomicAdd(&P.count[slot], w);
        atomicMin(&P.firsts[slot], f);
    }
}

// plain K-HASH (exported for the multi-GPU owner = hash mod G routing and for parity tests)
__global__ void __launch_bounds__(256) k_hash(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n,
                                              uint64_t *out)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int L = len ? __ldg(len + t) : uniform_len;
        out[t] = (L > 0 && L <= stride) ? hash_row(seq + (size_t)t * stride, L) : 0ull;
    }
}

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long *p, unsigned long long v, int64_t n)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

__global__ void __launch_bounds__(256) k_fill_len(int32_t *p, int32_t v, int64_t n)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

// compact the occupied slots into dense arrays (order irrelevant: everything is sorted afterwards)
__global__ void __launch_bounds__(256) k_compact(const unsigned long long *slots, const unsigned long long *count,
                                                 const unsigned long long *firsts, const uint64_t *hash, int64_t nslots,
                                                 unsigned long long *n_out, uint32_t *u_rep, uint64_t *u_hash,
                                                 uint64_t *u_first, uint64_t *u_count)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nslots; t += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long cur = slots[t];
        if (cur == 0ull) continue;
        const unsigned long long pos = atomicAdd(n_out, 1ull);
        const uint32_t rep = (uint32_t)(cur & 0xFFFFFFFFull) - 1u;
        u_rep[pos] = rep;
        u_hash[pos] = hash[rep];
        u_first[pos] = firsts[t];
        u_count[pos] = count[t];
    }
}

// ---------------------------------------------------------------------------------------------------
// K-ORDER helpers
// ---------------------------------------------------------------------------------------------------
__global__ void k_iota(uint32_t *p, uint32_t n) { for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) p[t] = t; }

// seq[p] = id of the unique at position p; bucket[p] = hash[id] % B
__global__ void k_bucket(const uint32_t *seq, const uint64_t *hash, uint64_t B, uint32_t m, uint32_t *bucket, uint32_t *pos)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        bucket[t] = (uint32_t)(hash[seq[t]] % B);
        pos[t] = t;
    }
}
// after sorting (bucket, pos) by bucket (stable): head index of each run (0 elsewhere) for a max-scan
__global__ void k_run_heads(const uint32_t *bucket_sorted, uint32_t m, uint32_t *head)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
        head[t] = (t == 0 || bucket_sorted[t] != bucket_sorted[t - 1]) ? t : 0u;
}
// key = (first-touch position of the bucket << 32) | own position ; value = unique id
__global__ void k_touch_keys(const uint32_t *pos_sorted, const uint32_t *head_scanned, const uint32_t *seq, uint32_t m,
                             uint64_t *key, uint32_t *val)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint32_t p = pos_sorted[t];
        const uint32_t touch = pos_sorted[head_scanned[t]];
        key[t] = ((uint64_t)touch << 32) | p;
        val[t] = seq[p];
    }
}
__global__ void k_reverse(const uint32_t *in, uint32_t *out, uint32_t n)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) out[t] = in[n - 1 - t];
}
__global__ void k_gather_u64(const uint64_t *src, const uint32_t *idx, uint64_t *dst, uint32_t n)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
__global__ void k_gather_rows(const uint8_t *keys, const int32_t *len, const uint32_t *rep, const uint32_t *perm, int stride,
                              uint32_t n, uint8_t *out_rows, int32_t *out_len)
{
    const int chunks = stride >> 4;
    const uint64_t total = (uint64_t)n * chunks;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(t / chunks);
        const int c = (int)(t - (uint64_t)i * chunks);
        const uint32_t r = rep[perm ? perm[i] : i];
        reinterpret_cast<uint4 *>(out_rows + (size_t)i * stride)[c] = __ldg(reinterpret_cast<const uint4 *>(keys + (size_t)r * stride) + c);
        if (c == 0 && out_len) out_len[i] = len[r];
    }
}

struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

}  // namespace fxg

using namespace fxg;

// bucket-count ladder of std::unordered_map growing from empty by single insertions (libstdc++ 13:
// _Prime_rehash_policy::_M_need_rehash/_M_next_bkt, max_load_factor 1.0)
static const uint64_t kLadder[] = { 13ull, 29ull, 59ull, 127ull, 257ull, 541ull, 1109ull, 2357ull, 5087ull, 10273ull, 20753ull,
    42043ull, 85229ull, 172933ull, 351061ull, 712697ull, 1447153ull, 2938679ull, 5967347ull, 12117689ull, 24607243ull,
    49969847ull, 101473717ull, 206062531ull, 418451333ull, 849749479ull, 1725587117ull, 3504151727ull };
static const int kLadderN = (int)(sizeof(kLadder) / sizeof(kLadder[0]));

#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(errbuf, errlen, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            rc = FXG_ERR_CUDA;                                                                     \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

static inline unsigned grid_for(uint64_t n) { uint64_t b = (n + 255) / 256; if (b > 148 * 32) b = 148 * 32; if (b < 1) b = 1; return (unsigned)b; }

// Final permutation of U uniques from (hash, first, count) — usable on one GPU or on the gathered triples
// of many.  perm[k] = index (into the input arrays) of the unique printed at rank k.
int fxg_order_impl(const uint64_t *d_hash, const uint64_t *d_first, const uint64_t *d_count, uint32_t U, uint32_t *d_perm,
                   cudaStream_t st, char *errbuf, size_t errlen, int64_t *launches)
{
    int rc = FXG_OK;
    if (U == 0) return rc;
    uint32_t *ids_a = NULL, *ids_b = NULL, *bucket_a = NULL, *bucket_b = NULL, *pos_a = NULL, *pos_b = NULL, *head = NULL, *ord = NULL;
    uint64_t *key_a = NULL, *key_b = NULL, *cnt_g = NULL, *cnt_s = NULL;
    void *tmp = NULL;
    size_t tmp_bytes = 0, need = 0;
    const unsigned G = grid_for(U);

    CKC(cudaMalloc(&ids_a, (size_t)U * 4)); CKC(cudaMalloc(&ids_b, (size_t)U * 4));
    CKC(cudaMalloc(&bucket_a, (size_t)U * 4)); CKC(cudaMalloc(&bucket_b, (size_t)U * 4));
    CKC(cudaMalloc(&pos_a, (size_t)U * 4)); CKC(cudaMalloc(&pos_b, (size_t)U * 4));
    CKC(cudaMalloc(&head, (size_t)U * 4)); CKC(cudaMalloc(&ord, (size_t)U * 4));
    CKC(cudaMalloc(&key_a, (size_t)U * 8)); CKC(cudaMalloc(&key_b, (size_t)U * 8));
    CKC(cudaMalloc(&cnt_g, (size_t)U * 8)); CKC(cudaMalloc(&cnt_s, (size_t)U * 8));
    // temp storage large enough for every CUB call below at size U
    cub::DeviceRadixSort::SortPairs(NULL, need, key_a, key_b, ids_a, ids_b, (int)U, 0, 64, st); tmp_bytes = need;
    cub::DeviceRadixSort::SortPairsDescending(NULL, need, key_a, key_b, ids_a, ids_b, (int)U, 0, 64, st); if (need > tmp_bytes) tmp_bytes = need;
    cub::DeviceRadixSort::SortPairs(NULL, need, bucket_a, bucket_b, pos_a, pos_b, (int)U, 0, 32, st); if (need > tmp_bytes) tmp_bytes = need;
    cub::DeviceScan::InclusiveScan(NULL, need, head, head, MaxOp(), (int)U, st); if (need > tmp_bytes) tmp_bytes = need;
    CKC(cudaMalloc(&tmp, tmp_bytes + 16));

    // 1. first-occurrence order: ids sorted by `first` ascending  -> ids_b
    k_iota<<<G, 256, 0, st>>>(ids_a, U);
    CKC(cudaMemcpyAsync(key_a, d_first, (size_t)U * 8, cudaMemcpyDeviceToDevice, st));
    need = tmp_bytes;
    CKC(cub::DeviceRadixSort::SortPairs(tmp, need, key_a, key_b, ids_a, ids_b, (int)U, 0, 64, st));
    *launches += 8;
    // ids_b = ids in first-occurrence order; `ord` holds the map's iteration order over the first m keys
    {
        uint32_t m = 0;   // keys already in the map
        for (int k = 0; k < kLadderN && m < U; k++) {
            const uint64_t B = kLadder[k];
            const uint32_t m2 = (B < (uint64_t)U) ? (uint32_t)B : U;      // the epoch ends when size reaches B
            // seq = ord[0..m) ++ ids_b[m..m2)
            if (m2 > m) CKC(cudaMemcpyAsync(ord + m, ids_b + m, (size_t)(m2 - m) * 4, cudaMemcpyDeviceToDevice, st));
            const unsigned g2 = grid_for(m2);
            k_bucket<<<g2, 256, 0, st>>>(ord, d_hash, B, m2, bucket_a, pos_a);
            need = tmp_bytes;
            CKC(cub::DeviceRadixSort::SortPairs(tmp, need, bucket_a, bucket_b, pos_a, pos_b, (int)m2, 0, 32, st));
            k_run_heads<<<g2, 256, 0, st>>>(bucket_b, m2, head);
            need = tmp_bytes;
            CKC(cub::DeviceScan::InclusiveScan(tmp, need, head, head, MaxOp(), (int)m2, st));
            k_touch_keys<<<g2, 256, 0, st>>>(pos_b, head, ord, m2, key_a, ids_a);
            need = tmp_bytes;
            CKC(cub::DeviceRadixSort::SortPairsDescending(tmp, need, key_a, key_b, ids_a, ord, (int)m2, 0, 64, st));
            *launches += 12;
            m = m2;
        }
        if (m < U) { snprintf(errbuf, errlen, "collapser: more than %llu unique sequences", (unsigned long long)kLadder[kLadderN - 1]); rc = FXG_ERR_UNSUPPORTED; goto done; }
    }
    // 2. count descending, ties by iteration position descending: reverse the order, then a stable sort by count
    k_reverse<<<G, 256, 0, st>>>(ord, ids_a, U);
    k_gather_u64<<<G, 256, 0, st>>>(d_count, ids_a, cnt_g, U);
    need = tmp_bytes;
    CKC(cub::DeviceRadixSort::SortPairsDescending(tmp, need, cnt_g, cnt_s, ids_a, d_perm, (int)U, 0, 64, st));
    *launches += 6;
    CKC(cudaStreamSynchronize(st));
done:
    cudaFree(ids_a); cudaFree(ids_b); cudaFree(bucket_a); cudaFree(bucket_b); cudaFree(pos_a); cudaFree(pos_b);
    cudaFree(head); cudaFree(ord); cudaFree(key_a); cudaFree(key_b); cudaFree(cnt_g); cudaFree(cnt_s); cudaFree(tmp);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// collapser object
// ---------------------------------------------------------------------------------------------------
struct fxg_collapser {
    int device;
    cudaStream_t st;
    int32_t stride;
    int64_t cap, rows;            // row capacity / rows stored
    uint8_t *keys;                // [cap][stride]
    int32_t *len;                 // [cap]
    uint64_t *hash;               // [cap]
    uint64_t nslots;
    unsigned long long *slots, *count, *firsts;
    unsigned long long *d_counters;   // CNT_WORDS (own)
    // results of finish()
    int64_t U;
    uint32_t *u_rep, *perm;
    uint64_t *u_hash, *u_first, *u_count;
    int64_t launches;
    char err[256];
};

extern "C" const char *fxg_collapse_error(const fxg_collapser *c) { return c ? c->err : "no collapser"; }
extern "C" int64_t fxg_collapse_launches(const fxg_collapser *c) { return c ? c->launches : 0; }

#define CKO(c, call)                                                                               \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((c)->err, sizeof((c)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return FXG_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

extern "C" void fxg_collapse_free(fxg_collapser *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->keys); cudaFree(c->len); cudaFree(c->hash); cudaFree(c->slots); cudaFree(c->count); cudaFree(c->firsts);
    cudaFree(c->d_counters); cudaFree(c->u_rep); cudaFree(c->perm); cudaFree(c->u_hash); cudaFree(c->u_first); cudaFree(c->u_count);
    if (c->st) cudaStreamDestroy(c->st);
    free(c);
}

extern "C" int fxg_collapse_new(int device, int64_t max_reads, int32_t stride, fxg_collapser **out)
{
    if (!out || max_reads <= 0 || max_reads >= 0xFFFFFFF0ll || stride <= 0 || (stride & 15)) return FXG_ERR_ARG;
    *out = NULL;
    if (cudaSetDevice(device) != cudaSuccess) return FXG_ERR_CUDA;
    fxg_collapser *c = (fxg_collapser *)calloc(1, sizeof(fxg_collapser));
    if (!c) return FXG_ERR_NOMEM;
    c->device = device; c->stride = stride; c->cap = max_reads;
    uint64_t ns = 1024;
    while (ns < (uint64_t)max_reads * 2) ns <<= 1;
    c->nslots = ns;
    const unsigned long long init[CNT_WORDS] = { 0, ~0ull };
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&c->keys, (size_t)max_reads * stride) != cudaSuccess || cudaMalloc(&c->len, (size_t)max_reads * 4) != cudaSuccess ||
        cudaMalloc(&c->hash, (size_t)max_reads * 8) != cudaSuccess || cudaMalloc(&c->slots, ns * 8) != cudaSuccess ||
        cudaMalloc(&c->count, ns * 8) != cudaSuccess || cudaMalloc(&c->firsts, ns * 8) != cudaSuccess ||
        cudaMalloc(&c->d_counters, CNT_WORDS * 8) != cudaSuccess ||
        cudaMemsetAsync(c->slots, 0, ns * 8, c->st) != cudaSuccess || cudaMemsetAsync(c->count, 0, ns * 8, c->st) != cudaSuccess ||
        cudaMemsetAsync(c->firsts, 0xFF, ns * 8, c->st) != cudaSuccess ||
        cudaMemcpyAsync(c->d_counters, init, sizeof(init), cudaMemcpyHostToDevice, c->st) != cudaSuccess ||
        cudaStreamSynchronize(c->st) != cudaSuccess) {
        cudaGetLastError();
        fxg_collapse_free(c);
        return FXG_ERR_NOMEM;
    }
    *out = c;
    return FXG_OK;
}

// Append rows (device or host memory — cudaMemcpyDefault) and insert them.  weight/first: see fxg.h.
extern "C" int fxg_collapse_add(fxg_collapser *c, const fxg_batch *b, const int32_t *weight, const int64_t *first, int64_t index_base)
{
    if (!c || !b || !b->seq || b->n < 0 || b->stride != c->stride) return FXG_ERR_ARG;
    if (c->rows + b->n > c->cap) { snprintf(c->err, sizeof(c->err), "collapser capacity %lld rows exceeded", (long long)c->cap); return FXG_ERR_ARG; }
    if (b->n == 0) return FXG_OK;
    CKO(c, cudaSetDevice(c->device));
    const int64_t row0 = c->rows;
    const size_t S = (size_t)c->stride;
    // stream the rows in chunks so that H2D copies overlap the insert kernel of the previous chunk
    const int64_t chunk = (64ll << 20) / (int64_t)S;
    int32_t *d_w = NULL; int64_t *d_f = NULL;
    if (weight) { CKO(c, cudaMalloc(&d_w, (size_t)b->n * 4)); CKO(c, cudaMemcpyAsync(d_w, weight, (size_t)b->n * 4, cudaMemcpyDefault, c->st)); }
    if (first) { CKO(c, cudaMalloc(&d_f, (size_t)b->n * 8)); CKO(c, cudaMemcpyAsync(d_f, first, (size_t)b->n * 8, cudaMemcpyDefault, c->st)); }
    if (b->len) CKO(c, cudaMemcpyAsync(c->len + row0, b->len, (size_t)b->n * 4, cudaMemcpyDefault, c->st));
    else { k_fill_len<<<grid_for((uint64_t)b->n), 256, 0, c->st>>>(c->len + row0, b->uniform_len, b->n); c->launches++; }
    for (int64_t r = 0; r < b->n; r += chunk) {
        const int64_t nr = (b->n - r < chunk) ? (b->n - r) : chunk;
        CKO(c, cudaMemcpyAsync(c->keys + (size_t)(row0 + r) * S, b->seq + (size_t)r * S, (size_t)nr * S, cudaMemcpyDefault, c->st));
        DedupParams p;
        p.keys = c->keys; p.len = c->len; p.stride = c->stride; p.row0 = row0 + r; p.n = nr;
        p.weight = d_w ? d_w + r : NULL; p.first = d_f ? d_f + r : NULL; p.index_base = index_base + r;
        p.hash = c->hash; p.slots = c->slots; p.mask = c->nslots - 1; p.count = c->count; p.firsts = c->firsts;
        p.counters = c->d_counters;
        k_hash_dedup<<<grid_for((uint64_t)nr), 256, 0, c->st>>>(p);
        CKO(c, cudaGetLastError());
        c->launches++;
    }
    CKO(c, cudaStreamSynchronize(c->st));
    cudaFree(d_w); cudaFree(d_f);
    c->rows += b->n;
    return FXG_OK;
}

// Compact the table; with order != 0 also compute the reference's output order.
extern "C" int fxg_collapse_finish(fxg_collapser *c, int order, int64_t *n_unique, int64_t *first_bad_read)
{
    if (!c) return FXG_ERR_ARG;
    CKO(c, cudaSetDevice(c->device));
    unsigned long long h_cnt[CNT_WORDS];
    CKO(c, cudaMemcpy(h_cnt, c->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost));
    if (first_bad_read) *first_bad_read = (h_cnt[CNT_FIRST_BAD] == ~0ull) ? -1 : (int64_t)h_cnt[CNT_FIRST_BAD];
    unsigned long long *d_n = NULL;
    CKO(c, cudaMalloc(&d_n, 8));
    CKO(c, cudaMemsetAsync(d_n, 0, 8, c->st));
    const int64_t ucap = c->rows > 0 ? c->rows : 1;
    cudaFree(c->u_rep); cudaFree(c->u_hash); cudaFree(c->u_first); cudaFree(c->u_count); cudaFree(c->perm);
    c->u_rep = NULL; c->u_hash = NULL; c->u_first = NULL; c->u_count = NULL; c->perm = NULL;
    CKO(c, cudaMalloc(&c->u_rep, (size_t)ucap * 4)); CKO(c, cudaMalloc(&c->u_hash, (size_t)ucap * 8));
    CKO(c, cudaMalloc(&c->u_first, (size_t)ucap * 8)); CKO(c, cudaMalloc(&c->u_count, (size_t)ucap * 8));
    k_compact<<<grid_for(c->nslots), 256, 0, c->st>>>(c->slots, c->count, c->firsts, c->hash, (int64_t)c->nslots, d_n,
                                                       c->u_rep, c->u_hash, c->u_first, c->u_count);
    c->launches++;
    unsigned long long U = 0;
    CKO(c, cudaMemcpyAsync(&U, d_n, 8, cudaMemcpyDeviceToHost, c->st));
    CKO(c, cudaStreamSynchronize(c->st));
    cudaFree(d_n);
    c->U = (int64_t)U;
    if (n_unique) *n_unique = c->U;
    if (order && U > 0) {
        CKO(c, cudaMalloc(&c->perm, (size_t)U * 4));
        int rc = fxg_order_impl(c->u_hash, c->u_first, c->u_count, (uint32_t)U, c->perm, c->st, c->err, sizeof(c->err), &c->launches);
        if (rc) return rc;
    }
    return FXG_OK;
}

// Copy the uniques out (device or host destinations), in output order when finish(order=1) ran, else in
// table order.  Any destination may be NULL.
extern "C" int fxg_collapse_fetch(fxg_collapser *c, uint8_t *out_seq, int32_t *out_len, uint64_t *out_count, int64_t *out_first,
                                  uint64_t *out_hash)
{
    if (!c) return FXG_ERR_ARG;
    if (c->U == 0) return FXG_OK;
    CKO(c, cudaSetDevice(c->device));
    const uint32_t U = (uint32_t)c->U;
    const unsigned G = grid_for(U);
    uint64_t *tmp = NULL;
    CKO(c, cudaMalloc(&tmp, (size_t)U * 8));
    const uint64_t *srcs[3] = { c->u_count, c->u_first, c->u_hash };
    void *dsts[3] = { out_count, out_first, out_hash };
    for (int k = 0; k < 3; k++) {
        if (!dsts[k]) continue;
        if (c->perm) { k_gather_u64<<<G, 256, 0, c->st>>>(srcs[k], c->perm, tmp, U); c->launches++; }
        CKO(c, cudaMemcpyAsync(dsts[k], c->perm ? tmp : srcs[k], (size_t)U * 8, cudaMemcpyDefault, c->st));
        CKO(c, cudaStreamSynchronize(c->st));
    }
    cudaFree(tmp);
    if (out_seq || out_len) {
        uint8_t *rows = NULL; int32_t *lens = NULL;
        CKO(c, cudaMalloc(&rows, (size_t)U * c->stride)); CKO(c, cudaMalloc(&lens, (size_t)U * 4));
        k_gather_rows<<<grid_for((uint64_t)U * (c->stride >> 4)), 256, 0, c->st>>>(c->keys, c->len, c->u_rep, c->perm, c->stride, U, rows, lens);
        c->launches++;
        if (out_seq) CKO(c, cudaMemcpyAsync(out_seq, rows, (size_t)U * c->stride, cudaMemcpyDefault, c->st));
        if (out_len) CKO(c, cudaMemcpyAsync(out_len, lens, (size_t)U * 4, cudaMemcpyDefault, c->st));
        CKO(c, cudaStreamSynchronize(c->st));
        cudaFree(rows); cudaFree(lens);
    }
    return FXG_OK;
}

// Stand-alone pieces for the multi-GPU path --------------------------------------------------------------
extern "C" int fxg_collapse_order_dev(int device, const uint64_t *hash_dev, const uint64_t *first_dev, const uint64_t *count_dev,
                                      int64_t n_unique, uint32_t *perm_dev)
{
    if (n_unique < 0 || n_unique >= 0xFFFFFFF0ll) return FXG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return FXG_ERR_CUDA;
    char err[256];
    int64_t launches = 0;
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return FXG_ERR_CUDA;
    cudaDeviceSynchronize();   // the triples usually come from another stream (collective output)
    int rc = fxg_order_impl(hash_dev, first_dev, count_dev, (uint32_t)n_unique, perm_dev, st, err, sizeof(err), &launches);
    if (rc) fprintf(stderr, "fxg_collapse_order_dev: %s\n", err);
    cudaStreamDestroy(st);
    return rc;
}

namespace fxg {
cudaError_t launch_hash(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint64_t *out, cudaStream_t st)
{
    k_hash<<<grid_for((uint64_t)n), 256, 0, st>>>(seq, len, uniform_len, stride, n, out);
    return cudaGetLastError();
}
}
