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
#include "fxg_collapse.cuh"

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

__global__ void __launch_bounds__(256) k_hash_dedup(const DedupParams P)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = P.row0 + t;
        RowMeta m;
        if (P.meta) m = P.meta[r];
        const int L = P.meta ? m.len : __ldg(P.len + r);
        const unsigned long long f = P.meta ? (unsigned long long)m.first
                                            : (P.first ? (unsigned long long)__ldg(P.first + t) : (unsigned long long)(P.index_base + t));
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
        if (bad) { atomicMin(&P.counters[CNT_FIRST_BAD], P.meta ? f : (unsigned long long)(P.index_base + t)); continue; }

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
                const int Lr = P.meta ? P.meta[rep].len : __ldg(P.len + rep);
                if (Lr == L && rows_equal(P.keys + (size_t)rep * P.stride, row, L)) break;
            }
            slot = (slot + 1) & P.mask;
        }
        const unsigned long long w = P.meta ? (unsigned long long)m.weight : (P.weight ? (unsigned long long)__ldg(P.weight + t) : 1ull);
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

// rows from one stride to another (the table grows its stride when a longer read arrives; shorter batches are padded)
__global__ void __launch_bounds__(256) k_copy_rows(const uint8_t *src, int src_stride, uint8_t *dst, int dst_stride, int64_t n)
{
    const int chunks = (src_stride < dst_stride ? src_stride : dst_stride) >> 4;
    const int64_t total = n * chunks;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        reinterpret_cast<uint4 *>(dst + (size_t)i * dst_stride)[c] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)i * src_stride) + c);
    }
}
// move every occupied slot of the old table into a larger one (keys are distinct: first empty slot on the probe path)
__global__ void __launch_bounds__(256) k_rehash(const unsigned long long *slots, const unsigned long long *count, const unsigned long long *firsts,
                                                const uint64_t *hash, int64_t nslots, unsigned long long *nslots_new_slots, uint64_t mask_new,
                                                unsigned long long *count_new, unsigned long long *firsts_new)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nslots; t += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long cur = slots[t];
        if (cur == 0ull) continue;
        const uint64_t h = hash[(cur & 0xFFFFFFFFull) - 1ull];
        uint64_t slot = h & mask_new;
        while (atomicCAS(&nslots_new_slots[slot], 0ull, cur) != 0ull) slot = (slot + 1) & mask_new;
        count_new[slot] = count[t];
        firsts_new[slot] = firsts[t];
    }
}

// compact the occupied slots into dense arrays (order irrelevant: everything is sorted afterwards).  There is ONE position
// counter for the whole table, and atomics on one address are serialised (~2.5 ns each), so a block takes the positions for a
// tile of 256 x CP_ITEMS slots with a single atomicAdd: ballot + popc inside the warp, the eight warp totals through shared
// memory.
constexpr int CP_ITEMS = 8;
__global__ void __launch_bounds__(256) k_compact(const unsigned long long *slots, const unsigned long long *count,
                                                 const unsigned long long *firsts, const uint64_t *hash, int64_t nslots,
                                                 unsigned long long *n_out, uint32_t *u_rep, uint64_t *u_hash,
                                                 uint64_t *u_first, uint64_t *u_count)
{
    __shared__ unsigned s_warp[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tile = 256 * CP_ITEMS;
    for (int64_t t0 = (int64_t)blockIdx.x * tile; t0 < nslots; t0 += (int64_t)gridDim.x * tile) {
        unsigned long long cur[CP_ITEMS];
        unsigned occ[CP_ITEMS];
        unsigned wtotal = 0;
#pragma unroll
        for (int i = 0; i < CP_ITEMS; i++) {
            const int64_t t = t0 + i * 256 + threadIdx.x;
            cur[i] = t < nslots ? slots[t] : 0ull;
            occ[i] = __ballot_sync(0xffffffffu, cur[i] != 0ull);
            wtotal += (unsigned)__popc(occ[i]);
        }
        if (lane == 0) s_warp[w] = wtotal;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned total = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { const unsigned c = s_warp[k]; s_warp[k] = total; total += c; }
            s_base = total ? atomicAdd(n_out, (unsigned long long)total) : 0ull;
        }
        __syncthreads();
        unsigned long long pos = s_base + s_warp[w];
#pragma unroll
        for (int i = 0; i < CP_ITEMS; i++) {
            if (cur[i] != 0ull) {
                const int64_t t = t0 + i * 256 + threadIdx.x;
                const unsigned long long q = pos + (unsigned long long)__popc(occ[i] & ((1u << lane) - 1u));
                const uint32_t rep = (uint32_t)(cur[i] & 0xFFFFFFFFull) - 1u;
                u_rep[q] = rep;
                u_hash[q] = hash[rep];
                u_first[q] = firsts[t];
                u_count[q] = count[t];
            }
            pos += (unsigned long long)__popc(occ[i]);
        }
        __syncthreads();              // s_warp / s_base are rewritten by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------
// K-ORDER helpers
// ---------------------------------------------------------------------------------------------------
__global__ void k_iota(uint32_t *p, uint32_t n) { for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) p[t] = t; }

// One epoch of the map's life (bucket count B, node sequence seq[0..m)): a node goes to the front of its bucket's chain,
// or to the global list head when the bucket is still empty.  So the epoch's result is seq sorted by (first-touch position
// of the node's bucket: descending, own position: descending) = the REVERSE of a stable ascending sort by the bucket's
// first-touch position alone (the input already is in position order).  k_touch_min finds that position per bucket
// (atomicMin into a table over the B buckets), k_touch_key reads it back as the 32-bit sort key.
__global__ void __launch_bounds__(256) k_touch_min(const uint32_t *seq, const uint64_t *hash, uint64_t B, uint32_t m, uint32_t *bucket, uint32_t *touch)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(hash[seq[t]] % B);
        bucket[t] = b;
        atomicMin(&touch[b], t);
    }
}
__global__ void __launch_bounds__(256) k_touch_key(const uint32_t *bucket, const uint32_t *touch, uint32_t m, uint32_t *key)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) key[t] = touch[bucket[t]];
}
// The same two steps for bucket tables that do not fit the L2 cache (B * 4 bytes against 126 MB): random atomicMin into
// DRAM-resident memory runs at a fraction of the L2 rate.  The nodes are first PARTITIONED by the top bits of their bucket
// number (one stable radix pass over (bucket, position << 32 | node)); the atomics and the read-back then walk the table
// region by region, a few MB at a time.  Nodes with equal sort keys share a bucket, hence a partition, and the stable pass
// keeps their positions ascending, so the stable sort by key that follows yields exactly the order of the direct version
// (tests/test_order_partition_model.py).
__global__ void __launch_bounds__(256) k_bucket_pos(const uint32_t *seq, const uint64_t *hash, uint64_t B, uint32_t m, uint32_t *bucket,
                                                    unsigned long long *pos_node)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) {
        const uint32_t node = seq[t];
        bucket[t] = (uint32_t)(hash[node] % B);
        pos_node[t] = ((unsigned long long)t << 32) | node;
    }
}
__global__ void __launch_bounds__(256) k_touch_min_part(const uint32_t *bucket_p, const unsigned long long *pos_node_p, uint32_t m, uint32_t *touch)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x)
        atomicMin(&touch[bucket_p[p]], (uint32_t)(pos_node_p[p] >> 32));
}
__global__ void __launch_bounds__(256) k_touch_key_part(const uint32_t *bucket_p, const unsigned long long *pos_node_p, const uint32_t *touch,
                                                        uint32_t m, uint32_t *key, uint32_t *node)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x) {
        key[p] = touch[bucket_p[p]];
        node[p] = (uint32_t)pos_node_p[p];
    }
}
__global__ void __launch_bounds__(256) k_reverse(const uint32_t *in, uint32_t *out, uint32_t n)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) out[t] = in[n - 1 - t];
}
__global__ void __launch_bounds__(256) k_gather_u64(const uint64_t *src, const uint32_t *idx, uint64_t *dst, uint32_t n)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
__global__ void __launch_bounds__(256) k_max_u64(const uint64_t *src, uint32_t n, unsigned long long *out)
{
    unsigned long long m = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) { const unsigned long long v = src[t]; m = v > m ? v : m; }
    for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o); m = v > m ? v : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
__global__ void k_gather_rows(const uint8_t *keys, const int32_t *len, const RowMeta *meta, const uint32_t *rep, const uint32_t *perm, int stride,
                              uint32_t n, uint8_t *out_rows, int32_t *out_len)
{
    const int chunks = stride >> 4;
    const uint64_t total = (uint64_t)n * chunks;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(t / chunks);
        const int c = (int)(t - (uint64_t)i * chunks);
        const uint32_t r = rep[perm ? perm[i] : i];
        reinterpret_cast<uint4 *>(out_rows + (size_t)i * stride)[c] = __ldg(reinterpret_cast<const uint4 *>(keys + (size_t)r * stride) + c);
        if (c == 0 && out_len) out_len[i] = meta ? meta[r].len : len[r];
    }
}

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

// bucket tables up to this size take the direct atomicMin path (they stay in the 126 MB L2 next to the streams beside them)
static const uint64_t kTouchDirectBytes = 32ull << 20;

static inline unsigned grid_for(uint64_t n) { uint64_t b = (n + 255) / 256; if (b > 148 * 32) b = 148 * 32; if (b < 1) b = 1; return (unsigned)b; }
static inline int bits_for(uint64_t max_value) { int b = 1; while (b < 64 && (max_value >> b)) b++; return b; }

// scratch of the ordering pass: stream-ordered allocations from the device's default pool (kept cached between calls)
static void pool_keep_memory(void)
{
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
}

// Final permutation of U uniques from (hash, first, count) — usable on one GPU or on the gathered triples
// of many.  perm[k] = index (into the input arrays) of the unique printed at rank k.
int fxg_order_impl(const uint64_t *d_hash, const uint64_t *d_first, const uint64_t *d_count, uint32_t U, uint32_t *d_perm,
                   uint64_t max_first, cudaStream_t st, char *errbuf, size_t errlen, int64_t *launches)
{
    int rc = FXG_OK;
    if (U == 0) return rc;
    uint32_t *ids_a = NULL, *ids_b = NULL, *ord = NULL, *ord_s = NULL, *bucket = NULL, *key = NULL, *key_s = NULL, *touch = NULL;
    uint64_t *k64_a = NULL, *k64_b = NULL;
    unsigned long long *d_max = NULL, h_max = 0;
    void *tmp = NULL;
    size_t tmp_bytes = 0, need = 0;
    const unsigned G = grid_for(U);
    const int ubits = bits_for(U);
    const int fbits = max_first ? bits_for(max_first) : 64;
    uint64_t Bmax = kLadder[kLadderN - 1];
    for (int k = 0; k < kLadderN; k++) if (kLadder[k] >= (uint64_t)U) { Bmax = kLadder[k]; break; }
    if (Bmax < (uint64_t)U) { snprintf(errbuf, errlen, "collapser: more than %llu unique sequences", (unsigned long long)kLadder[kLadderN - 1]); return FXG_ERR_UNSUPPORTED; }

    pool_keep_memory();
    CKC(cudaMallocAsync(&ids_a, (size_t)U * 4, st)); CKC(cudaMallocAsync(&ids_b, (size_t)U * 4, st));
    CKC(cudaMallocAsync(&ord, (size_t)U * 4, st)); CKC(cudaMallocAsync(&ord_s, (size_t)U * 4, st));
    CKC(cudaMallocAsync(&bucket, (size_t)U * 4, st)); CKC(cudaMallocAsync(&key, (size_t)U * 4, st)); CKC(cudaMallocAsync(&key_s, (size_t)U * 4, st));
    CKC(cudaMallocAsync(&touch, (size_t)Bmax * 4, st));
    CKC(cudaMallocAsync(&k64_a, (size_t)U * 8, st)); CKC(cudaMallocAsync(&k64_b, (size_t)U * 8, st));
    CKC(cudaMallocAsync(&d_max, 8, st));
    // temp storage large enough for every CUB call below at size U
    cub::DeviceRadixSort::SortPairs(NULL, need, k64_a, k64_b, ids_a, ids_b, (int)U, 0, 64, st); tmp_bytes = need;
    cub::DeviceRadixSort::SortPairsDescending(NULL, need, k64_a, k64_b, ids_a, ids_b, (int)U, 0, 64, st); if (need > tmp_bytes) tmp_bytes = need;
    cub::DeviceRadixSort::SortPairs(NULL, need, key, key_s, ord, ord_s, (int)U, 0, 32, st); if (need > tmp_bytes) tmp_bytes = need;
    cub::DeviceRadixSort::SortPairs(NULL, need, bucket, key_s, (unsigned long long *)k64_a, (unsigned long long *)k64_b, (int)U, 24, 32, st);
    if (need > tmp_bytes) tmp_bytes = need;
    CKC(cudaMallocAsync(&tmp, tmp_bytes + 16, st));

    // 1. first-occurrence order: ids sorted by `first` ascending  -> ids_b
    k_iota<<<G, 256, 0, st>>>(ids_a, U);
    CKC(cudaMemcpyAsync(k64_a, d_first, (size_t)U * 8, cudaMemcpyDeviceToDevice, st));
    need = tmp_bytes;
    CKC(cub::DeviceRadixSort::SortPairs(tmp, need, k64_a, k64_b, ids_a, ids_b, (int)U, 0, fbits, st));
    // the largest count sizes the last sort's passes
    CKC(cudaMemsetAsync(d_max, 0, 8, st));
    k_max_u64<<<G, 256, 0, st>>>(d_count, U, d_max);
    CKC(cudaMemcpyAsync(&h_max, d_max, 8, cudaMemcpyDeviceToHost, st));
    *launches += 2 + (fbits + 7) / 8 + 1;
    // 2. the epochs.  ord[0..m) = the map's iteration order over the first m keys; ord_s = its reverse after the last epoch
    {
        uint32_t m = 0;   // keys already in the map
        for (int k = 0; k < kLadderN && m < U; k++) {
            const uint64_t B = kLadder[k];
            const uint32_t m2 = (B < (uint64_t)U) ? (uint32_t)B : U;      // the epoch ends when size reaches B
            // seq = ord[0..m) ++ ids_b[m..m2)
            if (m2 > m) CKC(cudaMemcpyAsync(ord + m, ids_b + m, (size_t)(m2 - m) * 4, cudaMemcpyDeviceToDevice, st));
            const unsigned g2 = grid_for(m2);
            CKC(cudaMemsetAsync(touch, 0xFF, (size_t)B * 4, st));
            if (B * 4 <= kTouchDirectBytes) {
                k_touch_min<<<g2, 256, 0, st>>>(ord, d_hash, B, m2, bucket, touch);
                k_touch_key<<<g2, 256, 0, st>>>(bucket, touch, m2, key);
                need = tmp_bytes;
                CKC(cub::DeviceRadixSort::SortPairs(tmp, need, key, key_s, ord, ord_s, (int)m2, 0, bits_for(m2), st));
                *launches += 3 + (bits_for(m2) + 7) / 8 + 1;
            } else {
                // large table: partition by the top 8 bits of the bucket number first (k64_a / k64_b are free between steps 1 and 3)
                unsigned long long *pn = (unsigned long long *)k64_a, *pn_p = (unsigned long long *)k64_b;
                const int hb = bits_for(B - 1), lb = hb > 8 ? hb - 8 : 0;
                k_bucket_pos<<<g2, 256, 0, st>>>(ord, d_hash, B, m2, bucket, pn);
                need = tmp_bytes;
                CKC(cub::DeviceRadixSort::SortPairs(tmp, need, bucket, key_s, pn, pn_p, (int)m2, lb, hb, st));
                k_touch_min_part<<<g2, 256, 0, st>>>(key_s, pn_p, m2, touch);
                k_touch_key_part<<<g2, 256, 0, st>>>(key_s, pn_p, touch, m2, key, bucket);      // bucket[] now holds the nodes
                need = tmp_bytes;
                CKC(cub::DeviceRadixSort::SortPairs(tmp, need, key, key_s, bucket, ord_s, (int)m2, 0, bits_for(m2), st));
                *launches += 4 + 3 + (bits_for(m2) + 7) / 8 + 1;
            }
            m = m2;
            if (m < U) { k_reverse<<<g2, 256, 0, st>>>(ord_s, ord, m2); *launches += 1; }
        }
    }
    // 3. count descending, ties by iteration position descending: ord_s already is the reversed iteration order, so a
    //    stable descending sort by count finishes it
    CKC(cudaStreamSynchronize(st));      // h_max
    k_gather_u64<<<G, 256, 0, st>>>(d_count, ord_s, k64_a, U);
    need = tmp_bytes;
    CKC(cub::DeviceRadixSort::SortPairsDescending(tmp, need, k64_a, k64_b, ord_s, d_perm, (int)U, 0, bits_for(h_max), st));
    *launches += 2 + (bits_for(h_max) + 7) / 8;
    (void)ubits;
    CKC(cudaStreamSynchronize(st));
done:
    {
        void *scratch[] = { ids_a, ids_b, ord, ord_s, bucket, key, key_s, touch, k64_a, k64_b, d_max, tmp };
        for (void *q : scratch) if (q) cudaFreeAsync(q, st);
    }
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
    uint64_t max_first;           // upper bound of every first-occurrence index seen (0 = unknown: explicit `first` arrays)
    int first_unknown;
    char err[256];
};

extern "C" const char *fxg_collapse_error(const fxg_collapser *c) { return c ? c->err : "no collapser"; }
extern "C" int64_t fxg_collapse_launches(const fxg_collapser *c) { return c ? c->launches : 0; }
extern "C" int32_t fxg_collapse_stride(const fxg_collapser *c) { return c ? c->stride : 0; }

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
    if (!out || max_reads <= 0 || max_reads >= 0xFFFFFFF0ll || stride <= 0 || (stride & 15)) return FXG_ERR_ARG;   // both grow on demand
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

// Room for `rows` rows of `stride` bytes: the row store is re-allocated (and re-strided), the table rebuilt at twice the
// size when it would get more than half full.  A streaming caller does not know the number of reads, or the longest one,
// when it creates the table.
static int collapse_reserve(fxg_collapser *c, int64_t rows, int32_t stride)
{
    if (rows <= c->cap && stride <= c->stride) return FXG_OK;
    if (rows >= 0xFFFFFFF0ll) { snprintf(c->err, sizeof(c->err), "collapser: more than 2^32 rows"); return FXG_ERR_UNSUPPORTED; }
    int64_t ncap = c->cap;
    while (ncap < rows) ncap = ncap * 2 > 0xFFFFFFEFll ? 0xFFFFFFEFll : ncap * 2;
    const int32_t nstride = stride > c->stride ? stride : c->stride;
    if (ncap != c->cap || nstride != c->stride) {
        uint8_t *nk = NULL; int32_t *nl = NULL; uint64_t *nh = NULL;
        CKO(c, cudaMalloc(&nk, (size_t)ncap * nstride)); CKO(c, cudaMalloc(&nl, (size_t)ncap * 4)); CKO(c, cudaMalloc(&nh, (size_t)ncap * 8));
        if (c->rows > 0) {
            if (nstride == c->stride) CKO(c, cudaMemcpyAsync(nk, c->keys, (size_t)c->rows * nstride, cudaMemcpyDeviceToDevice, c->st));
            else { k_copy_rows<<<grid_for((uint64_t)c->rows * (c->stride >> 4)), 256, 0, c->st>>>(c->keys, c->stride, nk, nstride, c->rows); c->launches++; }
            CKO(c, cudaMemcpyAsync(nl, c->len, (size_t)c->rows * 4, cudaMemcpyDeviceToDevice, c->st));
            CKO(c, cudaMemcpyAsync(nh, c->hash, (size_t)c->rows * 8, cudaMemcpyDeviceToDevice, c->st));
        }
        CKO(c, cudaStreamSynchronize(c->st));
        cudaFree(c->keys); cudaFree(c->len); cudaFree(c->hash);
        c->keys = nk; c->len = nl; c->hash = nh; c->cap = ncap; c->stride = nstride;
    }
    uint64_t ns = c->nslots;
    while (ns < (uint64_t)ncap * 2) ns <<= 1;
    if (ns != c->nslots) {
        unsigned long long *s2 = NULL, *c2 = NULL, *f2 = NULL;
        CKO(c, cudaMalloc(&s2, ns * 8)); CKO(c, cudaMalloc(&c2, ns * 8)); CKO(c, cudaMalloc(&f2, ns * 8));
        CKO(c, cudaMemsetAsync(s2, 0, ns * 8, c->st)); CKO(c, cudaMemsetAsync(c2, 0, ns * 8, c->st)); CKO(c, cudaMemsetAsync(f2, 0xFF, ns * 8, c->st));
        k_rehash<<<grid_for(c->nslots), 256, 0, c->st>>>(c->slots, c->count, c->firsts, c->hash, (int64_t)c->nslots, s2, ns - 1, c2, f2);
        c->launches++;
        CKO(c, cudaStreamSynchronize(c->st));
        cudaFree(c->slots); cudaFree(c->count); cudaFree(c->firsts);
        c->slots = s2; c->count = c2; c->firsts = f2; c->nslots = ns;
    }
    return FXG_OK;
}

extern "C" int fxg_collapse_reserve(fxg_collapser *c, int64_t rows, int32_t stride)
{
    if (!c || stride < 0 || (stride & 15)) return FXG_ERR_ARG;
    CKO(c, cudaSetDevice(c->device));
    return collapse_reserve(c, rows > c->cap ? rows : c->cap, stride > c->stride ? stride : c->stride);
}

// Append rows (device or host memory — cudaMemcpyDefault) and insert them.  weight/first: see fxg.h.
extern "C" int fxg_collapse_add(fxg_collapser *c, const fxg_batch *b, const int32_t *weight, const int64_t *first, int64_t index_base)
{
    if (!c || !b || !b->seq || b->n < 0 || b->stride <= 0 || (b->stride & 15)) return FXG_ERR_ARG;
    if (b->n == 0) return FXG_OK;
    CKO(c, cudaSetDevice(c->device));
    { int rc = collapse_reserve(c, c->rows + b->n, b->stride); if (rc) return rc; }
    if (first || index_base < 0) c->first_unknown = 1;
    else if ((uint64_t)(index_base + b->n) > c->max_first) c->max_first = (uint64_t)(index_base + b->n);
    const int64_t row0 = c->rows;
    const size_t S = (size_t)c->stride;
    if (b->stride != c->stride) {
        // a batch of shorter rows: staged on the device, then copied row by row into the table's stride
        uint8_t *tmp = NULL; int32_t *d_w = NULL; int64_t *d_f = NULL;
        CKO(c, cudaMalloc(&tmp, (size_t)b->n * b->stride));
        CKO(c, cudaMemcpyAsync(tmp, b->seq, (size_t)b->n * b->stride, cudaMemcpyDefault, c->st));
        k_copy_rows<<<grid_for((uint64_t)b->n * (b->stride >> 4)), 256, 0, c->st>>>(tmp, b->stride, c->keys + (size_t)row0 * S, c->stride, b->n);
        if (weight) { CKO(c, cudaMalloc(&d_w, (size_t)b->n * 4)); CKO(c, cudaMemcpyAsync(d_w, weight, (size_t)b->n * 4, cudaMemcpyDefault, c->st)); }
        if (first) { CKO(c, cudaMalloc(&d_f, (size_t)b->n * 8)); CKO(c, cudaMemcpyAsync(d_f, first, (size_t)b->n * 8, cudaMemcpyDefault, c->st)); }
        if (b->len) CKO(c, cudaMemcpyAsync(c->len + row0, b->len, (size_t)b->n * 4, cudaMemcpyDefault, c->st));
        else k_fill_len<<<grid_for((uint64_t)b->n), 256, 0, c->st>>>(c->len + row0, b->uniform_len, b->n);
        DedupParams p;
        p.keys = c->keys; p.len = c->len; p.meta = NULL; p.stride = c->stride; p.row0 = row0; p.n = b->n;
        p.weight = d_w; p.first = d_f; p.index_base = index_base;
        p.hash = c->hash; p.slots = c->slots; p.mask = c->nslots - 1; p.count = c->count; p.firsts = c->firsts; p.counters = c->d_counters;
        k_hash_dedup<<<grid_for((uint64_t)b->n), 256, 0, c->st>>>(p);
        CKO(c, cudaGetLastError());
        c->launches += 3;
        CKO(c, cudaStreamSynchronize(c->st));
        cudaFree(tmp); cudaFree(d_w); cudaFree(d_f);
        c->rows += b->n;
        return FXG_OK;
    }
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
        p.keys = c->keys; p.len = c->len; p.meta = NULL; p.stride = c->stride; p.row0 = row0 + r; p.n = nr;
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

extern "C" int fxg_collapse_add_next(fxg_collapser *c, const fxg_batch *b)
{
    return c ? fxg_collapse_add(c, b, NULL, NULL, c->rows) : FXG_ERR_ARG;
}

// add_next() that also says whether THIS batch held a read the reader would reject (row index inside the batch, -1 = none):
// the GPU text path hands such a chunk back to the host parser, which words the reference's message
// first_base: first-occurrence index of the batch's row 0 (< 0: the number of rows added so far)
extern "C" int fxg_collapse_add_checked(fxg_collapser *c, const fxg_batch *b, const int32_t *weight, int64_t first_base, int64_t *first_bad_row)
{
    if (!c || !first_bad_row) return FXG_ERR_ARG;
    const int64_t base = first_base >= 0 ? first_base : c->rows;
    unsigned long long before = 0, after = 0;
    CKO(c, cudaSetDevice(c->device));
    CKO(c, cudaMemcpy(&before, c->d_counters + CNT_FIRST_BAD, 8, cudaMemcpyDeviceToHost));
    int rc = fxg_collapse_add(c, b, weight, NULL, base);
    if (rc) return rc;
    CKO(c, cudaMemcpy(&after, c->d_counters + CNT_FIRST_BAD, 8, cudaMemcpyDeviceToHost));
    *first_bad_row = (after != before && after != ~0ull && (int64_t)after >= base) ? (int64_t)after - base : -1;
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
        int rc = fxg_order_impl(c->u_hash, c->u_first, c->u_count, (uint32_t)U, c->perm, c->first_unknown ? 0 : c->max_first, c->st, c->err, sizeof(c->err), &c->launches);
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
        k_gather_rows<<<grid_for((uint64_t)U * (c->stride >> 4)), 256, 0, c->st>>>(c->keys, c->len, NULL, c->u_rep, c->perm, c->stride, U, rows, lens);
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
    int rc = fxg_order_impl(hash_dev, first_dev, count_dev, (uint32_t)n_unique, perm_dev, 0, st, err, sizeof(err), &launches);
    if (rc) fprintf(stderr, "fxg_collapse_order_dev: %s\n", err);
    cudaStreamDestroy(st);
    return rc;
}

namespace fxg {
cudaError_t launch_hash_dedup(const DedupParams &p, cudaStream_t st)
{
    k_hash_dedup<<<grid_for((uint64_t)p.n), 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_compact(const unsigned long long *slots, const unsigned long long *count, const unsigned long long *firsts,
                           const uint64_t *hash, int64_t nslots, unsigned long long *n_out, uint32_t *u_rep, uint64_t *u_hash,
                           uint64_t *u_first, uint64_t *u_count, cudaStream_t st)
{
    k_compact<<<grid_for((uint64_t)nslots), 256, 0, st>>>(slots, count, firsts, hash, nslots, n_out, u_rep, u_hash, u_first, u_count);
    return cudaGetLastError();
}
cudaError_t launch_gather_rows(const uint8_t *keys, const int32_t *len, const RowMeta *meta, const uint32_t *rep, const uint32_t *perm,
                               int stride, uint32_t n, uint8_t *out_rows, int32_t *out_len, cudaStream_t st)
{
    k_gather_rows<<<grid_for((uint64_t)n * (stride >> 4)), 256, 0, st>>>(keys, len, meta, rep, perm, stride, n, out_rows, out_len);
    return cudaGetLastError();
}
cudaError_t launch_hash(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, uint64_t *out, cudaStream_t st)
{
    k_hash<<<grid_for((uint64_t)n), 256, 0, st>>>(seq, len, uniform_len, stride, n, out);
    return cudaGetLastError();
}
}
