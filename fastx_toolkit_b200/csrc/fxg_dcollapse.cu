// fxg_dcollapse.cu — fastx_collapser across GPUs (BASELINE config (e)): the global count map of
// src/fastx_collapser/fastx_collapser.cpp:112-114 partitioned by owner = std::hash(seq) mod G.
//
//   K-ROUTE    per read: _Hash_bytes, owner = hash mod G, a slot in the send slab of that owner (warp-aggregated cursors)
//   exchange   key rows (stride bytes) + {first index, weight, len} (16 bytes) per read, grouped ncclSend/ncclRecv over NVLink
//   K-DEDUP    on the owner: the same exact-compare table as the one-GPU collapser, fed by the received rows
//   gather     (hash, first, count) of every owner's uniques -> the root GPU (24 bytes per unique; key rows stay where they are)
//   K-ORDER    on the root: the reference's output order (fastx_collapser.cpp:116-122 + libstdc++ iteration order)
//
// The object drives the `nlocal` GPUs of its communicator: all GPUs of the box in one process (the drop-in tool), or one
// GPU per process (torchrun).  Every phase is enqueued per local GPU on the communicator's streams; the host only reads
// the two small count matrices it needs to size the receive buffers.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fxg.h"
#include "fxg_collapse.cuh"
#include "fxg_comm.h"
#include "fxg_kernels.cuh"

namespace fxg {

constexpr int DC_MAX_RANKS = 64;

__device__ __forceinline__ uint64_t dc_shift_mix(uint64_t v) { return v ^ (v >> 47); }
// libstdc++ _Hash_bytes (same function as fxg_collapse.cu's hash_row; the owner must be hash mod G of THAT value so that a
// key always lands on the same GPU)
__device__ __forceinline__ uint64_t dc_hash_row(const uint8_t *row, int len)
{
    const uint64_t mul = (0xc6a4a793ull << 32) + 0x5bd1e995ull;
    uint64_t h = 0xc70f6907ull ^ ((uint64_t)len * mul);
    const int n8 = len >> 3;
    const uint2 *p = reinterpret_cast<const uint2 *>(row);
    for (int k = 0; k < n8; k++) {
        const uint2 v = __ldg(p + k);
        uint64_t d = ((uint64_t)v.y << 32) | v.x;
        d = dc_shift_mix(d * mul) * mul;
        h ^= d;
        h *= mul;
    }
    const int rem = len & 7;
    if (rem) {
        const uint2 v = __ldg(p + n8);
        uint64_t d = ((uint64_t)v.y << 32) | v.x;
        d &= (1ull << (8 * rem)) - 1ull;
        h ^= d;
        h *= mul;
    }
    h = dc_shift_mix(h) * mul;
    return dc_shift_mix(h);
}

// pass 1: owner of every read + reads per owner
__global__ void __launch_bounds__(256) k_route_count(const uint8_t *seq, const int32_t *len, int uniform_len, int stride, int64_t n, int G,
                                                     uint8_t *owner, unsigned long long *counts)
{
    __shared__ unsigned int s_cnt[DC_MAX_RANKS];
    if (threadIdx.x < DC_MAX_RANKS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = t0 + threadIdx.x;
        int o = -1;
        if (t < n) {
            const int L = len ? __ldg(len + t) : uniform_len;
            o = (L > 0 && L <= stride) ? (int)(dc_hash_row(seq + (size_t)t * stride, L) % (uint64_t)G) : 0;   // bad rows: the owner's table rejects them
            owner[t] = (uint8_t)o;
        }
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, o);
        if (o >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&s_cnt[o], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < G && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// pass 2: a position in the owner's segment of the send slab for every read, and its meta record
__global__ void __launch_bounds__(256) k_route_place(const uint8_t *owner, const int32_t *len, int uniform_len, const int32_t *weight,
                                                     const int64_t *first, int64_t n, int64_t index_base, unsigned long long *cursor, uint32_t *pos,
                                                     RowMeta *meta)
{
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = t0 + threadIdx.x;
        const int o = (t < n) ? (int)owner[t] : -1;
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, o);
        if (o < 0) continue;
        const unsigned lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&cursor[o], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint32_t p = (uint32_t)(base + __popc(peers & ((1u << lane) - 1u)));
        pos[t] = p;
        RowMeta m;
        m.first = first ? first[t] : index_base + t;
        m.weight = weight ? (uint32_t)__ldg(weight + t) : 1u;
        m.len = len ? __ldg(len + t) : uniform_len;
        meta[p] = m;
    }
}

// pass 3: the key rows, one 16-byte chunk per thread
__global__ void __launch_bounds__(256) k_route_rows(const uint8_t *seq, const uint32_t *pos, int stride, int64_t n, uint8_t *out)
{
    const int chunks = stride >> 4;
    const uint64_t total = (uint64_t)n * chunks;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = t / chunks;
        const int c = (int)(t - i * chunks);
        reinterpret_cast<uint4 *>(out + (size_t)pos[i] * stride)[c] = __ldg(reinterpret_cast<const uint4 *>(seq + (size_t)i * stride) + c);
    }
}

}  // namespace fxg

using namespace fxg;

namespace {

struct Buf {            // grow-only device buffer
    void *p; size_t cap;
};

struct Local {
    int device, rank;
    Buf owner, pos, srows, smeta, rrows, rmeta, hash, slots, count, firsts, u_rep, u_hash, u_first, u_count, small;
    unsigned long long *h_small;      // pinned: count matrices come back here
    int64_t m;                        // rows received
    uint64_t nslots;
    int64_t U;                        // uniques owned
    cudaEvent_t ev[6];
};

}  // namespace

struct fxg_dcollapse {
    fxg_comm *comm;
    int G, nlocal;
    int32_t stride;
    Local *loc;
    int root, root_local;
    Buf g_hash, g_first, g_count, g_perm;
    int64_t U_total;
    int64_t *u_off;                   // [G+1]
    int64_t launches;
    fxg_dcollapse_report rep;
    char err[256];
};

extern "C" const char *fxg_dcollapse_error(const fxg_dcollapse *d) { return d ? d->err : "no collapser"; }
extern "C" int64_t fxg_dcollapse_launches(const fxg_dcollapse *d) { return d ? d->launches : 0; }

#define CKD(d, call)                                                                               \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((d)->err, sizeof((d)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return FXG_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)
#define CKR(d, call)                                                                               \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != FXG_OK) {                                                                        \
            snprintf((d)->err, sizeof((d)->err), "%s", fxg_comm_error((d)->comm));                  \
            return r_;                                                                             \
        }                                                                                          \
    } while (0)

static inline unsigned dgrid(uint64_t n) { uint64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16; if (b < 1) b = 1; return (unsigned)b; }

static cudaError_t ensure(Buf *b, size_t bytes)
{
    if (bytes <= b->cap) return cudaSuccess;
    if (b->p) cudaFree(b->p);
    b->p = NULL; b->cap = 0;
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e == cudaSuccess) b->cap = want;
    return e;
}
static void drop(Buf *b) { if (b->p) cudaFree(b->p); b->p = NULL; b->cap = 0; }

extern "C" void fxg_dcollapse_free(fxg_dcollapse *d)
{
    if (!d) return;
    for (int i = 0; i < d->nlocal && d->loc; i++) {
        Local *l = &d->loc[i];
        cudaSetDevice(l->device);
        Buf *bs[] = { &l->owner, &l->pos, &l->srows, &l->smeta, &l->rrows, &l->rmeta, &l->hash, &l->slots, &l->count, &l->firsts,
                      &l->u_rep, &l->u_hash, &l->u_first, &l->u_count, &l->small };
        for (Buf *b : bs) drop(b);
        if (l->h_small) cudaFreeHost(l->h_small);
        for (cudaEvent_t e : l->ev) if (e) cudaEventDestroy(e);
        if (i == d->root_local) { drop(&d->g_hash); drop(&d->g_first); drop(&d->g_count); drop(&d->g_perm); }
    }
    free(d->loc); free(d->u_off);
    free(d);
}

extern "C" int fxg_dcollapse_new(fxg_comm *comm, int32_t stride, fxg_dcollapse **out)
{
    if (!out || !comm || stride <= 0 || (stride & 15) || comm->nranks > DC_MAX_RANKS) return FXG_ERR_ARG;
    *out = NULL;
    fxg_dcollapse *d = (fxg_dcollapse *)calloc(1, sizeof(fxg_dcollapse));
    if (!d) return FXG_ERR_NOMEM;
    d->comm = comm; d->G = comm->nranks; d->nlocal = comm->nlocal; d->stride = stride; d->root = 0; d->root_local = -1;
    d->loc = (Local *)calloc(d->nlocal, sizeof(Local));
    d->u_off = (int64_t *)calloc(d->G + 1, sizeof(int64_t));
    if (!d->loc || !d->u_off) { fxg_dcollapse_free(d); return FXG_ERR_NOMEM; }
    for (int i = 0; i < d->nlocal; i++) {
        Local *l = &d->loc[i];
        l->device = comm->devices[i]; l->rank = comm->ranks[i];
        if (cudaSetDevice(l->device) != cudaSuccess ||
            cudaMallocHost(&l->h_small, sizeof(unsigned long long) * (size_t)(d->G + 2) * d->G) != cudaSuccess) { fxg_dcollapse_free(d); return FXG_ERR_CUDA; }
        for (cudaEvent_t &e : l->ev) if (cudaEventCreate(&e) != cudaSuccess) { fxg_dcollapse_free(d); return FXG_ERR_CUDA; }
    }
    *out = d;
    return FXG_OK;
}

// One collapse of the reads resident on the local GPUs: batches[i] (DEVICE slabs on local GPU i, seq only; len == NULL for a
// uniform length) holds reads whose global indices are index_base[i] + row, or first_dev[i][row] when first_dev (and
// first_dev[i]) is given; weight_dev[i] (or weight_dev == NULL) as in fxg_collapse_add.  On return every owner holds its uniques, and the root's process holds the output order.
extern "C" int fxg_dcollapse_run(fxg_dcollapse *d, const fxg_batch *batches, const int64_t *index_base, const int32_t *const *weight_dev,
                                 const int64_t *const *first_dev, int root, fxg_dcollapse_report *rep)
{
    if (!d || !batches || !index_base || root < 0 || root >= d->G) return FXG_ERR_ARG;
    fxg_comm *c = d->comm;
    const int G = d->G, NL = d->nlocal;
    const size_t S = (size_t)d->stride;
    const int64_t bytes0 = c->bytes_sent;
    for (int i = 0; i < NL; i++) {
        if (!batches[i].seq || batches[i].stride != d->stride || batches[i].n < 0 || batches[i].n >= 0xFFFFFFF0ll) {
            snprintf(d->err, sizeof d->err, "batch %d: NULL slab, wrong stride or too many reads", i);
            return FXG_ERR_ARG;
        }
    }
    if (d->root_local >= 0 && c->ranks[d->root_local] != root) {      // the root moved: its buffers live on another GPU
        cudaSetDevice(d->loc[d->root_local].device);
        drop(&d->g_hash); drop(&d->g_first); drop(&d->g_count); drop(&d->g_perm);
    }
    d->root = root; d->root_local = -1;
    for (int i = 0; i < NL; i++) if (c->ranks[i] == root) d->root_local = i;

    // ---- K-ROUTE pass 1: owners and per-owner counts -------------------------------------------------------------
    for (int i = 0; i < NL; i++) {
        Local *l = &d->loc[i];
        const fxg_batch *b = &batches[i];
        cudaStream_t st = c->streams[i];
        CKD(d, cudaSetDevice(l->device));
        CKD(d, cudaEventRecord(l->ev[0], st));
        CKD(d, ensure(&l->small, sizeof(unsigned long long) * (size_t)(G + 2) * (G + 1)));
        CKD(d, ensure(&l->owner, (size_t)b->n + 16));
        CKD(d, cudaMemsetAsync(l->small.p, 0, sizeof(unsigned long long) * (size_t)(G + 2), st));
        if (b->n > 0) {
            k_route_count<<<dgrid((uint64_t)b->n), 256, 0, st>>>(b->seq, b->len, b->uniform_len, b->stride, b->n, G, (uint8_t *)l->owner.p,
                                                                 (unsigned long long *)l->small.p);
            CKD(d, cudaGetLastError());
            d->launches++;
        }
    }
    // every rank learns the whole count matrix: row r = reads rank r sends to each owner
    {
        const void *snd[DC_MAX_RANKS]; void *rcv[DC_MAX_RANKS];
        for (int i = 0; i < NL; i++) { snd[i] = d->loc[i].small.p; rcv[i] = (unsigned long long *)d->loc[i].small.p + (G + 2); }
        CKR(d, fxg_comm_allgather(c, snd, rcv, sizeof(unsigned long long) * (size_t)G));
        for (int i = 0; i < NL; i++) {
            CKD(d, cudaSetDevice(d->loc[i].device));
            CKD(d, cudaMemcpyAsync(d->loc[i].h_small, rcv[i], sizeof(unsigned long long) * (size_t)G * G, cudaMemcpyDeviceToHost, c->streams[i]));
        }
        CKR(d, fxg_comm_sync(c));
    }
    const unsigned long long *M = d->loc[0].h_small;      // M[s*G + o]: identical on every local GPU
    int64_t send_off[DC_MAX_RANKS * DC_MAX_RANKS], send_cnt[DC_MAX_RANKS * DC_MAX_RANKS], recv_off[DC_MAX_RANKS * DC_MAX_RANKS],
        recv_cnt[DC_MAX_RANKS * DC_MAX_RANKS];
    if ((size_t)NL * G > sizeof(send_off) / sizeof(send_off[0])) return FXG_ERR_ARG;
    // ---- K-ROUTE passes 2 and 3: fill the send slabs -----------------------------------------------------------------
    for (int i = 0; i < NL; i++) {
        Local *l = &d->loc[i];
        const fxg_batch *b = &batches[i];
        cudaStream_t st = c->streams[i];
        const int me = l->rank;
        int64_t so = 0, ro = 0;
        for (int p = 0; p < G; p++) {
            send_cnt[i * G + p] = (int64_t)M[(size_t)me * G + p]; send_off[i * G + p] = so; so += send_cnt[i * G + p];
            recv_cnt[i * G + p] = (int64_t)M[(size_t)p * G + me]; recv_off[i * G + p] = ro; ro += recv_cnt[i * G + p];
        }
        if (so != b->n) { snprintf(d->err, sizeof d->err, "route: %lld rows counted, %lld expected", (long long)so, (long long)b->n); return FXG_ERR_CUDA; }
        if (ro >= 0xFFFFFFF0ll) { snprintf(d->err, sizeof d->err, "owner %d would receive %lld rows (limit 2^32)", me, (long long)ro); return FXG_ERR_UNSUPPORTED; }
        l->m = ro;
        CKD(d, cudaSetDevice(l->device));
        CKD(d, ensure(&l->pos, (size_t)b->n * 4 + 16));
        CKD(d, ensure(&l->srows, (size_t)b->n * S + 16));
        CKD(d, ensure(&l->smeta, (size_t)b->n * sizeof(RowMeta) + 16));
        CKD(d, ensure(&l->rrows, (size_t)l->m * S + 16));
        CKD(d, ensure(&l->rmeta, (size_t)l->m * sizeof(RowMeta) + 16));
        unsigned long long *cursor = (unsigned long long *)l->small.p;      // reuse the counts words as cursors = send offsets
        unsigned long long h_cur[DC_MAX_RANKS];
        for (int p = 0; p < G; p++) h_cur[p] = (unsigned long long)send_off[i * G + p];
        CKD(d, cudaMemcpyAsync(cursor, h_cur, sizeof(unsigned long long) * (size_t)G, cudaMemcpyHostToDevice, st));
        if (b->n > 0) {
            k_route_place<<<dgrid((uint64_t)b->n), 256, 0, st>>>((const uint8_t *)l->owner.p, b->len, b->uniform_len,
                                                                 weight_dev ? weight_dev[i] : NULL, first_dev ? first_dev[i] : NULL, b->n, index_base[i], cursor,
                                                                 (uint32_t *)l->pos.p, (RowMeta *)l->smeta.p);
            k_route_rows<<<dgrid((uint64_t)b->n * (S >> 4)), 256, 0, st>>>(b->seq, (const uint32_t *)l->pos.p, d->stride, b->n, (uint8_t *)l->srows.p);
            CKD(d, cudaGetLastError());
            d->launches += 2;
        }
        CKD(d, cudaEventRecord(l->ev[1], st));
    }
    // ---- the exchange ------------------------------------------------------------------------------------------------
    {
        const void *snd[DC_MAX_RANKS]; void *rcv[DC_MAX_RANKS];
        for (int i = 0; i < NL; i++) { snd[i] = d->loc[i].srows.p; rcv[i] = d->loc[i].rrows.p; }
        CKR(d, fxg_comm_alltoallv(c, snd, send_off, send_cnt, rcv, recv_off, recv_cnt, S));
        for (int i = 0; i < NL; i++) { snd[i] = d->loc[i].smeta.p; rcv[i] = d->loc[i].rmeta.p; }
        CKR(d, fxg_comm_alltoallv(c, snd, send_off, send_cnt, rcv, recv_off, recv_cnt, sizeof(RowMeta)));
    }
    // ---- K-DEDUP on the owners -----------------------------------------------------------------------------------------
    for (int i = 0; i < NL; i++) {
        Local *l = &d->loc[i];
        cudaStream_t st = c->streams[i];
        CKD(d, cudaSetDevice(l->device));
        CKD(d, cudaEventRecord(l->ev[2], st));
        uint64_t ns = 1024;
        while (ns < (uint64_t)l->m * 2) ns <<= 1;
        l->nslots = ns;
        const int64_t ucap = l->m > 0 ? l->m : 1;
        CKD(d, ensure(&l->slots, ns * 8)); CKD(d, ensure(&l->count, ns * 8)); CKD(d, ensure(&l->firsts, ns * 8));
        CKD(d, ensure(&l->hash, (size_t)ucap * 8));
        CKD(d, ensure(&l->u_rep, (size_t)ucap * 4)); CKD(d, ensure(&l->u_hash, (size_t)ucap * 8));
        CKD(d, ensure(&l->u_first, (size_t)ucap * 8)); CKD(d, ensure(&l->u_count, (size_t)ucap * 8));
        CKD(d, cudaMemsetAsync(l->slots.p, 0, ns * 8, st)); CKD(d, cudaMemsetAsync(l->count.p, 0, ns * 8, st));
        CKD(d, cudaMemsetAsync(l->firsts.p, 0xFF, ns * 8, st));
        unsigned long long *cnt = (unsigned long long *)l->small.p;      // [0] = uniques, [1] = first bad read (CNT_FIRST_BAD), [2] = bound of the first indices put in here
        const unsigned long long init[3] = { 0ull, ~0ull, (first_dev && first_dev[i]) ? ~0ull : (unsigned long long)(index_base[i] + batches[i].n) };
        CKD(d, cudaMemcpyAsync(cnt, init, sizeof init, cudaMemcpyHostToDevice, st));
        if (l->m > 0) {
            DedupParams p;
            memset(&p, 0, sizeof p);
            p.keys = (const uint8_t *)l->rrows.p; p.meta = (const RowMeta *)l->rmeta.p; p.stride = d->stride; p.row0 = 0; p.n = l->m;
            p.hash = (uint64_t *)l->hash.p; p.slots = (unsigned long long *)l->slots.p; p.mask = ns - 1;
            p.count = (unsigned long long *)l->count.p; p.firsts = (unsigned long long *)l->firsts.p; p.counters = cnt;
            CKD(d, launch_hash_dedup(p, st));
            CKD(d, launch_compact(p.slots, p.count, p.firsts, p.hash, (int64_t)ns, cnt + CNT_OUT, (uint32_t *)l->u_rep.p, (uint64_t *)l->u_hash.p,
                                  (uint64_t *)l->u_first.p, (uint64_t *)l->u_count.p, st));
            d->launches += 2;
        }
        CKD(d, cudaEventRecord(l->ev[3], st));
    }
    // every rank learns {uniques, first bad read} of every owner
    {
        const void *snd[DC_MAX_RANKS]; void *rcv[DC_MAX_RANKS];
        for (int i = 0; i < NL; i++) { snd[i] = d->loc[i].small.p; rcv[i] = (unsigned long long *)d->loc[i].small.p + (G + 2); }
        CKR(d, fxg_comm_allgather(c, snd, rcv, sizeof(unsigned long long) * 3));
        for (int i = 0; i < NL; i++) {
            CKD(d, cudaSetDevice(d->loc[i].device));
            CKD(d, cudaMemcpyAsync(d->loc[i].h_small, rcv[i], sizeof(unsigned long long) * 3 * (size_t)G, cudaMemcpyDeviceToHost, c->streams[i]));
        }
        CKR(d, fxg_comm_sync(c));
    }
    int64_t ucnt[DC_MAX_RANKS];
    unsigned long long first_bad = ~0ull, max_first = 0;
    d->u_off[0] = 0;
    for (int s = 0; s < G; s++) {
        ucnt[s] = (int64_t)d->loc[0].h_small[3 * s];
        if (d->loc[0].h_small[3 * s + 1] < first_bad) first_bad = d->loc[0].h_small[3 * s + 1];
        if (d->loc[0].h_small[3 * s + 2] > max_first) max_first = d->loc[0].h_small[3 * s + 2];
        d->u_off[s + 1] = d->u_off[s] + ucnt[s];
    }
    d->U_total = d->u_off[G];
    for (int i = 0; i < NL; i++) d->loc[i].U = ucnt[d->loc[i].rank];
    if (d->U_total >= 0xFFFFFFF0ll) { snprintf(d->err, sizeof d->err, "%lld unique sequences (limit 2^32)", (long long)d->U_total); return FXG_ERR_UNSUPPORTED; }
    // ---- gather the triples on the root, K-ORDER there -------------------------------------------------------------------
    if (d->root_local >= 0) {
        CKD(d, cudaSetDevice(d->loc[d->root_local].device));
        const size_t u = (size_t)(d->U_total > 0 ? d->U_total : 1);
        CKD(d, ensure(&d->g_hash, u * 8)); CKD(d, ensure(&d->g_first, u * 8)); CKD(d, ensure(&d->g_count, u * 8)); CKD(d, ensure(&d->g_perm, u * 4));
    }
    {
        const void *snd[DC_MAX_RANKS];
        for (int i = 0; i < NL; i++) snd[i] = d->loc[i].u_hash.p;
        CKR(d, fxg_comm_gatherv(c, snd, ucnt, d->u_off, d->g_hash.p, root, 8));
        for (int i = 0; i < NL; i++) snd[i] = d->loc[i].u_first.p;
        CKR(d, fxg_comm_gatherv(c, snd, ucnt, d->u_off, d->g_first.p, root, 8));
        for (int i = 0; i < NL; i++) snd[i] = d->loc[i].u_count.p;
        CKR(d, fxg_comm_gatherv(c, snd, ucnt, d->u_off, d->g_count.p, root, 8));
    }
    for (int i = 0; i < NL; i++) { CKD(d, cudaSetDevice(d->loc[i].device)); CKD(d, cudaEventRecord(d->loc[i].ev[4], c->streams[i])); }
    if (d->root_local >= 0 && d->U_total > 0) {
        Local *l = &d->loc[d->root_local];
        CKD(d, cudaSetDevice(l->device));
        // every first index is below the largest index_base + n of the job: that bound sizes the radix passes of the first sort
        int rc = fxg_order_impl((const uint64_t *)d->g_hash.p, (const uint64_t *)d->g_first.p, (const uint64_t *)d->g_count.p, (uint32_t)d->U_total,
                                (uint32_t *)d->g_perm.p, max_first == ~0ull ? 0 : (uint64_t)max_first, c->streams[d->root_local], d->err, sizeof d->err, &d->launches);
        if (rc) return rc;
    }
    for (int i = 0; i < NL; i++) { CKD(d, cudaSetDevice(d->loc[i].device)); CKD(d, cudaEventRecord(d->loc[i].ev[5], c->streams[i])); }
    CKR(d, fxg_comm_sync(c));

    fxg_dcollapse_report *r = &d->rep;
    memset(r, 0, sizeof *r);
    r->n_unique = d->U_total;
    r->first_bad_read = (first_bad == ~0ull) ? -1 : (int64_t)first_bad;
    r->n_reads_local = 0; r->n_unique_local = 0; r->rows_received = 0;
    for (int i = 0; i < NL; i++) { r->n_reads_local += batches[i].n; r->n_unique_local += d->loc[i].U; r->rows_received += d->loc[i].m; }
    r->bytes_sent = c->bytes_sent - bytes0;
    {
        Local *l = &d->loc[d->root_local >= 0 ? d->root_local : 0];
        CKD(d, cudaSetDevice(l->device));
        for (int k = 0; k < 5; k++) { float ms = 0; cudaEventElapsedTime(&ms, l->ev[k], l->ev[k + 1]); r->ms[k] = ms; }
    }
    if (rep) *rep = *r;
    return FXG_OK;
}

// The uniques owned by local GPU i, in the owner's table order (the order fxg_dcollapse_fetch_order's indices refer to).
// Destinations may be device or host memory; any may be NULL.
extern "C" int fxg_dcollapse_fetch_local(fxg_dcollapse *d, int i, uint8_t *out_seq, int32_t *out_len, uint64_t *out_count, int64_t *out_first,
                                         uint64_t *out_hash)
{
    if (!d || i < 0 || i >= d->nlocal) return FXG_ERR_ARG;
    Local *l = &d->loc[i];
    if (l->U == 0) return FXG_OK;
    cudaStream_t st = d->comm->streams[i];
    CKD(d, cudaSetDevice(l->device));
    if (out_count) CKD(d, cudaMemcpyAsync(out_count, l->u_count.p, (size_t)l->U * 8, cudaMemcpyDefault, st));
    if (out_first) CKD(d, cudaMemcpyAsync(out_first, l->u_first.p, (size_t)l->U * 8, cudaMemcpyDefault, st));
    if (out_hash) CKD(d, cudaMemcpyAsync(out_hash, l->u_hash.p, (size_t)l->U * 8, cudaMemcpyDefault, st));
    if (out_seq || out_len) {
        uint8_t *rows = NULL; int32_t *lens = NULL;
        CKD(d, cudaMallocAsync(&rows, (size_t)l->U * d->stride, st)); CKD(d, cudaMallocAsync(&lens, (size_t)l->U * 4, st));
        CKD(d, launch_gather_rows((const uint8_t *)l->rrows.p, NULL, (const RowMeta *)l->rmeta.p, (const uint32_t *)l->u_rep.p, NULL, d->stride,
                                  (uint32_t)l->U, rows, lens, st));
        d->launches++;
        if (out_seq) CKD(d, cudaMemcpyAsync(out_seq, rows, (size_t)l->U * d->stride, cudaMemcpyDefault, st));
        if (out_len) CKD(d, cudaMemcpyAsync(out_len, lens, (size_t)l->U * 4, cudaMemcpyDefault, st));
        cudaFreeAsync(rows, st); cudaFreeAsync(lens, st);
    }
    CKD(d, cudaStreamSynchronize(st));
    return FXG_OK;
}

// On the process that holds the root GPU: the output order.  The unique printed at rank k is entry perm_index[k] of owner
// perm_owner[k]'s table (fxg_dcollapse_fetch_local); ordered_first / ordered_count (may be NULL) are its first-occurrence
// index and its count.  HOST destinations of n_unique entries.
extern "C" int fxg_dcollapse_fetch_order(fxg_dcollapse *d, int32_t *perm_owner, uint32_t *perm_index, int64_t *ordered_first, uint64_t *ordered_count)
{
    if (!d) return FXG_ERR_ARG;
    if (d->root_local < 0) { snprintf(d->err, sizeof d->err, "this process does not hold the root GPU (rank %d)", d->root); return FXG_ERR_ARG; }
    const int64_t U = d->U_total;
    if (U == 0) return FXG_OK;
    Local *l = &d->loc[d->root_local];
    cudaStream_t st = d->comm->streams[d->root_local];
    CKD(d, cudaSetDevice(l->device));
    uint32_t *perm = (uint32_t *)malloc((size_t)U * 4);
    if (!perm) return FXG_ERR_NOMEM;
    cudaError_t e = cudaMemcpyAsync(perm, d->g_perm.p, (size_t)U * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    uint64_t *tmp = NULL;
    if (e == cudaSuccess && (ordered_first || ordered_count)) {
        tmp = (uint64_t *)malloc((size_t)U * 8);
        if (!tmp) { free(perm); return FXG_ERR_NOMEM; }
        const void *srcs[2] = { d->g_first.p, d->g_count.p };
        uint64_t *dsts[2] = { (uint64_t *)ordered_first, ordered_count };
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            if (!dsts[k]) continue;
            e = cudaMemcpyAsync(tmp, srcs[k], (size_t)U * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) for (int64_t t = 0; t < U; t++) dsts[k][t] = tmp[perm[t]];
        }
    }
    if (e == cudaSuccess && (perm_owner || perm_index)) {
        for (int64_t t = 0; t < U; t++) {
            const int64_t g = perm[t];
            int lo = 0, hi = d->G;                      // owner s with u_off[s] <= g < u_off[s+1]
            while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (d->u_off[mid] <= g) lo = mid; else hi = mid; }
            if (perm_owner) perm_owner[t] = lo;
            if (perm_index) perm_index[t] = (uint32_t)(g - d->u_off[lo]);
        }
    }
    free(perm); free(tmp);
    if (e != cudaSuccess) { snprintf(d->err, sizeof d->err, "fetch_order: %s", cudaGetErrorString(e)); return FXG_ERR_CUDA; }
    return FXG_OK;
}
