// fxg_comm.cu — the collectives of the FASTX hot path, natively in the C ABI:
//   * all-reduce (sum, u64) of the per-GPU fastx_quality_stats histograms          fxg_comm_allreduce_u64
//   * the collapser's owner exchange: grouped ncclSend/ncclRecv, variable counts   fxg_comm_alltoallv
//   * small all-gather (count matrices) and gather-to-root (the (hash, first, count) triples of the uniques)
// A communicator drives the `nlocal` GPUs this PROCESS owns: all GPUs of the box for the drop-in tools
// (fxg_comm_init_all: ncclCommInitAll), or one GPU per process under torchrun / mpirun (fxg_comm_init_rank: the
// 128-byte id made by fxg_comm_unique_id() on one process is handed to the others by whatever launched them).
// NCCL is resolved at run time (dlopen "libnccl.so.2") so that libfxg.so has no load-time dependency on it and shares the
// NCCL instance a host program (e.g. PyTorch) may already have loaded.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <strings.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fxg.h"
#include "fxg_comm.h"

static char g_comm_err[256] = "";
extern "C" const char *fxg_comm_error(const fxg_comm *c) { return c ? c->err : g_comm_err; }

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load(void)
{
    if (g_nccl.lib) return FXG_OK;
    // the drop-in tools write their DATA to stdout and their stderr is part of the drop-in contract: NCCL's debug log must
    // never land on stdout, and a bare version banner (NCCL_DEBUG=VERSION, which this image exports) is not worth a line on stderr
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    { const char *lvl = getenv("NCCL_DEBUG"); if (lvl && !strcasecmp(lvl, "VERSION")) unsetenv("NCCL_DEBUG"); }
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { snprintf(g_comm_err, sizeof g_comm_err, "dlopen(libnccl.so.2): %s", dlerror()); return FXG_ERR_NCCL; }
    NcclApi a;
    memset(&a, 0, sizeof a);
    *(void **)&a.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
    *(void **)&a.CommInitRank = dlsym(lib, "ncclCommInitRank");
    *(void **)&a.CommInitAll = dlsym(lib, "ncclCommInitAll");
    *(void **)&a.CommDestroy = dlsym(lib, "ncclCommDestroy");
    *(void **)&a.AllReduce = dlsym(lib, "ncclAllReduce");
    *(void **)&a.AllGather = dlsym(lib, "ncclAllGather");
    *(void **)&a.Send = dlsym(lib, "ncclSend");
    *(void **)&a.Recv = dlsym(lib, "ncclRecv");
    *(void **)&a.GroupStart = dlsym(lib, "ncclGroupStart");
    *(void **)&a.GroupEnd = dlsym(lib, "ncclGroupEnd");
    *(void **)&a.GetErrorString = dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommInitAll || !a.CommDestroy || !a.AllReduce || !a.AllGather || !a.Send || !a.Recv ||
        !a.GroupStart || !a.GroupEnd || !a.GetErrorString) {
        snprintf(g_comm_err, sizeof g_comm_err, "libnccl.so.2 lacks a required symbol");
        return FXG_ERR_NCCL;
    }
    a.lib = lib;
    g_nccl = a;
    return FXG_OK;
}

// NCCL prints its version banner with printf() when NCCL_DEBUG=VERSION/WARN (this image exports NCCL_DEBUG=VERSION).  The
// drop-in tools' stdout is DATA and their stderr is compared with the reference's, so while communicators are created fd 1
// points at /dev/null (at stderr when the user asked for NCCL's INFO / TRACE log), then it is restored.
struct StdoutGuard {
    int saved;
    StdoutGuard()
    {
        fflush(stdout);
        saved = dup(STDOUT_FILENO);
        const char *lvl = getenv("NCCL_DEBUG");
        const bool chatty = lvl && (!strcasecmp(lvl, "INFO") || !strcasecmp(lvl, "TRACE") || !strcasecmp(lvl, "ABORT"));
        int to = chatty ? -1 : open("/dev/null", O_WRONLY);
        if (saved >= 0) dup2(to >= 0 ? to : STDERR_FILENO, STDOUT_FILENO);
        if (to >= 0) close(to);
    }
    ~StdoutGuard() { fflush(stdout); if (saved >= 0) { dup2(saved, STDOUT_FILENO); close(saved); } }
};

static int nccl_fail(fxg_comm *c, const char *what, ncclResult_t r)
{
    snprintf(c ? c->err : g_comm_err, 256, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return FXG_ERR_NCCL;
}

extern "C" void fxg_comm_free(fxg_comm *c)
{
    if (!c) return;
    for (int i = 0; i < c->nlocal; i++) {
        if (c->comms && c->comms[i] && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->comms[i]);
        if (c->own_streams && c->own_streams[i]) { cudaSetDevice(c->devices[i]); cudaStreamDestroy(c->own_streams[i]); }
    }
    free(c->comms); free(c->streams); free(c->own_streams); free(c->devices); free(c->ranks);
    free(c);
}

static fxg_comm *comm_alloc(int nlocal, int nranks)
{
    fxg_comm *c = (fxg_comm *)calloc(1, sizeof(fxg_comm));
    if (!c) return NULL;
    c->nlocal = nlocal; c->nranks = nranks;
    c->devices = (int *)calloc(nlocal, sizeof(int));
    c->ranks = (int *)calloc(nlocal, sizeof(int));
    c->comms = (void **)calloc(nlocal, sizeof(void *));
    c->streams = (cudaStream_t *)calloc(nlocal, sizeof(cudaStream_t));
    c->own_streams = (cudaStream_t *)calloc(nlocal, sizeof(cudaStream_t));
    if (!c->devices || !c->ranks || !c->comms || !c->streams || !c->own_streams) { fxg_comm_free(c); return NULL; }
    return c;
}

static int comm_make_streams(fxg_comm *c)
{
    for (int i = 0; i < c->nlocal; i++) {
        if (cudaSetDevice(c->devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_streams[i], cudaStreamNonBlocking) != cudaSuccess) {
            snprintf(g_comm_err, sizeof g_comm_err, "stream creation failed on device %d", c->devices[i]);
            return FXG_ERR_CUDA;
        }
        c->streams[i] = c->own_streams[i];
    }
    return FXG_OK;
}

extern "C" int fxg_comm_init_all(int ndev, const int *devices, fxg_comm **out)
{
    if (!out || ndev < 1 || !devices) return FXG_ERR_ARG;
    *out = NULL;
    int rc = nccl_load();
    if (rc) return rc;
    fxg_comm *c = comm_alloc(ndev, ndev);
    if (!c) return FXG_ERR_NOMEM;
    for (int i = 0; i < ndev; i++) { c->devices[i] = devices[i]; c->ranks[i] = i; }
    ncclResult_t r;
    { StdoutGuard g; r = g_nccl.CommInitAll((ncclComm_t *)c->comms, ndev, devices); }
    if (r != ncclSuccess) { nccl_fail(NULL, "ncclCommInitAll", r); fxg_comm_free(c); return FXG_ERR_NCCL; }
    if ((rc = comm_make_streams(c)) != FXG_OK) { fxg_comm_free(c); return rc; }
    *out = c;
    return FXG_OK;
}

extern "C" int fxg_comm_unique_id(void *id_out)
{
    if (!id_out) return FXG_ERR_ARG;
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(NULL, "ncclGetUniqueId", r);
    static_assert(sizeof(id) == FXG_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof id);
    return FXG_OK;
}

extern "C" int fxg_comm_init_rank(int device, int nranks, int rank, const void *id, fxg_comm **out)
{
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return FXG_ERR_ARG;
    *out = NULL;
    int rc = nccl_load();
    if (rc) return rc;
    fxg_comm *c = comm_alloc(1, nranks);
    if (!c) return FXG_ERR_NOMEM;
    c->devices[0] = device; c->ranks[0] = rank;
    if (cudaSetDevice(device) != cudaSuccess) { snprintf(g_comm_err, sizeof g_comm_err, "cudaSetDevice(%d) failed", device); fxg_comm_free(c); return FXG_ERR_CUDA; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclResult_t r;
    { StdoutGuard g; r = g_nccl.CommInitRank((ncclComm_t *)&c->comms[0], nranks, uid, rank); }
    if (r != ncclSuccess) { nccl_fail(NULL, "ncclCommInitRank", r); fxg_comm_free(c); return FXG_ERR_NCCL; }
    if ((rc = comm_make_streams(c)) != FXG_OK) { fxg_comm_free(c); return rc; }
    *out = c;
    return FXG_OK;
}

extern "C" int fxg_comm_nranks(const fxg_comm *c) { return c ? c->nranks : 0; }
extern "C" int fxg_comm_nlocal(const fxg_comm *c) { return c ? c->nlocal : 0; }
extern "C" int fxg_comm_rank(const fxg_comm *c, int i) { return (c && i >= 0 && i < c->nlocal) ? c->ranks[i] : -1; }
extern "C" int fxg_comm_device(const fxg_comm *c, int i) { return (c && i >= 0 && i < c->nlocal) ? c->devices[i] : -1; }
extern "C" int64_t fxg_comm_bytes_sent(const fxg_comm *c) { return c ? c->bytes_sent : 0; }
extern "C" int64_t fxg_comm_collectives(const fxg_comm *c) { return c ? c->n_collectives : 0; }

extern "C" int fxg_comm_set_stream(fxg_comm *c, int i, void *cuda_stream, int adopt)
{
    if (!c || i < 0 || i >= c->nlocal) return FXG_ERR_ARG;
    c->streams[i] = adopt ? (cudaStream_t)cuda_stream : c->own_streams[i];
    c->adopted = 0;
    for (int k = 0; k < c->nlocal; k++) if (c->streams[k] != c->own_streams[k]) c->adopted = 1;
    return FXG_OK;
}

extern "C" int fxg_comm_sync(fxg_comm *c)
{
    if (!c) return FXG_ERR_ARG;
    for (int i = 0; i < c->nlocal; i++) {
        cudaSetDevice(c->devices[i]);
        cudaError_t e = cudaStreamSynchronize(c->streams[i]);
        if (e != cudaSuccess) { snprintf(c->err, sizeof c->err, "collective failed on device %d: %s", c->devices[i], cudaGetErrorString(e)); return FXG_ERR_CUDA; }
    }
    return FXG_OK;
}

// producers that ran on other streams than the communicator's own: wait for the whole device once (the tools call the
// all-reduce once per job).  With an adopted stream the caller's stream order already covers it.
static void comm_wait_producers(fxg_comm *c)
{
    if (c->adopted) return;
    for (int i = 0; i < c->nlocal; i++) { cudaSetDevice(c->devices[i]); cudaDeviceSynchronize(); }
}

// In-place sum of bufs_dev[i] (count u64 words each, resident on the i-th local device) across all ranks.
extern "C" int fxg_comm_allreduce_u64(fxg_comm *c, uint64_t *const *bufs_dev, size_t count)
{
    if (!c || !bufs_dev) return FXG_ERR_ARG;
    comm_wait_producers(c);
    ncclResult_t r = g_nccl.GroupStart();
    for (int i = 0; r == ncclSuccess && i < c->nlocal; i++)
        r = g_nccl.AllReduce(bufs_dev[i], bufs_dev[i], count, ncclUint64, ncclSum, (ncclComm_t)c->comms[i], c->streams[i]);
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return nccl_fail(c, "ncclAllReduce", r);
    c->n_collectives++;
    c->bytes_sent += (int64_t)c->nlocal * (int64_t)count * 8 * 2 * (c->nranks - 1) / c->nranks;     // ring-equivalent volume
    return c->adopted ? FXG_OK : fxg_comm_sync(c);
}

// bytes-wide all-gather: every rank contributes `bytes` bytes; recv_dev[i] receives nranks * bytes, in rank order
extern "C" int fxg_comm_allgather(fxg_comm *c, const void *const *send_dev, void *const *recv_dev, size_t bytes)
{
    if (!c || !send_dev || !recv_dev) return FXG_ERR_ARG;
    ncclResult_t r = g_nccl.GroupStart();
    for (int i = 0; r == ncclSuccess && i < c->nlocal; i++)
        r = g_nccl.AllGather(send_dev[i], recv_dev[i], bytes, ncclChar, (ncclComm_t)c->comms[i], c->streams[i]);
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return nccl_fail(c, "ncclAllGather", r);
    c->n_collectives++;
    c->bytes_sent += (int64_t)c->nlocal * (int64_t)bytes * (c->nranks - 1);
    return FXG_OK;
}

// The collapser's exchange (grouped ncclSend/ncclRecv by owner = hash mod G).  For local device i (global
// rank r): elements [send_off[i*G + d], +send_cnt[i*G + d]) of send_dev[i] go to rank d; elements from rank s arrive at
// [recv_off[i*G + s], +recv_cnt[i*G + s]) of recv_dev[i].  Counts and offsets are in elements of elem_bytes bytes (host
// arrays); the caller has already exchanged the count matrix (fxg_comm_allgather).  Enqueued on the communicator's streams.
extern "C" int fxg_comm_alltoallv(fxg_comm *c, const void *const *send_dev, const int64_t *send_off, const int64_t *send_cnt,
                                  void *const *recv_dev, const int64_t *recv_off, const int64_t *recv_cnt, size_t elem_bytes)
{
    if (!c || !send_dev || !recv_dev || !send_off || !send_cnt || !recv_off || !recv_cnt || elem_bytes == 0) return FXG_ERR_ARG;
    const int G = c->nranks;
    ncclResult_t r = g_nccl.GroupStart();
    for (int i = 0; r == ncclSuccess && i < c->nlocal; i++) {
        for (int p = 0; r == ncclSuccess && p < G; p++) {
            const int64_t sc = send_cnt[(size_t)i * G + p], rc_ = recv_cnt[(size_t)i * G + p];
            if (sc > 0) {
                r = g_nccl.Send((const char *)send_dev[i] + (size_t)send_off[(size_t)i * G + p] * elem_bytes, (size_t)sc * elem_bytes, ncclChar, p,
                                (ncclComm_t)c->comms[i], c->streams[i]);
                if (p != c->ranks[i]) c->bytes_sent += sc * (int64_t)elem_bytes;
            }
            if (r == ncclSuccess && rc_ > 0)
                r = g_nccl.Recv((char *)recv_dev[i] + (size_t)recv_off[(size_t)i * G + p] * elem_bytes, (size_t)rc_ * elem_bytes, ncclChar, p,
                                (ncclComm_t)c->comms[i], c->streams[i]);
        }
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return nccl_fail(c, "ncclSend/ncclRecv", r);
    c->n_collectives++;
    return FXG_OK;
}

// Variable-size gather to one rank: local device i sends cnt[rank_i] elements; the local device whose rank is `root`
// receives every rank's block at off[s] (elements) of root_recv_dev.  cnt/off: host arrays of nranks entries.
extern "C" int fxg_comm_gatherv(fxg_comm *c, const void *const *send_dev, const int64_t *cnt, const int64_t *off, void *root_recv_dev,
                                int root, size_t elem_bytes)
{
    if (!c || !send_dev || !cnt || !off || root < 0 || root >= c->nranks || elem_bytes == 0) return FXG_ERR_ARG;
    ncclResult_t r = g_nccl.GroupStart();
    for (int i = 0; r == ncclSuccess && i < c->nlocal; i++) {
        const int me = c->ranks[i];
        if (cnt[me] > 0) {
            r = g_nccl.Send(send_dev[i], (size_t)cnt[me] * elem_bytes, ncclChar, root, (ncclComm_t)c->comms[i], c->streams[i]);
            if (me != root) c->bytes_sent += cnt[me] * (int64_t)elem_bytes;
        }
        if (me == root) {
            if (!root_recv_dev) { g_nccl.GroupEnd(); return FXG_ERR_ARG; }
            for (int s = 0; r == ncclSuccess && s < c->nranks; s++)
                if (cnt[s] > 0)
                    r = g_nccl.Recv((char *)root_recv_dev + (size_t)off[s] * elem_bytes, (size_t)cnt[s] * elem_bytes, ncclChar, s,
                                    (ncclComm_t)c->comms[i], c->streams[i]);
        }
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return nccl_fail(c, "ncclSend/ncclRecv (gather)", r);
    c->n_collectives++;
    return FXG_OK;
}
