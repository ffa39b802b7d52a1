// fxg_comm.cu — the one collective the C tools need natively: NCCL all-reduce (sum, u64) of the per-GPU
// fastx_quality_stats histograms when one process drives several GPUs.  NCCL is resolved at run
// time (dlopen "libnccl.so.2") so that libfxg.so has no load-time dependency on it and shares the NCCL instance a
// host program (e.g. PyTorch) may already have loaded.  Multi-process jobs (torchrun) use their own communicator.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fxg.h"

struct fxg_comm {
    int ndev;
    int *devices;
    ncclComm_t *comms;
    cudaStream_t *streams;
    void *lib;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
    char err[256];
};

static char g_comm_err[256] = "";
extern "C" const char *fxg_comm_error(const fxg_comm *c) { return c ? c->err : g_comm_err; }

extern "C" void fxg_comm_free(fxg_comm *c)
{
    if (!c) return;
    for (int i = 0; i < c->ndev; i++) {
        if (c->comms && c->comms[i] && c->CommDestroy) c->CommDestroy(c->comms[i]);
        if (c->streams && c->streams[i]) { cudaSetDevice(c->devices[i]); cudaStreamDestroy(c->streams[i]); }
    }
    free(c->comms); free(c->streams); free(c->devices);
    free(c);
}

extern "C" int fxg_comm_init_all(int ndev, const int *devices, fxg_comm **out)
{
    if (!out || ndev < 1 || !devices) return FXG_ERR_ARG;
    *out = NULL;
    fxg_comm *c = (fxg_comm *)calloc(1, sizeof(fxg_comm));
    if (!c) return FXG_ERR_NOMEM;
    c->ndev = ndev;
    c->devices = (int *)malloc(sizeof(int) * ndev);
    c->comms = (ncclComm_t *)calloc(ndev, sizeof(ncclComm_t));
    c->streams = (cudaStream_t *)calloc(ndev, sizeof(cudaStream_t));
    memcpy(c->devices, devices, sizeof(int) * ndev);
    // the drop-in tools write their DATA to stdout: NCCL's version banner / debug log must never land there
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    c->lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!c->lib) { snprintf(g_comm_err, sizeof g_comm_err, "dlopen(libnccl.so.2): %s", dlerror()); fxg_comm_free(c); return FXG_ERR_NCCL; }
    *(void **)&c->CommInitAll = dlsym(c->lib, "ncclCommInitAll");
    *(void **)&c->CommDestroy = dlsym(c->lib, "ncclCommDestroy");
    *(void **)&c->AllReduce = dlsym(c->lib, "ncclAllReduce");
    *(void **)&c->GroupStart = dlsym(c->lib, "ncclGroupStart");
    *(void **)&c->GroupEnd = dlsym(c->lib, "ncclGroupEnd");
    *(void **)&c->GetErrorString = dlsym(c->lib, "ncclGetErrorString");
    if (!c->CommInitAll || !c->CommDestroy || !c->AllReduce || !c->GroupStart || !c->GroupEnd) {
        snprintf(g_comm_err, sizeof g_comm_err, "libnccl.so.2 lacks a required symbol");
        fxg_comm_free(c);
        return FXG_ERR_NCCL;
    }
    // NCCL prints its version banner with printf() when NCCL_DEBUG=VERSION/WARN (this image exports NCCL_DEBUG=VERSION):
    // point fd 1 at stderr while the communicators are created, then restore it.
    fflush(stdout);
    const int saved_stdout = dup(STDOUT_FILENO);
    if (saved_stdout >= 0) dup2(STDERR_FILENO, STDOUT_FILENO);
    ncclResult_t r = c->CommInitAll(c->comms, ndev, devices);
    fflush(stdout);
    if (saved_stdout >= 0) { dup2(saved_stdout, STDOUT_FILENO); close(saved_stdout); }
    if (r != ncclSuccess) {
        snprintf(g_comm_err, sizeof g_comm_err, "ncclCommInitAll: %s", c->GetErrorString ? c->GetErrorString(r) : "error");
        fxg_comm_free(c);
        return FXG_ERR_NCCL;
    }
    for (int i = 0; i < ndev; i++) {
        if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
            snprintf(g_comm_err, sizeof g_comm_err, "stream creation failed on device %d", devices[i]);
            fxg_comm_free(c);
            return FXG_ERR_CUDA;
        }
    }
    *out = c;
    return FXG_OK;
}

// In-place sum of bufs_dev[i] (count u64 words each, resident on devices[i]) across all GPUs of the communicator.
extern "C" int fxg_comm_allreduce_u64(fxg_comm *c, uint64_t *const *bufs_dev, size_t count)
{
    if (!c || !bufs_dev) return FXG_ERR_ARG;
    for (int i = 0; i < c->ndev; i++) { cudaSetDevice(c->devices[i]); cudaDeviceSynchronize(); }   // producers may use other streams
    ncclResult_t r = c->GroupStart();
    for (int i = 0; r == ncclSuccess && i < c->ndev; i++)
        r = c->AllReduce(bufs_dev[i], bufs_dev[i], count, ncclUint64, ncclSum, c->comms[i], c->streams[i]);
    ncclResult_t r2 = c->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { snprintf(c->err, sizeof c->err, "ncclAllReduce: %s", c->GetErrorString ? c->GetErrorString(r) : "error"); return FXG_ERR_NCCL; }
    for (int i = 0; i < c->ndev; i++) {
        cudaSetDevice(c->devices[i]);
        if (cudaStreamSynchronize(c->streams[i]) != cudaSuccess) { snprintf(c->err, sizeof c->err, "all-reduce failed on device %d", c->devices[i]); return FXG_ERR_CUDA; }
    }
    return FXG_OK;
}
