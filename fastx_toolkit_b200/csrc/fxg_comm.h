// fxg_comm.h — the communicator object behind include/fxg.h's fxg_comm_* entry points (internal: shared by fxg_comm.cu,
// which owns the NCCL calls, and fxg_dcollapse.cu, the multi-GPU collapser built on them).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct fxg_comm {
    int nlocal;                 // GPUs driven by this process
    int nranks;                 // GPUs in the job
    int *devices;               // [nlocal] CUDA device ordinals
    int *ranks;                 // [nlocal] global rank of each local GPU
    void **comms;               // [nlocal] ncclComm_t
    cudaStream_t *streams;      // [nlocal] stream the collectives are enqueued on (own or adopted)
    cudaStream_t *own_streams;  // [nlocal]
    int adopted;                // some stream was adopted from the caller: no device-wide wait before a collective
    int64_t bytes_sent;         // payload bytes this process put on the wire (self-sends excluded)
    int64_t n_collectives;      // NCCL groups issued
    char err[256];
};
