#!/usr/bin/env python
"""Host<->device copy ceiling of the box, every GPU at once (torchrun, one rank per GPU) or alone:
pinned host memory, 256 MB blocks, H2D alone, D2H alone, both directions at once.  Prints per-rank and aggregate GB/s —
the ceiling the end-to-end text path (FASTQ text in, FASTQ text out) can reach at N GPUs."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
MB = int(os.environ.get("PROBE_MB", 256))
reps = int(os.environ.get("PROBE_REPS", 12))
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_in.fill_(65)
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
d_out = torch.ones(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(h2d, d2h):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return reps * MB * (1 << 20) / dt / 1e9


run(True, True)
res = {"h2d_only": run(True, False), "d2h_only": run(False, True), "both_each_direction": run(True, True)}
if rank == 0:
    print("pcie_probe world=%d block=%d MB: per-GPU GB/s %s ; aggregate GB/s %s" %
          (world, MB, {k: round(v, 1) for k, v in res.items()}, {k: round(v * world, 1) for k, v in res.items()}), flush=True)
if world > 1:
    dist.destroy_process_group()
