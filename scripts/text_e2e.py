#!/usr/bin/env python
"""Text -> text end-to-end throughput of the GPU text path (fxg_text_run_host) with W worker threads, each with its
own context + text pipeline (ctypes releases the GIL): pinned FASTQ text in host memory -> trimmed FASTQ text in host
memory.  Usage: text_e2e.py [reads] [workers] [chunk_reads]"""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402
import fastx_toolkit_b200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunk_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 250_000
base_reads = 1_000_000
txt = subprocess.run([os.path.join(ROOT, "bin", "fxg_synth"), "-n", str(base_reads), "-l", "150"], stdout=subprocess.PIPE, check=True).stdout
lines = txt.split(b"\n")
chunk = b"\n".join(lines[: 4 * chunk_reads]) + b"\n"
nchunks = n // chunk_reads
cb = len(chunk)
host_in = torch.empty(cb, dtype=torch.uint8).pin_memory()
host_in.numpy()[:] = np.frombuffer(chunk, np.uint8)
workers = []
for w in range(W):
    ctx = F.Context(0)
    tp = F.TextPipe(ctx, cb + 4096)
    out = torch.empty(cb + cb // 4 + 64, dtype=torch.uint8).pin_memory()
    workers.append((ctx, tp, out))
L = F.lib()


def run_chunks(w, count, res):
    ctx, tp, out = workers[w]
    rep = F.TextReport()
    tot = 0
    for _ in range(count):
        rc = L.fxg_text_run_host(tp.h, 0, host_in.data_ptr(), cb, 33, 20, 20, out.data_ptr(), C.byref(rep))
        assert rc == 0 and rep.anomaly == 0 and rep.n_records == chunk_reads
        tot += rep.out_bytes
    res[w] = tot


for trial in range(2):
    res = [0] * W
    per = [nchunks // W + (1 if w < nchunks % W else 0) for w in range(W)]
    t0 = time.perf_counter()
    th = [threading.Thread(target=run_chunks, args=(w, per[w], res)) for w in range(W)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    print("text e2e trial %d: %d reads in %d chunks of %.1f MB, %d workers: %.3f s  %.1f Mreads/s  (in %.1f GB/s, out %.1f GB/s)" %
          (trial, nchunks * chunk_reads, nchunks, cb / 1e6, W, dt, nchunks * chunk_reads / dt / 1e6, nchunks * cb / dt / 1e9, sum(res) / dt / 1e9), flush=True)
