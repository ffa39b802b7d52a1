#!/usr/bin/env python
"""Launch one op a few times on synthetic HBM-resident data (for ncu captures): run_ops.py <trim|filter|revcomp|stats|clip|collapse> [n] [L]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastx_toolkit_b200 as F  # noqa: E402

op = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 150
S = (L + 15) // 16 * 16
ctx = F.Context(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
kind = {"clip": 2, "pipeline": 2, "collapse": 3}.get(op, 0)
dseq = torch.empty((n, S), dtype=torch.uint8, device="cuda")
dqual = torch.empty((n, S), dtype=torch.uint8, device="cuda")
ctx.synth_dev(dseq, dqual, n, L, S, 20260926, kind, 33)
b = ctx.batch(dseq, dqual, n, S, L)
reps = 3


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if op == "trim":
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = timed(lambda: ctx.trim_dev(b, 33, 20, 20, out)); bytes_ = n * (2 * L + 4)
elif op == "filter":
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.filter_dev(b, 33, 20, 90, out)); bytes_ = n * (2 * L + 1)
elif op == "revcomp":
    os_, oq = torch.empty_like(dseq), torch.empty_like(dqual)
    ms = timed(lambda: ctx.revcomp_dev(b, 33, os_, oq)); bytes_ = n * 4 * L
elif op == "stats":
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ms = timed(lambda: ctx.stats_accum_dev(b, 33, hist, L)); bytes_ = n * 2 * L
elif op == "clip":
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    o = F.ClipOpts(adapter=b"AGATCGGAAGAGC", min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    ms = timed(lambda: ctx.clip_dev(b, None, 33, o, out)); bytes_ = n * (L + 4)
    print("clip: %.1f Gcells/s" % (n * L * 13 / ms / 1e6))
elif op == "pipeline":
    import ctypes as C
    fin = torch.empty(n, dtype=torch.int32, device="cuda")
    co = F.ClipOpts(adapter=b"AGATCGGAAGAGC", min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    stages = [F.Stage(2, 0, 0, C.addressof(co)), F.Stage(0, 20, 20, None), F.Stage(1, 20, 90, None)]
    alive = [0]
    def run_pipe():
        alive[0] = ctx.pipeline_dev(b, 33, stages, fin)
    ms = timed(run_pipe); bytes_ = n * (2 * L + 4)
    print("pipeline clip|trim|filter: %d of %d reads survive" % (alive[0], n))
elif op == "mask":
    os_ = torch.empty_like(dseq); fl = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.mask_dev(b, 33, 20, ord("N"), os_, fl)); bytes_ = n * 3 * L
elif op == "artifacts":
    fl = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.artifacts_dev(b, 33, fl)); bytes_ = n * (2 * L + 1)
elif op == "validate":
    ms = timed(lambda: ctx.validate_dev(b, 33)); bytes_ = n * 2 * L
elif op == "collapse":
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    col = F.Collapser(0, n, S)
    col.add(ctx.batch(dseq, None, n, S, L))
    t1 = time.perf_counter()
    u = col.finish(True)
    t2 = time.perf_counter()
    print("collapse n=%d L=%d: unique=%d  hash+dedup %.1f ms (%.1f Mkeys/s)  order %.1f ms  launches %d" %
          (n, L, u, (t1 - t0) * 1e3, n / (t1 - t0) / 1e6, (t2 - t1) * 1e3, col.launches()))
    t2 = time.perf_counter()
    u = col.finish(True)             # second pass: the ordering scratch now comes from the cached pool
    t3 = time.perf_counter()
    print("collapse (steady state) order %.1f ms for %d uniques" % ((t3 - t2) * 1e3, u))
    col.close()
    sys.exit(0)
print("%s n=%d L=%d: %.3f ms  %.1f GB/s algorithmic  %.2f Greads/s" % (op, n, L, ms, bytes_ / ms / 1e6, n / ms / 1e6))
