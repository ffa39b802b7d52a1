#!/usr/bin/env python
"""GPU tuning sweep for K-TRIM / K-FILTER / K-REVCOMP (FXG_TUNE=ring,g,stages,ctas). Prints GB/s (algorithmic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastx_toolkit_b200 as F  # noqa: E402

L, S = int(os.environ.get("TL", 150)), None
S = (L + 15) // 16 * 16
n = int(os.environ.get("TN", 50_000_000))
ctx = F.Context(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
dseq = torch.empty((n, S), dtype=torch.uint8, device="cuda")
dqual = torch.empty((n, S), dtype=torch.uint8, device="cuda")
ctx.synth_dev(dseq, dqual, n, L, S, 20260926, 0, 33)
out = torch.empty(n, dtype=torch.int32, device="cuda")
keep = torch.empty(n, dtype=torch.uint8, device="cuda")
b = ctx.batch(dseq, dqual, n, S, L)
torch.cuda.synchronize()
ref = None


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


configs = sys.argv[1:] or ["-1,0,0,0"]
for cfg in configs:
    os.environ["FXG_TUNE"] = cfg
    try:
        ms = timeit(lambda: ctx.trim_dev(b, 33, 20, 20, out))
        cur = out.clone()
        if ref is None:
            ref = cur
        ok = bool(torch.equal(cur, ref))
        msf = timeit(lambda: ctx.filter_dev(b, 33, 20, 90, keep))
        print("FXG_TUNE=%-10s trim %.3f ms %7.1f GB/s (%.2f Gr/s) same=%s | filter %.3f ms %7.1f GB/s" %
              (cfg, ms, n * (2 * L + 4) / ms / 1e6, n / ms / 1e6, ok, msf, n * (2 * L + 1) / msf / 1e6), flush=True)
    except Exception as e:  # noqa: BLE001
        print("FXG_TUNE=%s failed: %s" % (cfg, e), flush=True)
os.environ.pop("FXG_TUNE", None)
oseq, oqual = torch.empty_like(dseq), torch.empty_like(dqual)
for cfg in ["0,0,0,0", "-1,0,0,0", "1,2,0,6"]:
    os.environ["FXG_TUNE"] = cfg
    try:
        ms = timeit(lambda: ctx.revcomp_dev(b, 33, oseq, oqual))
        if cfg == "0,0,0,0":
            rs, rq = oseq.clone(), oqual.clone()
        else:
            assert torch.equal(rs, oseq) and torch.equal(rq, oqual), "revcomp variants disagree"
        print("FXG_TUNE=%-10s revcomp %.3f ms %7.1f GB/s (%.2f Gr/s)" % (cfg, ms, n * 4 * L / ms / 1e6, n / ms / 1e6), flush=True)
    except Exception as e:  # noqa: BLE001
        print("revcomp FXG_TUNE=%s failed: %s" % (cfg, e), flush=True)
os.environ.pop("FXG_TUNE", None)
hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
ns = min(n, 20_000_000)
bs = ctx.batch(dseq, dqual, ns, S, L)
for cfg in ["1,1,1,0", "1,2,1,0", "1,4,1,0", "1,2,2,0", "0,0,0,0"]:
    os.environ["FXG_TUNE"] = cfg
    ms = timeit(lambda: ctx.stats_accum_dev(bs, 33, hist, L), reps=3)
    print("FXG_TUNE=%-10s stats %.3f ms %7.1f GB/s (%.2f Gr/s, %.1f Gsamples/s)" % (cfg, ms, ns * 2 * L / ms / 1e6, ns / ms / 1e6, ns * L / ms / 1e6), flush=True)
os.environ.pop("FXG_TUNE", None)
nc = min(n, 4_000_000)
aseq = torch.empty((nc, S), dtype=torch.uint8, device="cuda"); aqual = torch.empty((nc, S), dtype=torch.uint8, device="cuda")
ctx.synth_dev(aseq, aqual, nc, L, S, 20260927, 2, 33)
olen = torch.empty(nc, dtype=torch.int32, device="cuda")
for ad, payload in [(b"AGATCGGAAGAGC", "0"), (b"AGATCGGAAGAGC", "1"), (b"AGATCGGAAGAGCACACGTCTGAACTCC", "0"), (b"AGATCGGAAGAGCACACGTCTGAACTCC", "1")]:
    os.environ["FXG_CLIP_PAYLOAD"] = payload
    o = F.ClipOpts(adapter=ad, min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    bc = ctx.batch(aseq, aqual, nc, S, L)
    ms = timeit(lambda: ctx.clip_dev(bc, None, 33, o, olen), reps=3)
    print("clip H=%d %s: %.3f ms %.1f Mreads/s %.1f Gcells/s" % (len(ad), "forward-payload" if payload == "1" else "origin-bits+backtrace", ms, nc / ms / 1e3, nc * L * len(ad) / ms / 1e6), flush=True)
rep = ctx.sync()
print("first_bad", rep.first_bad_read, "clip kept", rep.n_out)
