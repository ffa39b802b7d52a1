#!/usr/bin/env python
"""Exercise every kernel once on small data (meant to run under compute-sanitizer memcheck / racecheck)."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastx_toolkit_b200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ctx = F.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for L in (150, 100, 37, 250):
    S = (L + 15) // 16 * 16
    dseq = torch.empty((n, S), dtype=torch.uint8, device="cuda")
    dqual = torch.empty((n, S), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(dseq, dqual, n, L, S, 20260926 + L, 2, 33)
    lens = torch.randint(1, L + 1, (n,), dtype=torch.int32, device="cuda")
    for b in (ctx.batch(dseq, dqual, n, S, L), ctx.batch(dseq, dqual, n, S, 0, lens)):
        out = torch.empty(n, dtype=torch.int32, device="cuda"); keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        os_, oq = torch.empty_like(dseq), torch.empty_like(dqual)
        hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
        ctx.trim_dev(b, 33, 20, 20, out); ctx.filter_dev(b, 33, 20, 90, keep); ctx.revcomp_dev(b, 33, os_, oq)
        ctx.stats_accum_dev(b, 33, hist, L)
        o = F.ClipOpts(adapter=b"AGATCGGAAGAGC", min_length=5, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
        ctx.clip_dev(b, None, 33, o, out, keep)
        ctx.mask_dev(b, 33, 20, ord("N"), os_, keep); ctx.artifacts_dev(b, 33, keep); ctx.validate_dev(b, 33)
        hs = torch.empty(n, dtype=torch.int64, device="cuda"); ctx.hash_dev(b, hs)
        for cfg in ("0,0,0,0",):       # CTA-tile kernels too
            os.environ["FXG_TUNE"] = cfg
            ctx.trim_dev(b, 33, 20, 20, out); ctx.revcomp_dev(b, 33, os_, oq); ctx.stats_accum_dev(b, 33, hist, L)
            os.environ.pop("FXG_TUNE")
        rep = ctx.sync()
        assert rep.first_bad_read == -1, rep.first_bad_read
    col = F.Collapser(0, n, S)
    col.add(ctx.batch(dseq, None, n, S, L)); col.finish(True); col.close()
txt = subprocess.run([os.path.join(ROOT, "bin", "fxg_synth"), "-n", str(n), "-l", "150"], stdout=subprocess.PIPE, check=True).stdout
tp = F.TextPipe(ctx, len(txt) + 4096)
for op in (0, 1, 2):
    o, rep = tp.run(op, txt, 33, 20, 20 if op == 0 else 90)
    assert rep.anomaly == 0 and rep.n_records == n
tp.close()
print("smoke_ops OK: launches", ctx.launches())
