"""SURVEY §8(f-3), first version: fxg_pipeline_dev (map-type tools chained in HBM) vs the oracle's per-tool functions composed
stage by stage with the survivors compacted — the composition that tests/test_pipeline_oracle.py pins against the reference
binaries in a shell pipe."""
import os

import numpy as np
import pytest

import helpers as H
from test_pipeline_oracle import ADAPTER, stage_clip, stage_filter, stage_trim, synth_input

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def oracle_final_len(seq, qual, lens, stages):
    n = seq.shape[0]
    idx = np.arange(n)
    final = np.full(n, -1, np.int32)
    for st in stages:
        new = st(seq, qual, lens)
        keep = np.flatnonzero(new >= 0)
        seq, qual, lens, idx = np.ascontiguousarray(seq[keep]), np.ascontiguousarray(qual[keep]), new[keep].astype(np.int32), idx[keep]
    final[idx] = lens
    return final


def test_pipeline_matches_composed_oracle():
    import fastx_toolkit_b200 as F
    ctx = F.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    seq, qual, L = synth_input()
    n, stride = seq.shape
    dseq, dqual = torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda()
    clip = F.ClipOpts(adapter=ADAPTER, min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
    import ctypes as C
    cases = [
        ([F.Stage(2, 0, 0, C.addressof(clip)), F.Stage(0, 20, 20, None), F.Stage(1, 20, 90, None)],
         [lambda s, q, l: stage_clip(s, q, l, 20), lambda s, q, l: stage_trim(s, q, l, 20, 20), lambda s, q, l: stage_filter(s, q, l, 20, 90)], None),
        ([F.Stage(0, 25, 30, None), F.Stage(1, 30, 50, None)],
         [lambda s, q, l: stage_trim(s, q, l, 25, 30), lambda s, q, l: stage_filter(s, q, l, 30, 50)], None),
        ([F.Stage(1, 20, 90, None), F.Stage(0, 30, 10, None), F.Stage(1, 32, 80, None), F.Stage(0, 35, 40, None)],
         [lambda s, q, l: stage_filter(s, q, l, 20, 90), lambda s, q, l: stage_trim(s, q, l, 30, 10), lambda s, q, l: stage_filter(s, q, l, 32, 80),
          lambda s, q, l: stage_trim(s, q, l, 35, 40)], "ragged"),
        ([F.Stage(0, 41, 1, None), F.Stage(1, 20, 90, None)],                       # everything dropped by the first stage
         [lambda s, q, l: stage_trim(s, q, l, 41, 1), lambda s, q, l: stage_filter(s, q, l, 20, 90)], None),
        ([F.Stage(2, 0, 0, C.addressof(clip))], [lambda s, q, l: stage_clip(s, q, l, 20)], None),
    ]
    for stages, ostages, mode in cases:
        s2, q2 = seq.copy(), qual.copy()
        lens = np.full(n, L, np.int32)
        dlens = None
        if mode == "ragged":
            lens = H.ragged(s2, q2, np.random.default_rng(9), min_len=5)
            dlens = torch.from_numpy(lens).cuda()
        ds, dq = (dseq, dqual) if mode is None else (torch.from_numpy(s2).cuda(), torch.from_numpy(q2).cuda())
        final = torch.full((n,), 12345, dtype=torch.int32, device="cuda")
        ctx.report_reset()
        alive = ctx.pipeline_dev(ctx.batch(ds, dq, n, stride, L if dlens is None else 0, dlens), 33, stages, final)
        exp = oracle_final_len(s2, q2, lens, ostages)
        got = final.cpu().numpy()
        assert np.array_equal(got, exp), (len(stages), int((got != exp).sum()))
        assert alive == int((exp >= 0).sum())
    # the clipper after another stage is refused (stale-buffer semantics are not implemented on the device)
    with pytest.raises(F.FxgError):
        ctx.pipeline_dev(ctx.batch(dseq, dqual, n, stride, L), 33, [F.Stage(0, 20, 20, None), F.Stage(2, 0, 0, C.addressof(clip))],
                         torch.empty(n, dtype=torch.int32, device="cuda"))
    ctx.close()


@pytest.mark.skipif(os.environ.get("FXG_PIPE_STALE") != "1", reason="experimental: the device-side stale-buffer scan has not run on a GPU yet (set FXG_PIPE_STALE=1)")
def test_pipeline_clipper_after_trimmer_experimental():
    """trim | clip: the clipper on mixed lengths, stale-buffer rows from the prefix-overwrite scan (fxg_pipeline.cu)"""
    import ctypes as C
    import fastx_toolkit_b200 as F
    ctx = F.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    seq, qual, L = synth_input()
    n, stride = seq.shape
    clip = F.ClipOpts(adapter=ADAPTER, min_length=15, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=0, min_adapter_len=0)
    final = torch.full((n,), 12345, dtype=torch.int32, device="cuda")
    alive = ctx.pipeline_dev(ctx.batch(torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda(), n, stride, L), 33,
                             [F.Stage(0, 25, 30, None), F.Stage(2, 0, 0, C.addressof(clip))], final)
    exp = oracle_final_len(seq, qual, np.full(n, L, np.int32),
                           [lambda s, q, l: stage_trim(s, q, l, 25, 30), lambda s, q, l: stage_clip(s, q, l, 15, discard_unknown=0)])
    assert np.array_equal(final.cpu().numpy(), exp) and alive == int((exp >= 0).sum())
    ctx.close()
