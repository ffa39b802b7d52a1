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


def oracle_final_len(seq, qual, lens, stages, entering=None):
    n = seq.shape[0]
    idx = np.arange(n)
    final = np.full(n, -1, np.int32)
    for st in stages:
        if entering is not None:
            entering.append(seq.shape[0])
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
        entering = []
        exp = oracle_final_len(s2, q2, lens, ostages, entering)
        got = final.cpu().numpy()
        assert np.array_equal(got, exp), (len(stages), int((got != exp).sum()))
        assert alive == int((exp >= 0).sum())
        # the survivor counts stay on the device between the stages (one read-back per call); the report still counts what
        # each stage really saw, not the bound its kernels were launched over
        assert ctx.report().n_in == sum(entering), (ctx.report().n_in, entering)
    ctx.close()


def test_pipeline_clipper_on_mixed_lengths():
    """trim | clip and a ragged batch straight into the clipper: mixed lengths, the reference aligner's stale-buffer rows
     built on the device by the prefix-overwrite scan (fxg_pipeline.cu launch_stale_rows)"""
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
    # ragged input, clipper first, then two more stages
    s2, q2 = seq.copy(), qual.copy()
    lens = H.ragged(s2, q2, np.random.default_rng(3), min_len=8)
    final = torch.full((n,), 12345, dtype=torch.int32, device="cuda")
    alive = ctx.pipeline_dev(ctx.batch(torch.from_numpy(s2).cuda(), torch.from_numpy(q2).cuda(), n, stride, 0, torch.from_numpy(lens).cuda()), 33,
                             [F.Stage(2, 0, 0, C.addressof(clip)), F.Stage(1, 20, 80, None), F.Stage(0, 22, 12, None)], final)
    exp = oracle_final_len(s2, q2, lens, [lambda s, q, l: stage_clip(s, q, l, 15, discard_unknown=0), lambda s, q, l: stage_filter(s, q, l, 20, 80),
                                          lambda s, q, l: stage_trim(s, q, l, 22, 12)])
    assert np.array_equal(final.cpu().numpy(), exp) and alive == int((exp >= 0).sum())
    ctx.close()


def collapser_text(col, u):
    stride = col.stride
    oseq, olen, ocnt = np.zeros((u, stride), np.uint8), np.zeros(u, np.int32), np.zeros(u, np.uint64)
    col.fetch(oseq, olen, ocnt, None, None)
    return b"".join(b">%d-%d\n" % (k + 1, int(ocnt[k])) + oseq[k, :olen[k]].tobytes() + b"\n" for k in range(u))


@pytest.mark.parametrize("name", ["clip_trim_filter_collapse", "trim_clip_collapse"])
def test_pipeline_with_collapser_equals_reference_shell_pipe(name):
    """the whole chain on the device, the collapser as last stage: the FASTA it yields must be the bytes the reference tools
    produce in a shell pipe (digests in tests/golden/pipeline.json, made from the reference binaries; the same composition
    is replayed against the live binaries by tests/test_pipeline_oracle.py)"""
    import ctypes as C
    import hashlib
    import json
    import fastx_toolkit_b200 as F
    from test_pipeline_oracle import GOLD, oracle_pipeline
    ctx = F.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    seq, qual, L = synth_input()
    n, stride = seq.shape
    col = F.Collapser(0, n, stride)
    if name == "clip_trim_filter_collapse":
        clip = F.ClipOpts(adapter=ADAPTER, min_length=20, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=1, min_adapter_len=0)
        stages = [F.Stage(2, 0, 0, C.addressof(clip), None), F.Stage(0, 20, 20, None, None), F.Stage(1, 20, 90, None, None), F.Stage(3, 0, 0, None, col.h)]
    else:
        clip = F.ClipOpts(adapter=ADAPTER, min_length=15, keep_delta=0, discard_non_clipped=0, discard_clipped=0, discard_unknown=0, min_adapter_len=0)
        stages = [F.Stage(0, 25, 30, None, None), F.Stage(2, 0, 0, C.addressof(clip), None), F.Stage(3, 0, 0, None, col.h)]
    final = torch.full((n,), 12345, dtype=torch.int32, device="cuda")
    # two calls feeding one collapser would be two input streams for the clipper's history: one call, as one shell pipe
    alive = ctx.pipeline_dev(ctx.batch(torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda(), n, stride, L), 33, stages, final)
    u = col.finish(True)
    text = collapser_text(col, u)
    col.close()
    gold = json.load(open(GOLD))[name]
    assert u == gold["unique_sequences"] and alive == int((final >= 0).sum().item())
    assert hashlib.sha256(text).hexdigest() == gold["sha256"]
    assert text == oracle_pipeline(name)
    ctx.close()
