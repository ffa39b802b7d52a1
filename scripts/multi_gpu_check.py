#!/usr/bin/env python
"""Run under torchrun (one rank per GPU, NCCL): the multi-GPU forms of the three kinds of tools, checked on rank 0
against the CPU oracle on the whole input.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastx_toolkit_b200 as F  # noqa: E402
from fastx_toolkit_b200 import dist as D  # noqa: E402
import helpers as H  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = F.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n, L = int(os.environ.get("MG_N", 400000)), 50
    seq, qual = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    stride = seq.shape[1]
    lo, hi = D.shard_bounds(n, world)[rank]
    dseq, dqual = torch.from_numpy(seq[lo:hi]).cuda(), torch.from_numpy(qual[lo:hi]).cuda()
    b = ctx.batch(dseq, dqual, hi - lo, stride, L)

    # 1. map-type tool: trimmer on the shard, gathered in rank order == whole-input oracle
    out = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
    ctx.trim_dev(b, 33, 20, 20, out, lo)
    rep = ctx.sync()
    allout = D.gather_rows({"o": out})["o"].cpu().numpy()
    # 2. quality stats: all-reduce of the per-rank histograms
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.stats_accum_dev(b, 33, hist, L, None, lo)
    ctx.sync()
    D.allreduce_hist(hist)
    # 3. collapser: local dedup -> owner routing over NCCL -> owner merge -> gather -> one ordering pass
    col = F.Collapser(local, max(hi - lo, 1), stride)
    col.add(ctx.batch(dseq, None, hi - lo, stride, L), None, None, lo)
    u = col.finish(order=False)
    rows = torch.empty((u, stride), dtype=torch.uint8, device="cuda"); ln = torch.empty(u, dtype=torch.int32, device="cuda")
    cnt = torch.empty(u, dtype=torch.int64, device="cuda"); first = torch.empty(u, dtype=torch.int64, device="cuda")
    hsh = torch.empty(u, dtype=torch.int64, device="cuda")
    col.fetch(rows, ln, cnt, first, hsh)
    col.close()
    mine = D.route_to_owners({"rows": rows, "len": ln, "count": cnt, "first": first}, hsh)
    m = mine["len"].numel()
    own = F.Collapser(local, max(m, 1), stride)
    torch.cuda.synchronize()
    w32 = mine["count"].to(torch.int32)
    own.add(F.Batch(mine["rows"].data_ptr(), None, mine["len"].data_ptr(), 0, stride, m), w32, mine["first"], 0)
    u2 = own.finish(order=False)
    rows2 = torch.empty((u2, stride), dtype=torch.uint8, device="cuda"); ln2 = torch.empty(u2, dtype=torch.int32, device="cuda")
    cnt2 = torch.empty(u2, dtype=torch.int64, device="cuda"); first2 = torch.empty(u2, dtype=torch.int64, device="cuda")
    hsh2 = torch.empty(u2, dtype=torch.int64, device="cuda")
    own.fetch(rows2, ln2, cnt2, first2, hsh2)
    own.close()
    allu = D.gather_rows({"rows": rows2, "count": cnt2, "first": first2, "hash": hsh2})
    U = allu["count"].numel()
    perm = torch.empty(U, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    F.collapse_order_dev(local, allu["hash"], allu["first"], allu["count"], U, perm)
    p = perm.cpu().numpy().astype(np.int64)

    ok = True
    if rank == 0:
        exp, _ = H.o_trim(seq, qual, None, L, stride, 33, 20, 20)
        ok &= bool(np.array_equal(allout, exp))
        eh, _ = H.o_stats_hist(seq, qual, None, L, stride, 33, L)
        ok &= bool(np.array_equal(hist.cpu().numpy().astype(np.uint64), eh))
        efirst, ecnt = H.o_collapse(seq, None, L, stride)
        ok &= U == len(ecnt)
        ok &= bool(np.array_equal(allu["first"].cpu().numpy()[p], efirst))
        ok &= bool(np.array_equal(allu["count"].cpu().numpy()[p].astype(np.uint64), ecnt))
        top = allu["rows"].cpu().numpy()[p[0], :L].tobytes()
        ok &= top == seq[efirst[0], :L].tobytes()
        print("multi_gpu_check world=%d n=%d: trim %s, stats-allreduce %s, collapser U=%d %s" %
              (world, n, "OK" if np.array_equal(allout, exp) else "MISMATCH",
               "OK" if np.array_equal(hist.cpu().numpy().astype(np.uint64), eh) else "MISMATCH", U, "OK" if ok else "MISMATCH"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
