#!/usr/bin/env python
"""Multi-GPU forms of the three kinds of tools through the NATIVE collectives, checked on rank 0 against the CPU oracle on
the whole input.  Two launch styles:
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
        one process per GPU: fxg_comm_init_rank (id handed over by torch.distributed), fxg_comm_allreduce_u64, fxg_dcollapse_*
  python scripts/multi_gpu_check.py --single 2
        one process driving 2 GPUs (what the drop-in tools do): fxg_comm_init_all, same calls with 2 local GPUs
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastx_toolkit_b200 as F  # noqa: E402
from fastx_toolkit_b200 import dist as D  # noqa: E402
import helpers as H  # noqa: E402


def check_against_oracle(seq, qual, L, stride, allout, hist, order_first, order_count, U, rows_by_owner, perm_owner, perm_index, tag):
    exp, _ = H.o_trim(seq, qual, None, L, stride, 33, 20, 20)
    t_ok = bool(np.array_equal(allout, exp))
    eh, _ = H.o_stats_hist(seq, qual, None, L, stride, 33, L)
    s_ok = bool(np.array_equal(hist, eh))
    efirst, ecnt = H.o_collapse(seq, None, L, stride)
    c_ok = U == len(ecnt) and bool(np.array_equal(order_first, efirst)) and bool(np.array_equal(order_count, ecnt))
    # the key rows stay on their owners: the (owner, index) the order names must be the sequence the oracle prints there
    for k in (0, 1, U // 2, U - 1):
        row = rows_by_owner[perm_owner[k]][perm_index[k], :L].tobytes()
        c_ok &= row == seq[efirst[k], :L].tobytes()
    print("multi_gpu_check %s n=%d: trim %s, stats-allreduce %s, collapser U=%d %s" %
          (tag, seq.shape[0], "OK" if t_ok else "MISMATCH", "OK" if s_ok else "MISMATCH", U, "OK" if c_ok else "MISMATCH"), flush=True)
    return t_ok and s_ok and c_ok


def run_single(ngpu):
    """one process, ngpu GPUs"""
    n, L = int(os.environ.get("MG_N", 400000)), 50
    seq, qual = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    stride = seq.shape[1]
    comm = F.Comm.all(list(range(ngpu)))
    bounds = D.shard_bounds(n, ngpu)
    ctxs, batches, outs, hists, keep = [], [], [], [], []
    for g, (lo, hi) in enumerate(bounds):
        torch.cuda.set_device(g)
        ctx = F.Context(g)
        dseq, dqual = torch.from_numpy(seq[lo:hi]).cuda(g), torch.from_numpy(qual[lo:hi]).cuda(g)
        b = ctx.batch(dseq, dqual, hi - lo, stride, L)
        out = torch.empty(hi - lo, dtype=torch.int32, device="cuda:%d" % g)
        hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda:%d" % g)
        ctx.trim_dev(b, 33, 20, 20, out, lo)
        ctx.stats_accum_dev(b, 33, hist, L, None, lo)
        ctx.sync()
        ctxs.append(ctx); batches.append(ctx.batch(dseq, None, hi - lo, stride, L)); outs.append(out); hists.append(hist); keep.append((dseq, dqual))
    comm.allreduce_u64(hists, L * 5 * 109)
    dc = F.DCollapser(comm, stride)
    rep = dc.run(batches, [lo for lo, _ in bounds])
    U = rep.n_unique
    po, pi = np.empty(U, np.int32), np.empty(U, np.uint32)
    of, oc = np.empty(U, np.int64), np.empty(U, np.uint64)
    dc.fetch_order(po, pi, of, oc)
    rows = []
    for g in range(ngpu):
        r = np.zeros((U, stride), np.uint8)      # U (whole job) bounds every owner's share
        dc.fetch_local(g, r, None, None, None, None)
        rows.append(r)
    ok = all(np.array_equal(hists[0].cpu().numpy(), h.cpu().numpy()) for h in hists)
    ok &= check_against_oracle(seq, qual, L, stride, np.concatenate([o.cpu().numpy() for o in outs]), hists[0].cpu().numpy().astype(np.uint64),
                               of, oc, U, rows, po, pi, "single-process x%d" % ngpu)
    ok &= rep.first_bad_read == -1 and comm.collectives() > 0
    print("  exchange: %.1f MB over NVLink, %d NCCL groups, ms route/exchange/dedup/gather/order = %s" %
          (rep.bytes_sent / 1e6, comm.collectives(), ["%.2f" % x for x in rep.ms]), flush=True)
    dc.close(); comm.close()
    for c in ctxs:
        c.close()
    return ok


def run_rank():
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")          # plumbing only: hands the NCCL id over; every data-path collective is native
    ctx = F.Context(local)
    st = torch.cuda.current_stream()
    ctx.set_stream(st.cuda_stream)
    comm = D.native_comm(local)
    comm.set_stream(0, st.cuda_stream)
    n, L = int(os.environ.get("MG_N", 400000)), 50
    seq, qual = H.synth_slab(H.SEED_BASE + 4, n, L, H.DUPS)
    stride = seq.shape[1]
    lo, hi = D.shard_bounds(n, world)[rank]
    dseq, dqual = torch.from_numpy(seq[lo:hi]).cuda(), torch.from_numpy(qual[lo:hi]).cuda()
    b = ctx.batch(dseq, dqual, hi - lo, stride, L)
    out = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
    ctx.trim_dev(b, 33, 20, 20, out, lo)
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    ctx.stats_accum_dev(b, 33, hist, L, None, lo)
    comm.allreduce_u64([hist], L * 5 * 109)
    dc = F.DCollapser(comm, stride)
    rep = dc.run([ctx.batch(dseq, None, hi - lo, stride, L)], [lo])
    comm.sync()
    U = rep.n_unique
    # test-side plumbing (gloo): collect every rank's trim result and owned rows on rank 0 for the comparison
    mine = np.zeros((max(rep.n_unique_local, 1), stride), np.uint8)
    dc.fetch_local(0, mine, None, None, None, None)
    parts = [None] * world
    dist.gather_object((out.cpu().numpy(), mine[: rep.n_unique_local]), parts if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        po, pi = np.empty(U, np.int32), np.empty(U, np.uint32)
        of, oc = np.empty(U, np.int64), np.empty(U, np.uint64)
        dc.fetch_order(po, pi, of, oc)
        ok = check_against_oracle(seq, qual, L, stride, np.concatenate([p[0] for p in parts]), hist.cpu().numpy().astype(np.uint64), of, oc, U,
                                  [p[1] for p in parts], po, pi, "world=%d" % world)
        ok &= rep.first_bad_read == -1
        print("  exchange: %.1f MB sent by rank 0 over NVLink, %d NCCL groups, ms route/exchange/dedup/gather/order = %s" %
              (rep.bytes_sent / 1e6, comm.collectives(), ["%.2f" % x for x in rep.ms]), flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dc.close(); comm.close(); ctx.close()
    dist.destroy_process_group()
    return int(flag.item()) == 1


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--single":
        sys.exit(0 if run_single(int(sys.argv[2])) else 1)
    sys.exit(0 if run_rank() else 1)
