"""CPU, world_size 2 (gloo): the multi-GPU plumbing in fastx_toolkit_b200/dist.py — sharding, the histogram
all-reduce and the collapser's owner routing / gather — driven with the CPU oracle standing in for the kernels.
The sharded result must equal the single-process oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_world(fn, world=2):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


N, L = 20000, 50


def _stats_job(rank, world):
    from fastx_toolkit_b200 import dist as D
    seq, qual = H.synth_slab(H.SEED_BASE + 3, N, L, H.WITH_N)
    lo, hi = D.shard_bounds(N, world)[rank]
    part, _ = H.o_stats_hist(seq[lo:hi], qual[lo:hi], None, L, seq.shape[1], 33, L)
    t = torch.from_numpy(part.view(np.int64).copy())
    D.allreduce_hist(t)
    return t.numpy().view(np.uint64)


def test_stats_allreduce_gloo():
    seq, qual = H.synth_slab(H.SEED_BASE + 3, N, L, H.WITH_N)
    exp, _ = H.o_stats_hist(seq, qual, None, L, seq.shape[1], 33, L)
    for got in run_world(_stats_job):
        assert np.array_equal(got, exp)


def _trim_job(rank, world):
    from fastx_toolkit_b200 import dist as D
    seq, qual = H.synth_slab(H.SEED_BASE, N + 7, L)
    lo, hi = D.shard_bounds(N + 7, world)[rank]
    out, _ = H.o_trim(seq[lo:hi], qual[lo:hi], None, L, seq.shape[1], 33, 20, 20)
    return (lo, hi, out)


def test_map_sharding_preserves_order():
    seq, qual = H.synth_slab(H.SEED_BASE, N + 7, L)
    exp, _ = H.o_trim(seq, qual, None, L, seq.shape[1], 33, 20, 20)
    parts = sorted(run_world(_trim_job, 3))
    assert parts[0][0] == 0 and parts[-1][1] == N + 7 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert np.array_equal(np.concatenate([p[2] for p in parts]), exp)


def _local_uniques(seq, lo, hi, Lr):
    """stand-in for the per-GPU dedup table: uniques of one shard with count, first (global) index and std::hash"""
    import ctypes as C
    O = H.oracle()
    d = {}
    for i in range(lo, hi):
        k = seq[i, :Lr].tobytes()
        if k in d:
            d[k][0] += 1
        else:
            d[k] = [1, i]
    keys = list(d.keys())
    rows = np.zeros((len(keys), seq.shape[1]), np.uint8)
    cnt = np.zeros(len(keys), np.int64); first = np.zeros(len(keys), np.int64); hsh = np.zeros(len(keys), np.uint64)
    for j, k in enumerate(keys):
        rows[j, :Lr] = np.frombuffer(k, np.uint8)
        cnt[j], first[j] = d[k]
        buf = C.create_string_buffer(k, len(k) + 1)
        hsh[j] = O.fxo_hash_bytes(buf, len(k), 0xc70f6907)
    return rows, cnt, first, hsh


def _collapse_job(rank, world):
    from fastx_toolkit_b200 import dist as D
    seq, _ = H.synth_slab(H.SEED_BASE + 4, N, L, H.DUPS)
    lo, hi = D.shard_bounds(N, world)[rank]
    rows, cnt, first, hsh = _local_uniques(seq, lo, hi, L)
    mine = D.route_to_owners({"rows": torch.from_numpy(rows), "count": torch.from_numpy(cnt), "first": torch.from_numpy(first)},
                             torch.from_numpy(hsh.view(np.int64)))
    # owner-side merge (stand-in for fxg_collapse_add with weights + first indices)
    assert bool((D.owner_of(mine["hash"], world) == rank).all())
    merged = {}
    for r, c, f, h in zip(mine["rows"].numpy(), mine["count"].tolist(), mine["first"].tolist(), mine["hash"].tolist()):
        k = r[:L].tobytes()
        if k in merged:
            merged[k][0] += c; merged[k][1] = min(merged[k][1], f)
        else:
            merged[k] = [c, f, h]
    keys = list(merged.keys())
    allf = D.gather_rows({
        "rows": torch.from_numpy(np.array([np.frombuffer(k, np.uint8) for k in keys], np.uint8).reshape(len(keys), L)),
        "count": torch.tensor([merged[k][0] for k in keys], dtype=torch.int64),
        "first": torch.tensor([merged[k][1] for k in keys], dtype=torch.int64),
        "hash": torch.tensor([merged[k][2] for k in keys], dtype=torch.int64)})
    return {k: v.numpy() for k, v in allf.items()}


@pytest.mark.parametrize("world", [2, 3])
def test_collapser_owner_routing_gloo(world):
    seq, _ = H.synth_slab(H.SEED_BASE + 4, N, L, H.DUPS)
    efirst, ecnt = H.o_collapse(seq, None, L, seq.shape[1])
    for got in run_world(_collapse_job, world):
        # ordering pass stand-in: replay first occurrences (with their total counts) through the oracle's map model
        order = np.argsort(got["first"], kind="stable")
        O = H.oracle()
        c = O.fxo_collapser_new()
        for j in order:
            row = np.ascontiguousarray(got["rows"][j])
            O.fxo_collapser_add(c, row.ctypes.data_as(H.u8p), L, int(got["count"][j]))
        u = O.fxo_collapser_unique(c)
        f = np.empty(u, np.int64); cn = np.empty(u, np.uint64)
        O.fxo_collapser_order(c, H._p(f, H.i64p), H._p(cn, H.u64p))
        O.fxo_collapser_free(c)
        assert u == len(ecnt)
        assert np.array_equal(got["first"][order][f], efirst) and np.array_equal(cn, ecnt)
