"""Multi-GPU plumbing (one process per GPU, torch.distributed): the only places where the FASTX hot path really
exchanges data.

  * trim / filter / clip / revcomp are pure maps: ranks own contiguous read blocks (`shard_bounds`), no collective.
  * fastx_quality_stats: all-reduce (SUM) of the u64 hist[cycle][5][109] partials           -> `allreduce_hist`
  * fastx_collapser: local uniques routed to owner = std::hash mod world (all-to-all), owners merge, the
    (hash, first, count, key) rows of all owners are gathered on every rank for the single ordering pass
                                                                                               -> `route_to_owners`, `gather_rows`

Everything here is device-agnostic tensor plumbing (NCCL on GPUs, gloo in the CPU tests); the kernels stay behind
the C ABI (fastx_toolkit_b200._lib).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """[lo, hi) of every rank for n reads split into contiguous, order-preserving blocks."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def allreduce_hist(hist):
    """hist: int64 tensor holding u64 counters (same shape on every rank); summed in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def owner_of(hash_i64, world):
    """owner = (uint64) hash mod world, computed on int64 storage without unsigned support."""
    hi = (hash_i64 >> 32) & 0xFFFFFFFF
    lo = hash_i64 & 0xFFFFFFFF
    return ((hi % world) * ((1 << 32) % world) + (lo % world)) % world


def _all_to_all_rows(t, send_counts, recv_counts):
    """t is sorted by destination rank; rows are exchanged with all_to_all_single."""
    inner = t.shape[1:]
    width = 1
    for d in inner:
        width *= d
    out = torch.empty((int(sum(recv_counts)),) + tuple(inner), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out.view(-1), t.contiguous().view(-1),
                           output_split_sizes=[c * width for c in recv_counts],
                           input_split_sizes=[c * width for c in send_counts])
    return out


def route_to_owners(fields, hash_i64):
    """fields: dict name -> tensor with one row per local unique; returns the same dict for the rows this rank OWNS
    (owner = hash mod world), i.e. what the peers sent here.  `hash_i64` is routed too (key 'hash')."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    fields = dict(fields)
    fields["hash"] = hash_i64
    if world == 1:
        return fields
    own = owner_of(hash_i64, world)
    order = torch.argsort(own, stable=True)
    send = torch.bincount(own, minlength=world).to(torch.int64)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    send_l, recv_l = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
    return {k: _all_to_all_rows(v[order], send_l, recv_l) for k, v in fields.items()}


def gather_rows(fields):
    """all_gather of variable-length row sets: every rank gets the concatenation in rank order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return dict(fields)
    any_t = next(iter(fields.values()))
    n = torch.tensor([any_t.shape[0]], dtype=torch.int64, device=any_t.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    out = {}
    for k, v in fields.items():
        pad = torch.zeros((m,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[: v.shape[0]] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out[k] = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    return out
