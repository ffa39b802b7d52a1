#!/usr/bin/env python
"""bench.py — headline benchmark: Mreads/s of the fastq_quality_trimmer hot path on 150 bp reads.

  python bench.py --gpus N --steps K --warmup W            (torchrun launches N>1, one rank per GPU)
  python bench.py --impl reference ...                      (the reference CPU tool on the host cores)

A "step" is one pass of the trimmer loop body (`-t 20 -l 20 -Q33`, validation fused) over one batch of
synthetic reads that is already resident in HBM as SoA slabs (`value`, kernel K-TRIM), and — for
`e2e` — the same call through the C-ABI host entry point fxg_trim_host() with PINNED HOST slabs, the
H2D copy of the inputs and the D2H copy of the per-read result inside the timed region.
Reads shard contiguously across ranks; the trimmer has no exchange step, so there is no collective on
the data path ("scaling": "weak": every rank owns --reads reads).

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mreads/sec fastq_quality_trimmer 150bp"
UNIT = "Mreads/s"
L, STRIDE, Q, T, MINLEN = 150, 160, 33, 20, 20
ALGO_BYTES_PER_READ = 2 * L + 4          # validate seq + scan qual + 4 B result
# dram__bytes_read.sum + dram__bytes_write.sum of K-TRIM in profiles/r01_ncu_trim_full.txt (6.4826 GB per 20 M reads):
# both rows are fetched with their 10 padding bytes (stride 160) plus the 4-byte result.
NCU_TRAFFIC_BYTES_PER_READ = 324.13
SEED = 20260925 + 1


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (of fallback; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or f[0] != str(self.gpu):
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def ref_tool(name):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_sample_fastq(path, reads):
    exe = os.path.join(ROOT, "bin", "fxg_synth")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "tools"], stdout=subprocess.DEVNULL)
    subprocess.check_call([exe, "-n", str(reads), "-l", str(L), "-s", str(SEED), "-k", "plain", "-o", path])


def time_reference_cpu(sample_reads, procs, steps, warmup):
    """Run the UNMODIFIED reference fastq_quality_trimmer (oracle/_ref, gcc -O3) file -> file on
    `procs` host cores at once (the reference has no threads: one process per core, same input).
    Falls back to the oracle port when the reference binary is unavailable."""
    tool = ref_tool("fastq_quality_trimmer")
    tmp = tempfile.mkdtemp(prefix="fxg_bench_")
    try:
        if tool:
            fq = os.path.join(tmp, "sample.fq")
            make_sample_fastq(fq, sample_reads)
            def one_step():
                t0 = time.perf_counter()
                ps = [subprocess.Popen([tool, "-Q", str(Q), "-t", str(T), "-l", str(MINLEN), "-i", fq,
                                        "-o", os.path.join(tmp, "out%d.fq" % k)]) for k in range(procs)]
                rcs = [p.wait() for p in ps]
                dt = time.perf_counter() - t0
                if any(rcs):
                    raise RuntimeError("reference tool failed: %r" % rcs)
                return dt
            kind = "reference"
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers as H   # oracle port: checker/baseline only
            seq, qual = H.synth_slab(SEED, sample_reads, L)
            procs = 1
            def one_step():
                t0 = time.perf_counter()
                H.o_trim(seq, qual, None, L, STRIDE, Q, T, MINLEN)
                return time.perf_counter() - t0
            kind = "port"
        for _ in range(warmup):
            one_step()
        times = [one_step() for _ in range(steps)]
        total = sum(times)
        mreads = sample_reads * procs * steps / total / 1e6
        return mreads, kind, procs, total / steps
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    procs = host_threads()
    sample = max(100_000, min(args.ref_sample, 32_000_000 // max(procs, 1)))     # bounds the temp files to ~10 GB in + ~10 GB out per step
    val, kind, procs, sec = time_reference_cpu(sample, procs, args.steps, args.warmup)
    what = ("%d x fastq_quality_trimmer -t 20 -l 20 -Q33 (reference 0.0.14, gcc -O3) on the same %d x %d bp synthetic FASTQ, "
            "file->file, one process per host core" % (procs, sample, L)) if kind == "reference" else \
           ("oracle port fxo_trim_batch on %d x %d bp slabs, 1 thread (reference binary unavailable)" % (sample, L))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "fastq_quality_trimmer -t 20 -l 20 -Q33, %d bp synthetic reads" % L, "read_len": L,
                   "sample_reads_per_process": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind, "sample": what},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def leg_stats(args, ctx, comm, stream, dseq, dqual, n, rank, world, barrier, max_over_ranks, parity, H):
    """BASELINE config (d), this GPU's share: K-STATS over the shard + the native all-reduce of u64 hist[150][5][109]
    (fxg_comm_allreduce_u64 = ncclAllReduce on the launch stream), both inside the timed region."""
    import numpy as np
    import torch
    ns = min(args.stats_reads, n)
    words = L * 5 * 109
    hist = torch.zeros((L, 5, 109), dtype=torch.int64, device="cuda")
    b = ctx.batch(dseq, dqual, ns, STRIDE, L)

    def step():
        hist.zero_()
        ctx.stats_accum_dev(b, Q, hist, L, None, rank * ns)
        comm.allreduce_u64([hist], words)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, c0, b0 = ctx.launches(), comm.collectives(), comm.bytes_sent()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches, colls, sent = ctx.launches() - l0, comm.collectives() - c0, comm.bytes_sent() - b0
    total = int(hist.sum().item())
    ok_sum = total == world * ns * L
    # the collective alone
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        comm.allreduce_u64([hist], words)
    e1.record(stream)
    barrier()
    ms_ar = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # K-STATS alone (no collective)
    e0.record(stream)
    for _ in range(args.steps):
        ctx.stats_accum_dev(b, Q, hist, L, None, rank * ns)
    e1.record(stream)
    barrier()
    ms_k = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # parity: every rank's prefix, all-reduced natively, against the oracle over all the prefixes
    mpar = min(args.parity_reads, ns)
    hist.zero_()
    ctx.stats_accum_dev(ctx.batch(dseq, dqual, mpar, STRIDE, L), Q, hist, L, None, rank * ns)
    comm.allreduce_u64([hist], words)
    comm.sync()
    if rank == 0:
        exp = np.zeros((L, 5, 109), np.uint64)
        for r in range(world):
            ps, pq = H.synth_slab(SEED, mpar, L, H.PLAIN, first=r * n)
            eh, _ = H.o_stats_hist(ps, pq, None, L, STRIDE, Q, L)
            exp += eh
        parity["stats_allreduce"] = bool(np.array_equal(hist.cpu().numpy().astype(np.uint64), exp)) and ok_sum
    peak, _ = measured_hbm_peak()
    return {"workload": "fastx_quality_stats on %d x %d bp per GPU (BASELINE config (d): 500 M reads over 8 GPUs) + all-reduce of u64 hist[%d][5][109]" % (ns, L, L),
            "ms_per_step": ms, "value": world * ns / (ms * 1e-3) / 1e6, "unit": UNIT, "kernel_ms": ms_k, "allreduce_ms": ms_ar,
            "allreduce_bytes": words * 8, "nvlink_bytes_per_step_rank0": sent // max(args.steps, 1), "nccl_groups_per_step": colls / max(args.steps, 1),
            "comm_nranks_seen": comm.nranks, "gpu_launches": launches, "hist_sum_ok": ok_sum,
            "roofline": {"bound": "hbm", "achieved": ns * 2 * L / (ms_k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": ns * 2 * L / (ms_k * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_read": 2 * L, "kernel": "K-STATS"}}


def leg_collapse(args, ctx, comm, stream, rank, world, barrier, max_over_ranks, parity, H):
    """BASELINE config (e), this GPU's share: 200 M / 8 reads x 50 bp, ~40 % repeats.  One step = fxg_dcollapse_run: K-ROUTE ->
    exchange over NVLink (grouped ncclSend/ncclRecv) -> K-DEDUP on the owners -> gather of (hash, first, count) -> K-ORDER on rank 0."""
    import numpy as np
    import torch
    import fastx_toolkit_b200 as F
    Lc, Sc, seed = 50, 64, 20260925 + 4
    nc = args.collapse_reads
    cseq = torch.empty((nc, Sc), dtype=torch.uint8, device="cuda")
    cq = torch.empty((nc, Sc), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(cseq, cq, nc, Lc, Sc, seed, H.DUPS, Q, first_read=rank * nc, n_total=world * nc)
    ctx.sync()
    del cq
    dc = F.DCollapser(comm, Sc)
    batch = F.Batch(cseq.data_ptr(), None, None, Lc, Sc, nc)
    rep = None
    for _ in range(max(args.warmup, 3)):
        rep = dc.run([batch], [rank * nc])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, c0, b0 = dc.launches(), comm.collectives(), comm.bytes_sent()
    phases = np.zeros(5)
    e0.record(stream)
    for _ in range(args.steps):
        rep = dc.run([batch], [rank * nc])
        phases += np.array(list(rep.ms))
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches, colls, sent = dc.launches() - l0, comm.collectives() - c0, comm.bytes_sent() - b0
    U = int(rep.n_unique)
    ok = rep.first_bad_read == -1
    if rank == 0:
        oc = np.empty(U, np.uint64)
        dc.fetch_order(None, None, None, oc)
        ok = ok and int(oc.sum()) == world * nc and bool((oc[:-1] >= oc[1:]).all())
    # parity: a small job through the same object against the oracle over the whole small input
    mpar = min(args.parity_reads, nc)
    pseq, _ = H.synth_slab(seed, world * mpar, Lc, H.DUPS)
    dps = torch.from_numpy(pseq[rank * mpar:(rank + 1) * mpar]).cuda()
    torch.cuda.synchronize()
    prep = dc.run([F.Batch(dps.data_ptr(), None, None, Lc, Sc, mpar)], [rank * mpar])
    if rank == 0:
        pu = int(prep.n_unique)
        pf, pc = np.empty(pu, np.int64), np.empty(pu, np.uint64)
        dc.fetch_order(None, None, pf, pc)
        efirst, ecnt = H.o_collapse(pseq, None, Lc, Sc)
        parity["collapse_exchange"] = pu == len(ecnt) and bool(np.array_equal(pf, efirst)) and bool(np.array_equal(pc, ecnt)) and ok
    else:
        parity["collapse_exchange"] = bool(ok)
    dc.close()
    names = ["route", "exchange", "dedup", "gather", "order"]
    return {"workload": "fastx_collapser on %d x %d bp per GPU, ~40 %% repeats (BASELINE config (e): 200 M reads over 8 GPUs), global dedup by "
                        "owner = std::hash mod %d over NCCL send/recv, ordering pass on rank 0" % (nc, Lc, world),
            "ms_per_step": ms, "value": world * nc / (ms * 1e-3) / 1e6, "unit": UNIT, "keys_per_s": world * nc / (ms * 1e-3),
            "n_unique": U, "phase_ms_rank0": {k: float(v) / args.steps for k, v in zip(names, phases)},
            "nvlink_bytes_per_step_rank0": sent // max(args.steps, 1), "nccl_groups_per_step": colls / max(args.steps, 1),
            "comm_nranks_seen": comm.nranks, "gpu_launches": launches,
            "timing": "CUDA events on the communicator's (= launch) stream around K blocking fxg_dcollapse_run calls, max over ranks"}


def file_to_file(args, chunk, chunk_reads, world):
    """bin/fastq_quality_trimmer -t 20 -l 20 -Q33 file -> file on tmpfs (the input is page-cache resident by construction): the
    whole tool, process start and CUDA context creation included, with 1 GPU and with all `world` GPUs (FASTX_GPUS)."""
    tool = os.path.join(ROOT, "bin", "fastq_quality_trimmer")
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    tmp = tempfile.mkdtemp(prefix="fxg_f2f_", dir=d)
    try:
        reps = max(1, args.f2f_reads // chunk_reads)
        fin, fout = os.path.join(tmp, "in.fq"), os.path.join(tmp, "out.fq")
        with open(fin, "wb") as f:
            for _ in range(reps):
                f.write(chunk)
        res = {"reads": reps * chunk_reads, "input_bytes": reps * len(chunk), "filesystem": d,
               "command": "bin/fastq_quality_trimmer -Q33 -t 20 -l 20 -i in.fq -o out.fq (wall clock of the whole process)"}
        ref_out = None
        for g in sorted({1, world}):
            best = None
            for _ in range(2):
                env = dict(os.environ, FASTX_GPUS=str(g), FASTX_GPU="0")
                t0 = time.perf_counter()
                r = subprocess.run([tool, "-Q", str(Q), "-t", str(T), "-l", str(MINLEN), "-i", fin, "-o", fout], env=env, stderr=subprocess.PIPE)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError("file->file run failed: %s" % r.stderr.decode()[-300:])
                best = dt if best is None else min(best, dt)
            size = os.path.getsize(fout)
            if ref_out is None:
                ref_out = size
            assert size == ref_out, "the output size changes with the number of GPUs"
            res["gpus_%d" % g] = {"seconds": best, "value": reps * chunk_reads / best / 1e6, "unit": UNIT, "output_bytes": size}
        # the first chunk's worth of output must be the bytes the library path produced for that chunk
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (HBM-resident batch)")
    ap.add_argument("--e2e-reads", type=int, default=8_000_000, help="reads per GPU per e2e step (pinned host slabs)")
    ap.add_argument("--cpu-sample", type=int, default=3_000_000, help="reads in the single-core CPU baseline sample")
    ap.add_argument("--ref-sample", type=int, default=1_000_000, help="reads per process per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stats-reads", type=int, default=62_500_000, help="reads per GPU in the quality_stats leg (config (d): 500 M / 8)")
    ap.add_argument("--collapse-reads", type=int, default=25_000_000, help="reads per GPU in the collapser leg (config (e): 200 M / 8)")
    ap.add_argument("--parity-reads", type=int, default=100_000, help="reads per GPU in the oracle parity checks (outside every timed region)")
    ap.add_argument("--no-legs", action="store_true", help="skip the quality_stats / collapser legs")
    ap.add_argument("--no-f2f", action="store_true", help="skip the file -> file run of the drop-in binary")
    ap.add_argument("--e2e-only", action="store_true", help="diagnostics: small HBM batch, no legs, no file -> file, no CPU baseline — just the e2e numbers")
    ap.add_argument("--f2f-reads", type=int, default=10_000_000, help="reads in the file -> file input (tmpfs)")
    args = ap.parse_args()

    if args.e2e_only:
        args.reads = min(args.reads, 2_000_000); args.no_legs = True; args.no_f2f = True; args.no_cpu_baseline = True
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    import fastx_toolkit_b200 as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = F.Context(local_rank)
    stream = torch.cuda.Stream()          # all timed launches and the timing events share this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    n = args.reads
    free = torch.cuda.mem_get_info()[0]
    need = n * (2 * STRIDE + 4)
    if need > free * 0.9:
        n = int(free * 0.9 / (2 * STRIDE + 4)) // 4096 * 4096
    dseq = torch.empty((n, STRIDE), dtype=torch.uint8, device="cuda")
    dqual = torch.empty((n, STRIDE), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.synth_dev(dseq, dqual, n, L, STRIDE, SEED, 0, Q, first_read=rank * n, n_total=world * n)
    batch = ctx.batch(dseq, dqual, n, STRIDE, L)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- HBM-resident: K-TRIM ----------------
    for _ in range(max(args.warmup, 3)):
        ctx.trim_dev(batch, Q, T, MINLEN, out, rank * n)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the timed region long enough for the clock sampler to see it (>= ~1.5 s) without changing K's meaning
    reps = 1
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.trim_dev(batch, Q, T, MINLEN, out, rank * n)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - launches0
    rep = ctx.sync()
    if rep.first_bad_read != -1:
        raise SystemExit("bench.py: synthetic input flagged invalid at read %d" % rep.first_bad_read)
    kept = int((out >= 0).sum().item())
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    kernel_ms = ms_per_step / reps
    peak, peak_src = measured_hbm_peak()
    achieved = n * ALGO_BYTES_PER_READ / (kernel_ms * 1e-3) / 1e9

    # keep the GPU busy a little longer under the sampler if the timed region was very short
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < 1.0:
        for _ in range(5):
            ctx.trim_dev(batch, Q, T, MINLEN, out, rank * n)
        torch.cuda.synchronize()
    clocks = sampler.stop()

    # ---------------- parity of the sharded result (outside every timed region; the oracle is the checker) ----------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import helpers as H          # oracle bindings + the numpy twin of the generator
    mpar = min(args.parity_reads, n)
    pseq, pqual = H.synth_slab(SEED, mpar, L, H.PLAIN, first=rank * n)
    ptrim, _ = H.o_trim(pseq, pqual, None, L, STRIDE, Q, T, MINLEN)
    parity = {"trim": bool(np.array_equal(out[:mpar].cpu().numpy(), ptrim))}

    def all_ranks_true(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    legs = {}
    comm = None
    if not args.no_legs:
        from fastx_toolkit_b200 import dist as D
        comm = D.native_comm(local_rank)                  # fxg_comm_init_rank: the product's own NCCL communicator
        comm.set_stream(0, stream.cuda_stream)
        legs["stats"] = leg_stats(args, ctx, comm, stream, dseq, dqual, n, rank, world, barrier, max_over_ranks, parity, H)
        legs["collapse"] = leg_collapse(args, ctx, comm, stream, rank, world, barrier, max_over_ranks, parity, H)
    parity_ok = all_ranks_true(all(parity.values()))
    if not parity_ok:
        raise SystemExit("bench.py: rank %d: sharded result differs from the oracle: %r" % (rank, parity))

    # ---------------- end to end, slab level: pinned SoA slabs -> H2D -> K-TRIM -> D2H (fxg_trim_host) ----------------
    ne = min(args.e2e_reads, n)
    hseq = torch.empty((ne, STRIDE), dtype=torch.uint8).pin_memory()
    hqual = torch.empty((ne, STRIDE), dtype=torch.uint8).pin_memory()
    hout = torch.empty(ne, dtype=torch.int32).pin_memory()
    hseq.copy_(dseq[:ne]); hqual.copy_(dqual[:ne])
    torch.cuda.synchronize()
    hb = ctx.batch(hseq, hqual, ne, STRIDE, L)
    for _ in range(max(args.warmup, 3)):
        ctx.trim_host(hb, Q, T, MINLEN, hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.trim_host(hb, Q, T, MINLEN, hout)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    assert bool((hout == out[:ne].cpu()).all()), "e2e result differs from the HBM-resident result"
    e2e_slab_value = world * ne * args.steps / dt / 1e6
    del hseq, hqual

    # ---------------- end to end, text level (the call a FASTQ user makes): FASTQ text in pinned host memory ->
    # H2D -> K-LINES/K-RECS/K-PACK -> K-TRIM -> K-EMIT -> D2H -> trimmed FASTQ text in host memory (fxg_text_run_host),
    # W worker threads per GPU, each with its own context so that copies and kernels of neighbouring chunks overlap
    import ctypes as C
    import threading
    import numpy as np
    chunk_reads = int(os.environ.get("FXG_BENCH_CHUNK_READS", "250000"))     # tuning knobs for the e2e leg only
    NROT = int(os.environ.get("FXG_BENCH_ROTATE", "4"))                       # distinct input chunks per rank, sent in rotation
    exe = os.path.join(ROOT, "bin", "fxg_synth")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "tools"], stdout=subprocess.DEVNULL)
    chunks, host_ins = [], []
    for j in range(NROT):
        cj = subprocess.run([exe, "-n", str(chunk_reads), "-l", str(L), "-s", str(SEED), "-f", str((rank * NROT + j) * chunk_reads)],
                            stdout=subprocess.PIPE, check=True).stdout
        hj = torch.empty(len(cj), dtype=torch.uint8).pin_memory()
        hj.numpy()[:] = np.frombuffer(cj, np.uint8)
        chunks.append(cj); host_ins.append(hj)
    chunk = chunks[0]
    cb = max(len(c) for c in chunks)
    nchunks = max(1, args.e2e_reads // chunk_reads)
    W = int(os.environ.get("FXG_BENCH_WORKERS", "3"))
    workers = []
    for _ in range(W):
        wctx = F.Context(local_rank)
        workers.append((wctx, F.TextPipe(wctx, cb + 4096), torch.empty(cb + cb // 4 + 64, dtype=torch.uint8).pin_memory()))
    Lib = F.lib()
    stat = {"in_bytes": 0, "out_bytes": 0}
    stat_lock = threading.Lock()

    dec_len = [torch.empty(chunk_reads, dtype=torch.int32).pin_memory() for _ in range(W)]
    dec_start = [torch.empty(chunk_reads * 4, dtype=torch.int32).pin_memory() for _ in range(W)]

    def text_worker(w, first_chunk, count, mode):
        """mode 0: FASTQ text out; 1: the same text deflated on the GPU (-z); 2: per-record decisions + line table out"""
        wctx, tp, wout = workers[w]
        rep = F.TextReport()
        ib = ob = 0
        for k in range(count):
            hin = host_ins[(first_chunk + k) % NROT]
            if mode == 2:
                rc = Lib.fxg_text_decide_host(tp.h, 0, hin.data_ptr(), hin.numel(), Q, T, MINLEN, dec_len[w].data_ptr(), dec_start[w].data_ptr(), C.byref(rep))
                nout = chunk_reads * 20
            else:
                rc = Lib.fxg_text_run_host(tp.h, 0, hin.data_ptr(), hin.numel(), Q, T, MINLEN, wout.data_ptr(), C.byref(rep))
                nout = int(rep.out_bytes)
            if rc != 0 or rep.anomaly != 0 or rep.n_records != chunk_reads:
                raise RuntimeError("text path failed: rc=%d anomaly=%d" % (rc, rep.anomaly))
            ib += hin.numel(); ob += nout
        with stat_lock:
            stat["in_bytes"] += ib; stat["out_bytes"] += ob

    def text_pass(mode=0):
        per = [nchunks // W + (1 if w < nchunks % W else 0) for w in range(W)]
        starts = [sum(per[:w]) for w in range(W)]
        th = [threading.Thread(target=text_worker, args=(w, starts[w], per[w], mode)) for w in range(W)]
        [t.start() for t in th]
        [t.join() for t in th]

    def timed_passes(mode):
        for _ in range(2):
            text_pass(mode)
        barrier()
        stat["in_bytes"] = stat["out_bytes"] = 0
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            text_pass(mode)
        torch.cuda.synchronize()
        loc = time.perf_counter() - t0_
        d = max_over_ranks(loc)
        return {"value": world * nchunks * chunk_reads * args.steps / d / 1e6, "unit": UNIT,
                "h2d_GBps_this_gpu": stat["in_bytes"] / loc / 1e9, "d2h_GBps_this_gpu": stat["out_bytes"] / loc / 1e9,
                "d2h_bytes_per_read": stat["out_bytes"] / max(1, nchunks * chunk_reads * args.steps)}

    for _ in range(max(args.warmup, 3)):
        text_pass()
    # the emitted text must be what the reference writer produces for the kernel's own decisions
    first = out[: chunk_reads].cpu().numpy() if rank == 0 else None
    barrier()
    l0 = sum(int(Lib.fxg_text_launches(w[1].h)) + w[0].launches() for w in workers)
    stat["in_bytes"] = stat["out_bytes"] = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        text_pass()
    torch.cuda.synchronize()
    dt_local = time.perf_counter() - t0
    dt = max_over_ranks(dt_local)
    launches_e2e = sum(int(Lib.fxg_text_launches(w[1].h)) + w[0].launches() for w in workers) - l0
    e2e_value = world * nchunks * chunk_reads * args.steps / dt / 1e6
    e2e_h2d_step, e2e_d2h_step = stat["in_bytes"] // args.steps, stat["out_bytes"] // args.steps
    pcie = {"h2d_GBps_this_gpu": stat["in_bytes"] / dt_local / 1e9, "d2h_GBps_this_gpu": stat["out_bytes"] / dt_local / 1e9}
    if rank == 0:
        rep0 = F.TextReport()
        Lib.fxg_text_run_host(workers[0][1].h, 0, host_ins[0].data_ptr(), host_ins[0].numel(), Q, T, MINLEN, workers[0][2].data_ptr(), C.byref(rep0))
        got = workers[0][2].numpy()[: int(rep0.out_bytes)].tobytes().split(b"\n")
        src = chunk.split(b"\n")
        k = 0
        for i in range(chunk_reads):
            if first[i] >= 0:
                assert got[4 * k] == src[4 * i] and got[4 * k + 1] == src[4 * i + 1][: first[i]] and got[4 * k + 3] == src[4 * i + 3][: first[i]], "text path output differs"
                k += 1
        assert len(got) == 4 * k + 1
    # the same input with less coming back over PCIe: (1) the output deflated on the GPU (what `-z` does), (2) only the
    # per-record decisions + line table (for a writer that gathers from its own copy of the input)
    variants = {}
    for w in workers:
        w[1].set_deflate(True)
    variants["text_in_gzip_out"] = timed_passes(1)
    for w in workers:
        w[1].set_deflate(False)
    variants["text_in_decisions_out"] = timed_passes(2)
    if rank == 0:      # the decisions must be the kernel's own
        Lib.fxg_text_decide_host(workers[0][1].h, 0, host_ins[0].data_ptr(), host_ins[0].numel(), Q, T, MINLEN, dec_len[0].data_ptr(), dec_start[0].data_ptr(), C.byref(rep0))
        assert bool((dec_len[0].numpy() == first).all()), "decisions differ from the HBM-resident result"
    for w in workers:
        w[1].close(); w[0].close()
    del workers, host_ins

    # ---------------- file -> file: the drop-in binary itself (BASELINE.md regime iii) on a page-cache-resident file ----------------
    f2f = None
    if not args.no_f2f:
        del dseq, dqual, out, batch
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            f2f = file_to_file(args, chunk, chunk_reads, world)
        barrier()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": "fastq_quality_trimmer -t 20 -l 20 -Q33 (+ fused record validation) on %d x %d bp synthetic reads per GPU, "
                        "HBM-resident SoA slabs (stride %d)" % (n, L, STRIDE),
            "reads_per_gpu": n, "read_len": L, "stride": STRIDE, "kept_reads_rank0": kept,
            "cache": "inputs are %.1f GB per GPU, far larger than the 126 MB L2: no flush needed" % (2 * n * STRIDE / 1e9),
            "sharding": "contiguous read blocks per rank, no data-path collective",
            "timing": "CUDA events on the launch stream, barrier+synchronize both sides, max over ranks",
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": n * NCU_TRAFFIC_BYTES_PER_READ, "traffic_source": "ncu --set full capture of the same kernel (profiles/r01_ncu_trim_full.txt), bytes/read x reads per launch",
                     "kernel": "fxg::k_scan_w<G=1,TRIM,HAS_SEQ> (warp-private TMA ring, lane per read)", "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ,
                     "kernel_ms": kernel_ms, "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * e2e_h2d_step, "d2h_bytes_per_step": world * e2e_d2h_step,
                "pcie_rank0": pcie, "input_rotation": "%d distinct chunks of %d reads per GPU (%.0f MB, larger than the host L3) sent in rotation" % (NROT, chunk_reads, NROT * cb / 1e6),
                "less_d2h": variants,
                "file_to_file": f2f,
                "reads_per_step_per_gpu": nchunks * chunk_reads,
                "api": "fxg_text_run_host: FASTQ text in pinned host memory -> H2D -> parse/pack/K-TRIM/emit on the GPU -> D2H -> trimmed FASTQ text "
                       "in host memory; %d worker threads per GPU, chunks of %d reads" % (W, chunk_reads),
                "timing": "host wall clock around the blocking C-ABI calls, synchronize both sides, max over ranks",
                "gpu_launches": launches_e2e,
                "slab_level": {"value": e2e_slab_value, "unit": UNIT, "api": "fxg_trim_host: pinned SoA slabs -> H2D -> K-TRIM -> D2H int32 per read",
                               "h2d_bytes_per_step": world * ne * 2 * STRIDE, "d2h_bytes_per_step": world * ne * 4}},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity_checked": parity_ok,
        "parity": {"checked_against": "oracle (CPU restatement) on %d reads per rank, outside the timed regions" % mpar, "results_rank0": parity},
        "legs": legs,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, kind, procs, sec = time_reference_cpu(args.cpu_sample, 1, 1, 0)
        line["cpu_baseline"] = {
            "value": val, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": ("reference fastq_quality_trimmer 0.0.14 (gcc -O3), -t 20 -l 20 -Q33, %d x %d bp synthetic FASTQ file -> file, "
                       "1 process, %.1f s" % (args.cpu_sample, L, sec)) if kind == "reference" else
                      ("oracle port on %d x %d bp slabs, 1 thread, %.1f s" % (args.cpu_sample, L, sec)),
            "host_threads_available": host_threads(),
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
