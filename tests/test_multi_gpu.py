"""-m gpu, needs >= 2 GPUs (skipped otherwise): torchrun + NCCL run of scripts/multi_gpu_check.py."""
import os
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_two_rank_nccl_stats_allreduce_and_collapser_exchange():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MG_N="300000")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(H.ROOT, "scripts", "multi_gpu_check.py")],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"collapser U=" in r.stdout and b"MISMATCH" not in r.stdout


def test_single_process_two_gpus_native_collectives():
    """one process driving 2 GPUs (the drop-in tools' style): fxg_comm_init_all + fxg_comm_allreduce_u64 + fxg_dcollapse_*"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MG_N="300000")
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "scripts", "multi_gpu_check.py"), "--single", "2"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"collapser U=" in r.stdout and b"MISMATCH" not in r.stdout


def test_native_nccl_allreduce_and_multi_gpu_stats_tool(tmp_path):
    """fxg_comm_allreduce_u64 (NCCL from C, one process driving 2 GPUs) and fastx_quality_stats with FASTX_GPUS=2."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import ctypes as C
    import numpy as np
    import fastx_toolkit_b200 as F
    L = F.lib()
    devs = (C.c_int * 2)(0, 1)
    comm = C.c_void_p()
    assert L.fxg_comm_init_all(2, devs, C.byref(comm)) == 0, L.fxg_comm_error(None)
    a = torch.arange(0, 1000, dtype=torch.int64, device="cuda:0")
    b = torch.arange(0, 1000, dtype=torch.int64, device="cuda:1") * 3
    bufs = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert L.fxg_comm_allreduce_u64(comm, bufs, 1000) == 0, L.fxg_comm_error(comm)
    exp = np.arange(1000, dtype=np.int64) * 4
    assert np.array_equal(a.cpu().numpy(), exp) and np.array_equal(b.cpu().numpy(), exp)
    L.fxg_comm_free(comm)

    if H.ref_tool("fastx_quality_stats") is None:
        pytest.skip("oracle/_ref not built")
    from test_tools_cli import BIN, run_tool
    fq = str(tmp_path / "in.fq")
    seq, qual = H.synth_slab(H.SEED_BASE + 17, 600000, 150, H.WITH_N)       # > 64 MB of text: several chunks
    H.write_fastq(fq, seq, qual, None, 150)
    ref = run_tool(H.ref_tool("fastx_quality_stats"), ["-N", "-i", fq])
    env = dict(os.environ, FASTX_GPUS="2")
    r = subprocess.run([os.path.join(BIN, "fastx_quality_stats"), "-N", "-i", fq], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert (r.returncode, r.stdout) == (ref[0], ref[1]), r.stderr.decode()[-500:]
    env["FASTX_TEXT_PATH"] = "0"
    env["FASTX_BATCH_READS"] = "100000"
    r = subprocess.run([os.path.join(BIN, "fastx_quality_stats"), "-i", fq], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    ref = run_tool(H.ref_tool("fastx_quality_stats"), ["-i", fq])
    assert (r.returncode, r.stdout) == (ref[0], ref[1]), r.stderr.decode()[-500:]


def test_drop_in_tools_on_two_gpus(tmp_path):
    """FASTX_GPUS=2: the streaming engine spreads the text chunks over both GPUs (in-order writer); the collapser merges the two
    partial count maps through the native owner exchange.  Output, report and exit status must be the reference's."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if H.ref_tool("fastx_collapser") is None:
        pytest.skip("oracle/_ref not built")
    from test_tools_cli import assert_same
    fq = str(tmp_path / "in.fq")
    seq, qual = H.synth_slab(H.SEED_BASE + 4, 400000, 50, H.DUPS)
    H.write_fastq(fq, seq, qual, None, 50)
    fa = str(tmp_path / "in.fa")
    H.write_fasta(fa, seq[:150000], None, 50, prefix="7-")
    env = dict(FASTX_GPUS="2", FASTX_GPUS_FORCE="1", FASTX_CHUNK_BYTES="2000000")
    os.environ.update(env)
    try:
        assert_same("fastq_quality_trimmer", ["-t", "20", "-l", "20", "-v", "-i", fq])
        assert_same("fastq_quality_filter", ["-q", "20", "-p", "80", "-v", "-i", fq])
        assert_same("fastx_reverse_complement", ["-v", "-i", fq])
        assert_same("fastx_clipper", ["-a", "AGATCGGAAGAGC", "-l", "10", "-v", "-i", fq])
        assert_same("fastx_collapser", ["-v", "-i", fq])
        assert_same("fastx_collapser", ["-v", "-i", fa])
        assert_same("fastx_quality_stats", ["-i", fq])
    finally:
        for k in env:
            os.environ.pop(k, None)
