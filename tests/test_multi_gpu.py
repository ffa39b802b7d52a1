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
