"""bench.py's JSON contract (one line per run; keys the driver reads): the reference arm runs here on the CPU; for the GPU arm the
committed builder lines under profiles/ are checked for the same schema."""
import glob
import json
import os
import subprocess
import sys

import pytest

import helpers as H

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"].startswith("Mreads/sec fastq_quality_trimmer") and d["unit"] == "Mreads/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u8" and "workload" in d["config"]
    if d["n_gpus"] == 1:                      # the CPU baseline is timed on rank 0 at N = 1 only
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["value"] > 0 and d["ms_per_step"] > 0


@pytest.mark.skipif(H.ref_tool("fastq_quality_trimmer") is None, reason="oracle/_ref not built")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "100000"],
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.decode().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    check_common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_committed_builder_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(H.ROOT, "profiles", "r02_bench_n*_builder*.json")))
    assert files
    for f in files:
        d = None
        for line in open(f):
            if line.lstrip().startswith("{"):
                d = json.loads(line)
        assert d is not None, f
        check_common(d)
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]), f
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-6
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]), f
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["parity_checked"] is True and {"stats", "collapse"} <= set(d["legs"]), f
