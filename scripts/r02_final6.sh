#!/bin/bash
# round-2 last GPU call: K-ORDER with the partitioned first-touch pass + block-aggregated compaction — parity, then timings
mkdir -p gpurun_out
exec > gpurun_out/final6.log 2>&1
set -x
timeout 200 python -m pytest tests/test_gpu_collapse.py -q -m gpu -k "not 200m" 2>&1 | tail -6
timeout 100 python scripts/run_ops.py collapse 200000000 50
timeout 60 python scripts/run_ops.py collapse 25000000 50
timeout 160 python -m pytest tests/test_gpu_collapse.py -q -m gpu -k "200m" 2>&1 | tail -4
