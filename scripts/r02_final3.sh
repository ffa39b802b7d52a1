#!/bin/bash
# round-2 last checks (2 GPUs): tools on two GPUs, collapser growth, K-ARTIFACT table counting (parity + timing)
mkdir -p gpurun_out
exec > gpurun_out/final3.log 2>&1
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_collapse.py -q -m gpu -k "grows or dcollapse or epoch" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_extra_tools.py -q -m gpu 2>&1 | tail -6
timeout 200 python scripts/run_ops.py artifacts 50000000 150
timeout 200 python scripts/run_ops.py validate 50000000 150
