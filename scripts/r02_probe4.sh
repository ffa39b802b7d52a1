#!/bin/bash
# round-2 fourth GPU call (1 GPU): ncu capture of k_stats4, tool CLI tests with the advisor fixes, K-ORDER steady state + launch list
mkdir -p gpurun_out
exec > gpurun_out/probe4.log 2>&1
set -x
FXG_STATS_V=4 bash scripts/gpu_prof_ops.sh r02a "stats"
timeout 900 python -m pytest tests/test_tools_cli.py tests/test_gpu_text.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python scripts/run_ops.py collapse 200000000 50
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_collapse_r02a.csv python scripts/run_ops.py collapse 50000000 50 > gpurun_out/ncu_collapse_r02a.log 2>&1
tail -3 gpurun_out/ncu_collapse_r02a.log
