#!/bin/bash
# round-2 second GPU call (1 GPU): new collapser paths (K-ORDER rewrite, fxg_dcollapse on one rank), BASELINE-size tests, bench legs
mkdir -p gpurun_out
exec > gpurun_out/probe2.log 2>&1
set -x
timeout 1500 python -m pytest tests/test_gpu_collapse.py -q -m gpu -x --durations=8 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_stats_clip.py -q -m gpu -k "config" --durations=4 2>&1 | tail -15
FXG_PIPE_STALE=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/run_ops.py collapse 200000000 50
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_probe2.json 2> gpurun_out/bench_probe2.err; echo bench rc=$?
tail -c 3000 gpurun_out/bench_probe2.err
