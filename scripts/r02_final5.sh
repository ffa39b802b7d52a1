#!/bin/bash
# round-2: survivor counts of fused pipelines kept on the device — parity of every kernel that learned n_dev, then timings
mkdir -p gpurun_out
exec > gpurun_out/final5.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py tests/test_gpu_stats_clip.py -q -m gpu 2>&1 | tail -8
timeout 200 python scripts/run_ops.py pipeline 20000000 100
timeout 200 python scripts/run_ops.py pipeline 2000000 100
timeout 200 python scripts/run_ops.py trim 100000000 50
timeout 300 python bench.py --steps 5 --warmup 3 --no-legs --no-f2f --cpu-sample 200000 > gpurun_out/bench_final5.json 2> gpurun_out/bench_final5.err; echo bench rc=$?
tail -c 300 gpurun_out/bench_final5.err
