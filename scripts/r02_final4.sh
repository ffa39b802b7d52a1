#!/bin/bash
# round-2 last sanity (1 GPU): the tools whose engine path changed last (ordered apply for stats / collapser), then a short bench
mkdir -p gpurun_out
exec > gpurun_out/final4.log 2>&1
set -x
timeout 900 python -m pytest tests/test_tools_cli.py -q -m gpu -k "trim_filter_revcomp_stats or collapser or numeric or small_windows" 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 300000 --f2f-reads 2000000 > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo bench rc=$?
tail -c 300 gpurun_out/bench_final4.err
