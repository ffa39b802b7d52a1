#!/bin/bash
# round-2: launch list of the collapser at config (e) size on one GPU (132 M uniques), then a short bench line with the legs
mkdir -p gpurun_out
exec > gpurun_out/final7.log 2>&1
set -x
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_collapse_200m.csv python scripts/run_ops.py collapse 200000000 50
timeout 150 python bench.py --steps 3 --warmup 3 --no-f2f --cpu-sample 200000 > gpurun_out/bench_final7.json 2> gpurun_out/bench_final7.err; echo bench rc=$?
tail -c 300 gpurun_out/bench_final7.err
