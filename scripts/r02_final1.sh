#!/bin/bash
# round-2 validation on one GPU: the whole GPU suite, K-STATS timing + ncu capture, smoke, bench N=1 + launch list
mkdir -p gpurun_out
exec > gpurun_out/final1.log 2>&1
set -x
timeout 2400 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -40
for L in 150 100 50; do timeout 300 python scripts/run_ops.py stats 60000000 $L; done
timeout 300 python scripts/run_ops.py collapse 200000000 50
bash scripts/gpu_prof_ops.sh r02c "stats"
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_final1.json 2> gpurun_out/bench_final1.err; echo bench rc=$?
tail -c 800 gpurun_out/bench_final1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-f2f > gpurun_out/bench_under_ncu.log 2>&1; echo ncu rc=$?
