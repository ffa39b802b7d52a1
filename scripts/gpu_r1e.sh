#!/bin/bash
TAG=${1:-r01e}
mkdir -p gpurun_out
echo "== tests"
(timeout 300 python -m pytest tests/test_gpu_stats_clip.py -m gpu -x -q -k "stats" 2>&1 | tail -5) | tee gpurun_out/pytest_stats_$TAG.log
(FXG_STATS_B=0 timeout 300 python -m pytest tests/test_gpu_stats_clip.py -m gpu -x -q -k "stats2 or generations" 2>&1 | tail -3)
echo "== perf"
for v in "FXG_STATS_B=0" "FXG_STATS_B=1" "FXG_STATS_B=1 FXG_TUNE=-1,20,0,0"; do
  echo "-- $v"; env $v timeout 200 python scripts/run_ops.py stats 50000000 150 2>&1 | tail -1
done
for L in 100 50 250; do echo "-- B=1 L=$L"; timeout 200 python scripts/run_ops.py stats 30000000 $L 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stats2 -s 1 -c 1 -f -o gpurun_out/prof_stats2_$TAG \
    python scripts/run_ops.py stats 10000000 > gpurun_out/ncu_stats2_$TAG.log 2>&1; tail -1 gpurun_out/ncu_stats2_$TAG.log | cut -c1-160
