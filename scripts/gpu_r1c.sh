#!/bin/bash
# Second visit for the second-generation kernels: parity of the touched paths, A/B timing, ncu captures.
TAG=${1:-r01c}
mkdir -p gpurun_out
nvidia-smi -L
echo "== tests (stats/clip)"
(timeout 600 python -m pytest tests/test_gpu_stats_clip.py -m gpu -x -q 2>&1 | tail -25) | tee gpurun_out/pytest_statsclip_$TAG.log
echo "== perf: stats"
for v in "FXG_STATS_V=2" "FXG_STATS_V=2 FXG_TUNE=-1,20,0,0" "FXG_STATS_V=2 FXG_TUNE=-1,16,0,0"; do
  echo "-- $v"; env $v timeout 300 python scripts/run_ops.py stats 50000000 150 2>&1 | tail -1
done
for L in 100 50 250; do echo "-- v2 L=$L"; timeout 300 python scripts/run_ops.py stats 30000000 $L 2>&1 | tail -1; done
echo "== perf: clip"
timeout 300 python scripts/run_ops.py clip 50000000 150 2>&1 | tail -2
timeout 300 python scripts/run_ops.py clip 20000000 100 2>&1 | tail -2
echo "== ncu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stats2 -s 1 -c 1 -f -o gpurun_out/prof_stats2_$TAG \
    python scripts/run_ops.py stats 10000000 > gpurun_out/ncu_stats2_$TAG.log 2>&1; tail -1 gpurun_out/ncu_stats2_$TAG.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_clip_dpx -s 1 -c 1 -f -o gpurun_out/prof_clipdpx_$TAG \
    python scripts/run_ops.py clip 4000000 > gpurun_out/ncu_clipdpx_$TAG.log 2>&1; tail -1 gpurun_out/ncu_clipdpx_$TAG.log | cut -c1-200
ls -la gpurun_out | tail -8
