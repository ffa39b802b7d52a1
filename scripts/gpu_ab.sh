#!/bin/bash
# A/B visit used for the second-generation kernels in round 1 (parity of the touched paths, timing of variants, ncu
# captures).  Usage under gpurun: bash scripts/gpu_ab.sh <tag>
#   variants are selected by environment: FXG_STATS_V=1|2, FXG_STATS_B=0|1, FXG_CLIP_V=1|2, FXG_TUNE=ring,warps,stages,ctas
TAG=${1:-ab}
mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_stats_clip.py tests/test_extra_tools.py tests/test_barcode_splitter.py -m gpu -x -q 2>&1 | tail -8) | tee gpurun_out/pytest_ab_$TAG.log
for v in "FXG_STATS_V=1" "FXG_STATS_V=2" "FXG_STATS_V=2 FXG_STATS_B=1" "FXG_STATS_V=2 FXG_TUNE=-1,20,0,0"; do
  echo "-- $v"; env $v timeout 300 python scripts/run_ops.py stats 50000000 150 2>&1 | tail -1
done
for v in "FXG_CLIP_V=1" "FXG_CLIP_V=2"; do echo "-- $v"; env $v timeout 300 python scripts/run_ops.py clip 20000000 150 2>&1 | tail -2; done
for k in stats2 clip_dpx; do
  op=stats; n=10000000; [ $k = clip_dpx ] && { op=clip; n=4000000; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_$TAG \
      python scripts/run_ops.py $op $n > gpurun_out/ncu_${k}_$TAG.log 2>&1; tail -1 gpurun_out/ncu_${k}_$TAG.log | cut -c1-160
done
# read the captures here with:  ncu -i X.ncu-rep --page raw --csv > raw.csv ; --page source --csv > src.csv ;
#   python scripts/ncu_summary.py raw.csv ; python scripts/ncu_source_segments.py src.csv <units>
