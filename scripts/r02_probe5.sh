#!/bin/bash
# round-2 fifth GPU call (1 GPU): k_stats4 with warp pairs, numeric / FASTA / collapser text path, CLI tests with the CR fixes
mkdir -p gpurun_out
exec > gpurun_out/probe5.log 2>&1
set -x
FXG_STATS_V=4 timeout 900 python -m pytest tests/test_gpu_stats_clip.py -q -m gpu -k "stats" 2>&1 | tail -8
for pair in 1 2; do for L in 150 100 50; do
  FXG_STATS_V=4 FXG_STATS_PAIR=$pair timeout 300 python scripts/run_ops.py stats 60000000 $L
done; done
timeout 900 python -m pytest tests/test_gpu_text.py -q -m gpu 2>&1 | tail -25
timeout 900 python -m pytest tests/test_tools_cli.py -q -m gpu -k "broken or clipper_fallback or numeric" 2>&1 | tail -15
FXG_STATS_V=4 bash scripts/gpu_prof_ops.sh r02b "stats"
