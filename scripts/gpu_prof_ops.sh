#!/bin/bash
# ncu full captures of the secondary kernels (one GPU). Usage under gpurun: bash scripts/gpu_prof_ops.sh <tag> "<ops>"
TAG=${1:-r01}
OPS=${2:-"stats clip revcomp"}
mkdir -p gpurun_out
for op in $OPS; do
  case $op in
    stats) K=k_stats; N=10000000;; clip) K=k_clip; N=2000000;; revcomp) K=k_revcomp; N=20000000;; filter) K=k_scan; N=20000000;; *) K=k_; N=10000000;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${op}_$TAG \
      python scripts/run_ops.py $op $N > gpurun_out/ncu_${op}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${op}_$TAG.log | cut -c1-200
done
