#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, then the ncu launch list and one full capture of K-TRIM.
# Usage (under gpurun): bash scripts/gpu_check.sh [tag] [skip_ncu]
TAG=${1:-r01}
SKIP_NCU=${2:-0}
mkdir -p gpurun_out
nvidia-smi -L
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_gpu_$TAG.log 2>&1
cat gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -5 gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
if [ "$SKIP_NCU" = "0" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 3 --warmup 3 --reads 20000000 --e2e-reads 2000000 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 2 -f -o gpurun_out/prof_trim_$TAG \
      python bench.py --steps 3 --warmup 3 --reads 20000000 --e2e-reads 2000000 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
  ls -la gpurun_out/
fi
