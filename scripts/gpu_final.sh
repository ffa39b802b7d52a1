#!/bin/bash
# Round-end visit: full GPU test suite, smoke, bench (own arm), launch list, BASELINE-sized runs of every kernel.
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -22 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
{
echo "# scripts/run_ops.py on one B200, BASELINE.json-sized batches, HBM-resident, CUDA events (round 1, final kernels)"
timeout 200 python scripts/run_ops.py filter 100000000 150 2>&1 | tail -1
timeout 200 python scripts/run_ops.py clip 100000000 150 2>&1 | tail -2
timeout 200 python scripts/run_ops.py stats 100000000 150 2>&1 | tail -1
timeout 200 python scripts/run_ops.py revcomp 100000000 150 2>&1 | tail -1
timeout 300 python scripts/run_ops.py collapse 200000000 50 2>&1 | tail -1
} | tee gpurun_out/full_size_$TAG.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --reads 20000000 --e2e-reads 2000000 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -1 gpurun_out/ncu_bench_$TAG.log | cut -c1-200
ls -la gpurun_out | tail -8
