#!/bin/bash
# round-2 validation on two GPUs: K-STATS (tail rotation) parity + timing, native collectives and the tools on 2 GPUs, bench --gpus 2
mkdir -p gpurun_out
exec > gpurun_out/final2.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_stats_clip.py -q -m gpu -k "stats" 2>&1 | tail -5
for L in 150 100 50 250; do timeout 300 python scripts/run_ops.py stats 60000000 $L; done
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/bench_final2_n2.json 2> gpurun_out/bench_final2_n2.err; echo bench rc=$?
tail -c 600 gpurun_out/bench_final2_n2.err
