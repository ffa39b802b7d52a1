#!/bin/bash
# round-2 third GPU call (2 GPUs): k_stats4 parity + timing, pipelines with the collapser, native collectives on 2 GPUs, bench --gpus 2
mkdir -p gpurun_out
exec > gpurun_out/probe3.log 2>&1
set -x
nvidia-smi -L; nproc
FXG_STATS_V=4 timeout 900 python -m pytest tests/test_gpu_stats_clip.py -q -m gpu -k "stats" 2>&1 | tail -15
for L in 150 100 50; do
  FXG_STATS_V=4 timeout 300 python scripts/run_ops.py stats 60000000 $L
done
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x 2>&1 | tail -25
MG_N=400000 timeout 300 python scripts/multi_gpu_check.py --single 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_probe3_n2.json 2> gpurun_out/bench_probe3_n2.err; echo bench rc=$?
tail -c 2000 gpurun_out/bench_probe3_n2.err
