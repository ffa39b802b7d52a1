#!/bin/bash
# round-2 final 8-GPU call: bench.py --gpus 8 exactly as the driver launches it (collective legs, e2e + less-D2H variants, file->file)
mkdir -p gpurun_out
exec > gpurun_out/final8.log 2>&1
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_final8_n8.json 2> gpurun_out/bench_final8_n8.err; echo bench rc=$?
tail -c 600 gpurun_out/bench_final8_n8.err
