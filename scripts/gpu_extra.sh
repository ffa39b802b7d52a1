#!/bin/bash
mkdir -p gpurun_out
(timeout 32 python scripts/run_ops.py pipeline 20000000 150 2>&1 | tail -3) | tee gpurun_out/pipeline_timing.log
