#!/bin/bash
mkdir -p gpurun_out
(timeout 48 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -12) | tee gpurun_out/pytest_pipeline.log
