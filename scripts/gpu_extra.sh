#!/bin/bash
# short visit: the (f-2)/(f-4) tools and kernels
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_extra_tools.py -m gpu -x -q 2>&1 | tail -25) | tee gpurun_out/pytest_extra.log
