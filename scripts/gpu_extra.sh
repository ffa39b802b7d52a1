#!/bin/bash
# short visit: the (f-2)/(f-4) tools and kernels
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_barcode_splitter.py -m gpu -x -q 2>&1 | tail -30) | tee gpurun_out/pytest_barcode.log
