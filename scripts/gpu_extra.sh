#!/bin/bash
# short visit: sanity of the tool binaries that sit on the kernels changed last (quality stats, clipper) + smoke
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_tools_cli.py tests/test_gpu_text.py -m gpu -x -q -k "stats_binaries or clipper_binaries or golden_fixtures or text" 2>&1 | tail -8) | tee gpurun_out/pytest_final_subset.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
