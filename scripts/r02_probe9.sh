#!/bin/bash
# round-2 ninth GPU call (1 GPU): deflate / decisions / FASTA-weighted clipper tests, ring-kernel extras, K-STATS A/B, bench N=1
mkdir -p gpurun_out
exec > gpurun_out/probe9.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_text.py -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_tools_cli.py -q -m gpu -k "gzip or numeric or output_file" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_extra_tools.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -8
for op in validate artifacts mask; do timeout 200 python scripts/run_ops.py $op 50000000 150; FXG_EXTRA_PLAIN=1 timeout 200 python scripts/run_ops.py $op 50000000 150; done
for v in "" "FXG_STATS_NOBFAST=1"; do for pair in 2 1; do
  env $v FXG_STATS_V=4 FXG_STATS_PAIR=$pair timeout 300 python scripts/run_ops.py stats 60000000 150
done; done
timeout 300 python scripts/run_ops.py stats 60000000 150
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-sample 1000000 > gpurun_out/bench_probe9.json 2> gpurun_out/bench_probe9.err; echo bench rc=$?
tail -c 1500 gpurun_out/bench_probe9.err
