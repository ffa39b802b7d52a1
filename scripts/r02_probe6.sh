#!/bin/bash
# round-2 sixth GPU call (1 GPU): k_stats4 B-fast block, text tests, the tools on the streaming engine (CLI suite), file->file throughput
mkdir -p gpurun_out
exec > gpurun_out/probe6.log 2>&1
set -x
FXG_STATS_V=4 timeout 900 python -m pytest tests/test_gpu_stats_clip.py -q -m gpu -k "stats" 2>&1 | tail -5
for pair in 2 1; do for L in 150 100 50; do
  FXG_STATS_V=4 FXG_STATS_PAIR=$pair timeout 300 python scripts/run_ops.py stats 60000000 $L
done; done
timeout 900 python -m pytest tests/test_gpu_text.py -q -m gpu 2>&1 | tail -12
timeout 1500 python -m pytest tests/test_tools_cli.py tests/test_extra_tools.py -q -m gpu 2>&1 | tail -25
# file -> file throughput of the trimmer on tmpfs (page-cache-resident by construction)
bin/fxg_synth -n 20000000 -l 150 -s 20260926 -k plain -o /dev/shm/in20m.fq
ls -la /dev/shm/in20m.fq
for w in 3 2; do for rt in 4 1; do
  FASTX_TIMING=1 FASTX_WORKERS=$w FASTX_READ_THREADS=$rt bash -c 'time bin/fastq_quality_trimmer -Q33 -t 20 -l 20 -i /dev/shm/in20m.fq -o /dev/shm/out20m.fq' 2>&1 | tail -8
done; done
FASTX_TIMING=1 bash -c 'time bin/fastq_quality_trimmer -Q33 -t 20 -l 20 -i /dev/shm/in20m.fq -o /dev/null' 2>&1 | tail -6
FASTX_TIMING=1 bash -c 'time bin/fastq_quality_filter -Q33 -q 20 -p 90 -i /dev/shm/in20m.fq -o /dev/shm/out20m.fq' 2>&1 | tail -6
FASTX_TIMING=1 bash -c 'time bin/fastx_quality_stats -i /dev/shm/in20m.fq -o /dev/shm/stats.txt' 2>&1 | tail -6
head -c 1500000000 /dev/shm/in20m.fq > /dev/null
bash -c 'time oracle/_ref/fastq_quality_trimmer -Q33 -t 20 -l 20 -i /dev/shm/in20m.fq -o /dev/shm/ref20m.fq' 2>&1 | tail -4
cmp /dev/shm/out20m.fq /dev/shm/ref20m.fq; echo "cmp(filter out vs trimmer ref: expected to differ) rc=$?"
bin/fastq_quality_trimmer -Q33 -t 20 -l 20 -i /dev/shm/in20m.fq -o /dev/shm/out20m.fq; cmp /dev/shm/out20m.fq /dev/shm/ref20m.fq; echo "trimmer cmp rc=$?"
rm -f /dev/shm/*.fq /dev/shm/stats.txt
