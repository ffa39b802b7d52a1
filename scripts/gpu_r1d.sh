#!/bin/bash
# Third visit: parity of the touched paths, timing, ncu captures, e2e knob sweep.
TAG=${1:-r01d}
mkdir -p gpurun_out
nvidia-smi -L
echo "== tests (stats/clip/extra)"
(timeout 600 python -m pytest tests/test_gpu_stats_clip.py tests/test_extra_tools.py -m gpu -x -q 2>&1 | tail -25) | tee gpurun_out/pytest_statsclip_$TAG.log
echo "== perf"
for L in 150 100 50; do timeout 300 python scripts/run_ops.py stats 50000000 $L 2>&1 | tail -1; done
timeout 300 python scripts/run_ops.py clip 50000000 150 2>&1 | tail -2
timeout 300 python scripts/run_ops.py clip 20000000 100 2>&1 | tail -2
for op in validate mask artifacts; do timeout 300 python scripts/run_ops.py $op 50000000 150 2>&1 | tail -1; done
echo "== ncu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stats2 -s 1 -c 1 -f -o gpurun_out/prof_stats2_$TAG \
    python scripts/run_ops.py stats 10000000 > gpurun_out/ncu_stats2_$TAG.log 2>&1; tail -1 gpurun_out/ncu_stats2_$TAG.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_clip_dpx -s 1 -c 1 -f -o gpurun_out/prof_clipdpx_$TAG \
    python scripts/run_ops.py clip 4000000 > gpurun_out/ncu_clipdpx_$TAG.log 2>&1; tail -1 gpurun_out/ncu_clipdpx_$TAG.log | cut -c1-200
echo "== e2e sweep"
for cfg in "3 250000" "4 250000" "3 500000" "6 125000"; do
  set -- $cfg
  echo "-- workers=$1 chunk=$2"
  FXG_BENCH_WORKERS=$1 FXG_BENCH_CHUNK_READS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --reads 20000000 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'slab',round(d['e2e']['slab_level']['value'],1))"
done
ls -la gpurun_out | tail -6
