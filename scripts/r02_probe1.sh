#!/bin/bash
# round-2 first GPU call: run the two never-run device paths (k_stats3, stale-row scan) and record the host topology
mkdir -p gpurun_out
exec > gpurun_out/probe1.log 2>&1
set -x
nproc; lscpu | sed -n 1,40p; numactl -H; free -g; nvidia-smi topo -m; nvidia-smi -L
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q '^0x0302' $d/class; then echo $d $(cat $d/numa_node) $(cat $d/current_link_speed) $(cat $d/current_link_width); fi; done
df -h /dev/shm /tmp
python -c "import torch" 
FXG_STATS_V=3 timeout 900 python -m pytest tests/test_gpu_stats_clip.py -x -q -m gpu -k "stats" 2>&1 | tail -15
for L in 150 100 50; do
  timeout 300 python scripts/run_ops.py stats 60000000 $L
  FXG_STATS_V=3 timeout 300 python scripts/run_ops.py stats 60000000 $L
done
FXG_PIPE_STALE=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -15
