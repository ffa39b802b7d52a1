#!/bin/bash
# round-2 8-GPU call: the host<->device copy ceiling of the box, bench.py --gpus 8 (collective legs at N=8), e2e with 2/3/4 workers per GPU
mkdir -p gpurun_out
exec > gpurun_out/probe8.log 2>&1
set -x
nproc; free -g | head -2; nvidia-smi topo -m | head -12; lscpu | grep -E "Model name|Socket|NUMA|L3"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 python scripts/pcie_probe.py
timeout 200 $TR --master-port 29541 scripts/pcie_probe.py
timeout 900 $TR --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_probe8_n8.json 2> gpurun_out/bench_probe8_n8.err; echo bench rc=$?
tail -c 1500 gpurun_out/bench_probe8_n8.err
for w in 2 4; do
  FXG_BENCH_WORKERS=$w timeout 300 $TR --master-port 2955$w bench.py --gpus 8 --steps 5 --warmup 3 --e2e-only > gpurun_out/bench_probe8_e2e_w$w.json 2>/dev/null; echo e2e w=$w rc=$?
done
FXG_BENCH_WORKERS=3 FXG_BENCH_CHUNK_READS=100000 timeout 300 $TR --master-port 29557 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-only > gpurun_out/bench_probe8_e2e_c100k.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_probe8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value'],1), 'slab', round(d['e2e']['slab_level']['value'],1), d['e2e'].get('pcie_rank0'), d['e2e'].get('file_to_file'))
        for k,v in d.get('legs',{}).items(): print('   ', k, round(v['ms_per_step'],2), 'ms', round(v['value']), 'Mreads/s', v.get('phase_ms_rank0'), v.get('allreduce_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
