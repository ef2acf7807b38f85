#!/bin/bash
# 2-GPU bench with and without binding each rank to its GPU's NUMA node; prints the topology first.
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|^CPU(s)"
for v in "" 1; do
  echo "== RB_BENCH_NO_NUMA=$v"
  RB_BENCH_NO_NUMA=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
     bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('resident', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d.get('numa'))"
done
