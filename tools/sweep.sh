#!/bin/bash
# Tuning sweep on the GPU box: rebuild librbcuda.so with each set of -D knobs and print the per-kernel times.
#   tools/sweep.sh "<defs 1>" "<defs 2>" ...      (results: gpurun_out/sweep.txt)
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for defs in "$@"; do
  echo "=== $defs" | tee -a gpurun_out/sweep.txt
  RB_NVCC_DEFS="$defs" python -m rustybam_b200.build --force > gpurun_out/sweep_build.log 2>&1 || { echo BUILD FAILED | tee -a gpurun_out/sweep.txt; tail -5 gpurun_out/sweep_build.log; continue; }
  grep -E "k_liftE|k_scan_liftILb0" -A2 gpurun_out/sweep_build.log | grep -E "spill|Used" | tr '\n' ' ' >> gpurun_out/sweep.txt; echo >> gpurun_out/sweep.txt
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('step_ms', round(d['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],2), {k:round(v,3) for k,v in d['roofline']['all_kernels_ms'].items()})" | tee -a gpurun_out/sweep.txt
done
python -m rustybam_b200.build --force > /dev/null 2>&1
