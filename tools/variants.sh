#!/bin/bash
# Build tuning variants of librbcuda.so HERE (no GPU needed): tools/variants.sh name "-DX=1 -DY=2" [name2 "..."] ...
# -> rustybam_b200/variants/librbcuda_<name>.so ; on the GPU box: RBCUDA_LIB=$PWD/rustybam_b200/variants/librbcuda_<name>.so tools/quick_bench.sh <tag>
set -e
cd "$(dirname "$0")/.."
mkdir -p rustybam_b200/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $defs -shared \
       -o rustybam_b200/variants/librbcuda_${name}.so rustybam_b200/csrc/*.cu 2> rustybam_b200/variants/${name}_ptxas.log &
done
wait
grep -A2 "k_emit" rustybam_b200/variants/*_ptxas.log | grep -E "spill|registers"
