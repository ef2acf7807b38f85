#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the hot kernels.
#   tools/profile.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.ncu-rep
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/run_steps.py --steps 3 --warmup 2 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_(tokenise|scan_lift|lift|serialise)$' -s 8 -c 4 \
    -o gpurun_out/${TAG}_full -f python tools/run_steps.py --steps 2 --warmup 2 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out
