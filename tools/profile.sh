#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the hot kernels (C4 workload, resident steps).
#   tools/profile.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.ncu-rep
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/run_steps.py --steps 3 --warmup 2 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(k_tokenise|k_samples2|k_lift|k_emit|k_rec_prep|k_lift_plan)' -s 12 -c 7 \
    -o gpurun_out/${TAG}_full -f python tools/run_steps.py --steps 2 --warmup 2 > gpurun_out/${TAG}_full.log 2>&1
# the same for the headline's window width (10 kb: no block fits k_emit's staging area, k_lift lifts everything)
ncu --set full --clock-control none --import-source on -k regex:'^(k_tokenise|k_samples2|k_lift|k_emit)' -s 8 -c 4 \
    -o gpurun_out/${TAG}_w10k_full -f python tools/run_steps.py --steps 2 --warmup 2 --window 10000 > gpurun_out/${TAG}_w10k_full.log 2>&1
ls -la gpurun_out | tail
