#!/bin/bash
# compute-sanitizer passes over the small parity tests (on the GPU box): memcheck, racecheck, synccheck.
#   tools/sanitize.sh   -> gpurun_out/{memcheck,racecheck,synccheck}.log
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "c1_bundled or random_eqx_tiling or random_streaming or random_sliced or general_path or noncanonical or tiny_records or long_numbers or break_paf_random or qbed_random or largest or stats_text or multi_device" \
  > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "c1_tiling or random_eqx_tiling or break_paf_bundled or stats_text_mode" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "random_eqx_tiling or random_sliced or stats_text_mode" > gpurun_out/synccheck.log 2>&1; echo "synccheck rc=$?"
grep -hE "SUMMARY|passed|failed" gpurun_out/memcheck.log gpurun_out/racecheck.log gpurun_out/synccheck.log
