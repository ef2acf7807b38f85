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
# round 2: BGZF inflate, trim-paf under both search policies (k_trim_rescan), the fused tokeniser + scan kernel (opt-in)
compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_trim_paf.py -m gpu -x -q \
  -k "every_block_type or random_payloads or many_blocks or corrupt or trim_reference or trim_random or trim_bundled" > gpurun_out/memcheck2.log 2>&1; echo "memcheck2 rc=$?"
RB_TOKSCAN=1 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "c1_bundled or random_eqx_tiling or noncanonical or tiny_records or long_numbers" > gpurun_out/memcheck3.log 2>&1; echo "memcheck3 (RB_TOKSCAN=1) rc=$?"
RB_TOKSCAN=1 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trim_paf.py -m gpu -x -q \
  -k "c1_tiling or trim_reference or trim_bundled" > gpurun_out/racecheck2.log 2>&1; echo "racecheck2 rc=$?"
grep -hE "SUMMARY|passed|failed" gpurun_out/memcheck.log gpurun_out/racecheck.log gpurun_out/synccheck.log gpurun_out/memcheck2.log gpurun_out/memcheck3.log gpurun_out/racecheck2.log
