mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_trim_paf.py -q --tb=short > gpurun_out/r01z_trim_tests.log 2>&1; echo "trim tests rc=$?"; tail -25 gpurun_out/r01z_trim_tests.log
timeout 200 python -m pytest tests/test_gpu_scale.py -m gpu -q --tb=short -k "trim" > gpurun_out/r01z_trim_scale.log 2>&1; echo "trim scale rc=$?"; tail -25 gpurun_out/r01z_trim_scale.log
timeout 90 python tools/trim_time.py gpurun_out/r01z_trim_time.jsonl > gpurun_out/r01z_trim_time.log 2>&1; echo "trim_time rc=$?"; tail -5 gpurun_out/r01z_trim_time.log
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_trim_paf.py -m gpu -x -q -k "reference_vectors or many_names or errors" > gpurun_out/r01z_trim_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -hE "ERROR SUMMARY|passed|failed" gpurun_out/r01z_trim_memcheck.log
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_trim_paf.py -m gpu -x -q -k "reference_vectors or many_names" > gpurun_out/r01z_trim_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -hE "RACECHECK SUMMARY|passed|failed" gpurun_out/r01z_trim_racecheck.log
if [ "$1" = "bench" ]; then timeout 300 python bench.py > gpurun_out/r01z_bench.json 2> gpurun_out/r01z_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r01z_bench.json; fi
