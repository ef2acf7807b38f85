mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r01z_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r01z_gpu_suite.log
timeout 90 python tools/trim_time.py gpurun_out/r01z_trim_time.jsonl > gpurun_out/r01z_trim_time.log 2>&1; echo "trim_time rc=$?"; cut -c1-1200 gpurun_out/r01z_trim_time.log | tail -3
