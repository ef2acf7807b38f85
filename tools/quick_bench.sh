#!/bin/bash
# quick resident / e2e timing of the C4 workload on the GPU box: tools/quick_bench.sh <tag>
TAG=${1:-q}
python bench.py --only-c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<P
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("resident ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
print({k: round(v, 4) for k, v in d["roofline"]["all_kernels_ms"].items()})
P
tail -3 gpurun_out/${TAG}_bench.err
