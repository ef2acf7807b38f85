#!/usr/bin/env python
"""The box's host<->device copy ceilings, the denominator of the end-to-end numbers (VERDICT r1 #5a):
H2D alone, D2H alone and both at once (pinned memory, 1 GiB copies, CUDA events), on 1, 2, 4, ... GPUs CONCURRENTLY
(one host thread per GPU).  Prints one JSON object.
    python tools/pcie_ceiling.py [--gpus 1,2,4,8] > profiles/rNN_pcie_ceiling.json"""
import argparse
import json
import threading

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default=None)
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
n_all = torch.cuda.device_count()
sets = [int(x) for x in args.gpus.split(",")] if args.gpus else [n for n in (1, 2, 4, 8) if n <= n_all]
N = args.mib << 20
bufs = {}
for d in range(max(sets)):
    torch.cuda.set_device(d)
    bufs[d] = (torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory(),
               torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}"),
               torch.cuda.Stream(device=d), torch.cuda.Stream(device=d))


def run(d, mode, out, barrier):
    torch.cuda.set_device(d)
    h_in, h_out, d_in, d_out, sa, sb = bufs[d]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(d)
    barrier.wait()
    e0.record(torch.cuda.current_stream(d))
    sa.wait_event(e0)
    sb.wait_event(e0)
    for _ in range(args.reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(sa):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(sb):
                h_out.copy_(d_out, non_blocking=True)
    cur = torch.cuda.current_stream(d)
    cur.wait_stream(sa)
    cur.wait_stream(sb)
    e1.record(cur)
    e1.synchronize()
    out[d] = e0.elapsed_time(e1)


res = {"copy_mib": args.mib, "reps": args.reps, "gpu": torch.cuda.get_device_name(0), "sets": {}}
for n in sets:
    row = {}
    for mode in ("h2d", "d2h", "both"):
        best = None
        for _ in range(3):
            out, barrier = {}, threading.Barrier(n)
            th = [threading.Thread(target=run, args=(d, mode, out, barrier)) for d in range(n)]
            [t.start() for t in th]
            [t.join() for t in th]
            ms = max(out.values())
            best = ms if best is None else min(best, ms)
        nbytes = N * args.reps * n * (2 if mode == "both" else 1)
        row[mode] = {"ms": best, "aggregate_gb_per_s": nbytes / (best * 1e-3) / 1e9, "per_gpu_gb_per_s": nbytes / n / (best * 1e-3) / 1e9}
    res["sets"][str(n)] = row
print(json.dumps(res))
