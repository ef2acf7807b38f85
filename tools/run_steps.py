#!/usr/bin/env python
"""Runs W + K resident liftover(+stats) steps of the bench workload (C4, 1 GPU) and nothing else —
the target command for `ncu` launch lists and full captures (see tools/profile.sh)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--window", type=int, default=1000)
ap.add_argument("--stats-only", action="store_true")
args = ap.parse_args()

from rustybam_b200 import capi, hostlib

paf = hostlib.HostPaf.synth(scale=args.scale)
wins = paf.tiling_windows(args.window)
ctx = capi.Context(0)
b = ctx.upload(paf, None if args.stats_only else wins)
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for i in range(args.warmup + args.steps):
    if os.environ.get('RB_FLUSH'):
        flush.fill_(1)
    s = ctx.batch_stats(b) if args.stats_only else ctx.batch_liftover(b, with_stats=True, want=capi.WANT_TEXT)
print(s)
ctx.batch_free(b)
ctx.close()
