#!/usr/bin/env python
"""Where the end-to-end time of rb_liftover() goes on the bench workload: upload / kernels / download, wall clock."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rustybam_b200 import capi, hostlib

paf = hostlib.HostPaf.synth(scale=1.0)
wins = paf.tiling_windows(1000)
lib = capi.load()
for ptr, nbytes in ((paf.c.cigar, paf.cigar_nbytes), (wins.c.st, wins.n_win * 8), (wins.c.en, wins.n_win * 8),
                    (wins.c.bed_row, wins.n_win * 4), (wins.c.t_id, wins.n_win * 4)):
    lib.rb_host_register(C.c_void_p(C.cast(ptr, C.c_void_p).value), nbytes)
ctx = capi.Context(0)


def t(f, n=5):
    f(); f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sum(ts) / len(ts), r


b = [None]


def up():
    if b[0]:
        ctx.batch_free(b[0])
    b[0] = ctx.upload(paf, wins)


print("upload (alloc+free each time) ms min/avg", t(up)[:2])
print("kernels ms", t(lambda: ctx.batch_liftover(b[0], with_stats=True, want=capi.WANT_TEXT))[:2])
print("download ms", t(lambda: ctx.batch_download_lift(b[0], want=capi.WANT_TEXT, stats=True, copy=False))[:2])
print("rb_liftover ms", t(lambda: ctx.liftover(paf, wins, want=capi.WANT_TEXT, stats=True, copy=False))[:2])
os.environ["RB_TRACE"] = "1"
ctx.liftover(paf, wins, want=capi.WANT_TEXT, stats=True, copy=False)
del os.environ["RB_TRACE"]
# raw PCIe: pinned copies of the same sizes with torch
h = torch.empty(655_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty(655_000_000, dtype=torch.uint8, device="cuda")
print("torch D2H 655 MB ms", t(lambda: h.copy_(d, non_blocking=True))[:2])
print("torch H2D 196 MB ms", t(lambda: d[:196_000_000].copy_(h[:196_000_000], non_blocking=True))[:2])
