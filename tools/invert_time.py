#!/usr/bin/env python
"""rb invert on the bench PAF (C4's records): end-to-end call times and the per-kernel split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustybam_b200 import capi, hostlib

ctx = capi.Context(0)
full = hostlib.HostPaf.synth(scale=1.0)
for i in range(4):
    t = time.perf_counter()
    r = ctx.invert(full, want=capi.WANT_TEXT, copy=False)
    print("invert e2e ms %.2f" % ((time.perf_counter() - t) * 1e3), r)
ctx.set_profiling(True)
ctx.invert(full, want=capi.WANT_TEXT, copy=False)
print({k: (n, round(ms, 3)) for k, (n, ms) in ctx.kernel_times().items()})
