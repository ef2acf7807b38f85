#!/usr/bin/env python
"""Does a host->device copy proceed while a device->host copy is in flight (separate DMA engines)?"""
import time
import torch

h_out = torch.empty(655_000_000, dtype=torch.uint8).pin_memory()
d_out = torch.empty(655_000_000, dtype=torch.uint8, device="cuda")
h_in = torch.empty(20_000_000, dtype=torch.uint8).pin_memory()
d_in = torch.empty(20_000_000, dtype=torch.uint8, device="cuda")
A, B = torch.cuda.Stream(), torch.cuda.Stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


for trial in range(3):
    torch.cuda.synchronize()
    a0, a1, b0, b1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(B):
        b0.record()
        h_out.copy_(d_out, non_blocking=True)
        b1.record()
    time.sleep(0.002)  # the D2H is running now
    with torch.cuda.stream(A):
        a0.record()
        d_in.copy_(h_in, non_blocking=True)
        a1.record()
    torch.cuda.synchronize()
    print(f"D2H 655 MB {b0.elapsed_time(b1):.2f} ms ({0.655 / b0.elapsed_time(b1) * 1e3:.1f} GB/s) | H2D 20 MB issued 2 ms later took {a0.elapsed_time(a1):.2f} ms, "
          f"finished {b0.elapsed_time(a1):.2f} ms after the D2H started")
    # kernel on A while the D2H runs
    torch.cuda.synchronize()
    with torch.cuda.stream(B):
        b0.record()
        h_out.copy_(d_out, non_blocking=True)
        b1.record()
    time.sleep(0.002)
    with torch.cuda.stream(A):
        a0.record()
        d_in.add_(1)
        a1.record()
    torch.cuda.synchronize()
    print(f"   kernel on A during the D2H: {a0.elapsed_time(a1):.3f} ms, finished {b0.elapsed_time(a1):.2f} ms after the D2H started")
# chunked D2H (93 MB x 7) with H2D in between
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(B):
    for k in range(7):
        h_out[k * 93_000_000:(k + 1) * 93_000_000].copy_(d_out[k * 93_000_000:(k + 1) * 93_000_000], non_blocking=True)
with torch.cuda.stream(A):
    for k in range(7):
        d_in.copy_(h_in, non_blocking=True)
torch.cuda.synchronize()
print(f"7 x 93 MB D2H on B + 7 x 20 MB H2D on A concurrently: {(time.perf_counter() - t0) * 1e3:.2f} ms")
