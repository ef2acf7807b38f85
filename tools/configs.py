#!/usr/bin/env python
"""BASELINE.json configs C2-C5 on ONE B200: resident (kernels only, CUDA events on the launch stream, L2 flushed)
and end-to-end (pinned host buffers in, pinned host buffers out) times.  Writes one JSON object per config.
    python tools/configs.py [--haps 16] > profiles/rNN_configs.jsonl"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rustybam_b200 import capi, hostlib

ap = argparse.ArgumentParser()
ap.add_argument("--haps", type=int, default=16, help="haplotypes of the C5-style run (94 = the full config)")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

lib = capi.load()
ctx = capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


PINNED = []


def pin(obj_ptrs):
    for ptr, nbytes in obj_ptrs:
        addr = C.cast(ptr, C.c_void_p).value
        if addr and nbytes:
            rc = lib.rb_host_register(C.c_void_p(addr), nbytes)
            if rc != 0:
                sys.stderr.write(f"rb_host_register({nbytes} bytes) -> {rc}: this buffer stays pageable\n")
            else:
                PINNED.append(addr)


def unpin_all():
    """Page-locked buffers must be unregistered BEFORE their memory is freed (a later allocation that reuses the
    addresses would otherwise collide with the stale registration)."""
    torch.cuda.synchronize()
    while PINNED:
        lib.rb_host_unregister(C.c_void_p(PINNED.pop()))


def timed(f, steps):
    f()
    ms = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = f()
        e1.record(stream)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sum(ms) / len(ms), r


def wall(f, steps):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sum(ts) / len(ts), r


def run(name, paf, width, steps, resident=True):
    out = {"config": name, "records": paf.n_rec, "cigar_bytes": paf.cigar_nbytes}
    if width is None:  # rb stats --paf
        b = ctx.upload(paf)
        out["resident_ms"], s = timed(lambda: ctx.batch_stats(b), steps)
        ctx.batch_free(b)
        out["e2e_ms"], r = wall(lambda: ctx.stats(paf, copy=False), steps)
        out["rows"] = r["n"]
        out["cigar_ops"] = s["n_ops"]
    else:
        wins = paf.tiling_windows(width)
        pin(((wins.c.st, wins.n_win * 8), (wins.c.en, wins.n_win * 8), (wins.c.bed_row, wins.n_win * 4), (wins.c.t_id, wins.n_win * 4)))
        out["bed_rows"] = wins.n_win
        if resident:
            b = ctx.upload(paf, wins)
            out["resident_ms"], s = timed(lambda: ctx.batch_liftover(b, with_stats=True, want=capi.WANT_TEXT), steps)
            ctx.batch_free(b)
            out.update(cigar_ops=s["n_ops"], pairs=s["n_pairs"], rows=s["n_out"], out_bytes=s["out_bytes"])
        out["e2e_ms"], r = wall(lambda: ctx.liftover(paf, wins, want=capi.WANT_TEXT, stats=True, copy=False), steps)
        out.update(rows=r["n_out"], out_bytes=r["paf_nbytes"], pairs=r["n_pairs"])
    if width is not None:
        unpin_win = PINNED[-4:]  # the four window columns registered above die with `wins`
        torch.cuda.synchronize()
        for a in unpin_win:
            lib.rb_host_unregister(C.c_void_p(a))
            PINNED.remove(a)
        wins.close()
    out["rows_per_s_e2e"] = out["rows"] / (out["e2e_ms"] * 1e-3)
    out["cigar_gb_per_s_e2e"] = out["cigar_bytes"] / (out["e2e_ms"] * 1e-3) / 1e9
    if "resident_ms" in out:
        out["rows_per_s_resident"] = out["rows"] / (out["resident_ms"] * 1e-3)
        out["cigar_gb_per_s_resident"] = out["cigar_bytes"] / (out["resident_ms"] * 1e-3) / 1e9
    print(json.dumps(out), flush=True)


one = hostlib.HostPaf.synth(scale=1.0)
pin(((one.c.cigar, one.cigar_nbytes),))
run("C2: rb stats --paf, 1 haplotype (~50 M ops)", one, None, args.steps)
run("C3: liftover 100 kb windows + stats, 1 haplotype", one, 100_000, args.steps)
run("C4: liftover 1 kb windows + stats, 1 haplotype", one, 1000, args.steps)
unpin_all()
one.close()
if args.haps > 1:
    t0 = time.time()
    many = hostlib.HostPaf.synth(scale=1.0, n_hap=args.haps, threads=os.cpu_count() or 8)
    pin(((many.c.cigar, many.cigar_nbytes),))
    sys.stderr.write(f"generated {args.haps} haplotypes in {time.time() - t0:.1f} s ({many.cigar_nbytes / 1e9:.2f} GB of CIGAR text)\n")
    run(f"C5-style: {args.haps} haplotypes vs CHM13-like, 10 kb windows + stats, ONE GPU (sliced, emission-order gather)", many, 10_000,
        max(2, args.steps // 2), resident=(args.haps <= 24))
    unpin_all()
ctx.close()
