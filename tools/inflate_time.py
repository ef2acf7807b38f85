#!/usr/bin/env python
"""BGZF inflate: the device (rb_inflate_bgzf, compressed bytes in host memory -> inflated text in pinned host memory, copies
included; kernel time separately) beside the host's block-parallel zlib reader (rbhost read_all on every host core), on the
synthetic C4 PAF (one haplotype) written as BGZF.  Outputs compared.
    python tools/inflate_time.py [n_hap] > profiles/rNN_inflate.json"""
import json
import os
import struct
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustybam_b200 import capi, hostlib

n_hap = int(sys.argv[1]) if len(sys.argv) > 1 else 1
paf = hostlib.HostPaf.synth(scale=1.0, n_hap=n_hap, threads=os.cpu_count() or 8)
text = paf.text()
BLK = 65280


def one(off):
    chunk = text[off:off + BLK]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = co.compress(chunk) + co.flush()
    return (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(cdata) + 25) + cdata +
            struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
    z = b"".join(ex.map(one, range(0, len(text), BLK)))
z += one(len(text))  # empty EOF block
path = "/dev/shm/rb_inflate_time.paf.bgz"
open(path, "wb").write(z)
res = {"text_bytes": len(text), "bgzf_bytes": len(z), "blocks": (len(text) + BLK - 1) // BLK, "host_cores": os.cpu_count()}

ctx = capi.Context(0)
ctx.set_profiling(True)
got = ctx.inflate_bgzf(z)
assert got == text, "device inflate differs from the text"
ctx.kernel_times(reset=True)
t = []
for _ in range(5):
    t0 = time.perf_counter()
    p, n = capi.C.c_void_p(), capi.C.c_uint64()
    ctx._check(ctx.lib.rb_inflate_bgzf(ctx.h, z, len(z), capi.C.byref(p), capi.C.byref(n)))
    t.append(time.perf_counter() - t0)
    ctx.lib.rb_free_text(ctx.h, p)
kt = ctx.kernel_times(reset=True)
k_ms = kt["k_inflate_bgzf"][1] / max(kt["k_inflate_bgzf"][0], 1)
res["gpu_call_s"] = min(t)
res["gpu_call_text_gb_per_s"] = len(text) / min(t) / 1e9
res["gpu_kernel_ms"] = k_ms
res["gpu_kernel_text_gb_per_s"] = len(text) / (k_ms * 1e-3) / 1e9
ctx.close()

t = []
for _ in range(3):
    t0 = time.perf_counter()
    h = hostlib.read_all(path)
    t.append(time.perf_counter() - t0)
assert h == text
res["host_read_all_s"] = min(t)
res["host_text_gb_per_s"] = len(text) / min(t) / 1e9
res["note"] = ("gpu_call = rb_inflate_bgzf from pageable host bytes to pinned host text (header hop, H2D of the compressed bytes, kernel, D2H of the "
               "text); host = rbhost read_all of the same file from /dev/shm (file read + zlib on every core)")
print(json.dumps(res))
