#!/usr/bin/env python
"""Per-source-line hot spots of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
   python tools/ncu_lines.py rep.ncu-rep k_serialise [n]"""
import csv, io, subprocess, sys, collections, re
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# layout: blocks per file: "File Name",path / header row / for each source line a row, followed by its SASS rows
agg = collections.defaultdict(lambda: [0, 0, 0])  # (file,line) -> [warp instr, thread instr, stall samples]
src = {}
hdr = None; cur = None; fname = None; first_kernel_done = False
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] in ("Line No", "#"): hdr = r; continue
    if hdr is None: continue
    try:
        ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("Warp Stall Sampling (All Samples)")
    except ValueError:
        continue
    if len(r) != len(hdr): continue
    key = (fname, r[0])
    if r[0].isdigit():
        try:
            agg[key][0] += int(r[ci] or 0); agg[key][1] += int(r[ti] or 0); agg[key][2] += int(r[si] or 0)
        except ValueError:
            pass
        src[key] = r[1]
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
print(f"{kern}: {tot/1e6:.1f}M warp-instr, {tots} stall samples")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    f = key[0].split("/")[-1]
    print(f"{100*v[0]/tot:5.1f}% instr {100*v[2]/tots:5.1f}% stall  act {v[1]/max(v[0],1):4.1f}  {f}:{key[1]:>5}  {src.get(key,'')[:100].strip()}")
