#!/usr/bin/env python
"""Top SASS instructions of one kernel of an .ncu-rep by executed count / stall samples.
   python tools/ncu_hot.py rep.ncu-rep k_tokenise [n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first capture of that kernel only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
h = rows[start]; body = [r for r in rows[start + 1:end] if len(r) == len(h)]
ci, si, ii = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
tot = sum(int(r[ci]) for r in body); tots = sum(int(r[si]) for r in body)
print(f"{kern}: {len(body)} SASS instrs, {tot} warp-instr executed, {tots} stall samples")
# by opcode class
from collections import Counter
ops = Counter(); stall = Counter()
for r in body:
    op = r[ii].split()[0] if not r[ii].strip().startswith("@") else r[ii].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r[ci]); stall[op] += int(r[si])
print("by opcode (exec%, stall%):", [(k, round(100 * v / tot, 1), round(100 * stall[k] / max(tots, 1), 1)) for k, v in ops.most_common(14)])
print("top by stall samples:")
for r in sorted(body, key=lambda r: -int(r[si]))[:n]:
    print(f"  {int(r[si]):6d} {int(r[ci]):9d}  {r[ii].strip()[:90]}")
