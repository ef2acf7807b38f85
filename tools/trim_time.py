"""Times `rb trim-paf` on the GPU (rb_trim_paf, host buffers in, pinned rows out) beside the literal CPU oracle on the same
input: the bundled fixture and a synthetic pile of long overlapping records.  One JSON line per workload.
    python tools/trim_time.py [out.jsonl]"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gen  # noqa: E402
import orc  # noqa: E402
from rustybam_b200 import capi  # noqa: E402
from rustybam_b200.paf import Paf  # noqa: E402


def measure(ctx, name, paf_text, reps=7):
    recs = Paf.from_text(paf_text).pack()
    t = time.perf_counter()
    want = orc.run_trim_paf(paf_text)
    cpu_s = time.perf_counter() - t
    got = ctx.trim_paf(recs, want=capi.WANT_TEXT, stats=False)["paf_text"]
    assert got == want, name
    ms = []
    for _ in range(reps):
        t = time.perf_counter()
        ctx.trim_paf(recs, want=capi.WANT_TEXT, stats=False, copy=False)
        ms.append((time.perf_counter() - t) * 1e3)
    ctx.set_profiling(True)
    ctx.kernel_times(reset=True)
    ctx.trim_paf(recs, want=capi.WANT_TEXT, stats=False, copy=False)
    kt = ctx.kernel_times(reset=True)
    ctx.set_profiling(False)
    changed = sum(a != b for a, b in zip(sorted(paf_text.splitlines()), sorted(want.splitlines())))
    return dict(workload=name, records=int(recs.c.n_rec), cigar_bytes=int(recs.c.cigar_nbytes), out_bytes=len(want), gpu_e2e_ms=statistics.median(ms),
                gpu_e2e_ms_all=[round(x, 3) for x in ms], oracle_cpu_s=cpu_s, oracle_cores=1, parity="byte-identical",
                kernels={k: dict(launches=v[0], ms=round(v[1], 4)) for k, v in kt.items()}, rows_changed=changed)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    ctx = capi.Context(0)
    lines = [measure(ctx, "bundled .test/asm_small.paf (249 records, 5 query names)", orc.golden_paf())]
    big = gen.random_trim_paf(11, n_names=60, recs_per_name=6, max_ops=30000, span=400000, lead_trail=False)
    lines.append(measure(ctx, "synthetic: 60 query names x <= 6 overlapping records of <= 30 k ops", big))
    ctx.close()
    for ln in lines:
        print(json.dumps(ln))
    if out:
        with open(out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
