#!/usr/bin/env python
"""Time-bounded random parity campaign on the GPU box: random PAFs / BEDs / policies / pipeline modes against the
oracle, beyond the fixed seeds of tests/.  Stops at the first mismatch and leaves the inputs in gpurun_out/.
    python tools/fuzz_gpu.py [seconds] [first_seed]"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import orc  # noqa: E402
from rustybam_b200 import bamstats, capi, hostlib, liftover  # noqa: E402
from rustybam_b200.paf import ReferencePanic  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
os.environ.setdefault("RB_MULTI_MIN_BYTES", "0")
ctx = capi.Context(0)
multi = capi.Context(devices=[0, 0])


def both(label, ref_fn, gpu_fn, dump):
    """Same bytes, or a panic on both sides."""
    try:
        want = ref_fn()
    except orc.ReferencePanic:
        want = "PANIC"
    try:
        got = gpu_fn()
    except ReferencePanic:
        got = "PANIC"
    if got != want:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        for name, data in dump.items():
            open(os.path.join(ROOT, "gpurun_out", f"fuzz_fail_{name}"), "wb").write(data)
        print("MISMATCH", label, "want", (want if want == "PANIC" else len(want)), "got", (got if got == "PANIC" else len(got)))
        sys.exit(1)
    return want


t_end = time.time() + budget
n = counts = 0
panics = 0
n_trim_panics = 0
seed = seed0
while time.time() < t_end:
    rng = random.Random(seed)
    style = rng.choice(["eqx", "eqx", "all"])
    canonical = rng.random() < 0.6
    paf_text, contigs = gen.random_paf(seed, n_contigs=rng.randint(1, 5), recs_per_contig=rng.randint(1, 12), style=style,
                                       canonical=canonical, allow_zero=(not canonical and rng.random() < 0.5),
                                       lead_trail=rng.random() < 0.7, max_ops=rng.choice([5, 40, 300, 1500]),
                                       clips=(style == "all" and rng.random() < 0.3))
    if rng.random() < 0.15:  # break one record the way real files are broken: the reference panics (or skips the line), so must we
        lines = paf_text.splitlines()
        k = rng.randrange(len(lines))
        f = lines[k].split(b"\t")
        cg = [i for i, x in enumerate(f) if x.startswith(b"cg:Z:")][0]
        how = rng.choice(["span_t", "span_q", "lead_d", "bad_op", "no_len", "empty", "bad_col", "all_indel", "huge"])
        if how == "span_t": f[8] = str(int(f[8]) + rng.choice([1, 5])).encode()
        elif how == "span_q": f[3] = str(int(f[3]) + 1).encode()
        elif how == "lead_d": f[cg] = b"cg:Z:3D" + f[cg][5:]; f[8] = str(int(f[8]) + 3).encode()
        elif how == "bad_op": f[cg] = f[cg][:-1] + rng.choice([b"Z", b"m", b"*", b" "])
        elif how == "no_len": f[cg] = b"cg:Z:=" + f[cg][5:]
        elif how == "empty": f[cg] = b"cg:Z:"; f[8] = f[7]; f[3] = f[2]
        elif how == "bad_col": f[rng.choice([1, 2, 3, 6, 7, 8, 9, 10, 11])] = rng.choice([b"x", b"-1", b"", b"1.5"])
        elif how == "all_indel": f[cg] = b"cg:Z:2I3D1I"; f[8] = str(int(f[7]) + 3).encode(); f[3] = str(int(f[2]) + 3).encode()
        elif how == "huge": f[cg] = b"cg:Z:99999999999=" + f[cg][5:]
        lines[k] = b"\t".join(f)
        paf_text = b"\n".join(lines) + b"\n"
    kind = rng.choice(["tile", "tile", "random", "random_ids", "sorted"])
    if kind == "tile":
        bed_text = gen.tiling_bed(contigs, rng.choice([1, 2, 3, 7, 25, 100, 1000]), with_ids=rng.random() < 0.3)
    else:
        bed_text = gen.random_bed(seed, contigs, rng.randint(1, 120), max_w=rng.choice([3, 50, 400]), with_ids=(kind == "random_ids"),
                                  sort=(kind == "sorted"))
    policy = rng.randint(0, 1)
    dump = {"in.paf": paf_text, "in.bed": bed_text, "info.txt": f"seed {seed} policy {policy} style {style} kind {kind}\n".encode()}
    mode = rng.choice(["plain", "sliced", "stream"])
    if mode == "sliced":
        ctx.set_slicing(rng.choice([16, 64, 512]))
    if mode == "stream":
        ctx.set_lift_mode(capi.LIFT_STREAM)
    try:
        lifted = both(f"liftover[{mode}] seed {seed}", lambda: orc.run_liftover(paf_text, bed_text, policy=policy, threads=2),
                      lambda: liftover.run_liftover(ctx, paf_text, bed_text, policy=policy), dump)
    finally:
        ctx.set_slicing()
        ctx.set_lift_mode(capi.LIFT_SEARCH)
    if lifted != "PANIC" and lifted:
        both(f"stats of lifted seed {seed}", lambda: orc.run_stats(lifted), lambda: bamstats.run_stats(ctx, lifted), dump)
        # RB_WANT_STATS_TEXT: the device prints those stats rows itself; and the same call on a multi-device context (two
        # contexts on this GPU), sometimes in forced slices
        hp = hostlib.HostPaf.from_text(paf_text)
        hw = hp.windows_from_bed_text(bed_text)
        hdr = bamstats.print_cigar_stats_header().encode()
        both(f"stats-text seed {seed}", lambda: orc.run_stats(lifted),
             lambda: hdr + ctx.liftover(hp, hw, policy=policy, want=capi.WANT_STATS_TEXT, stats=False)["paf_text"], dump)
        if rng.random() < 0.5:
            multi.set_slicing(rng.choice([0, 16, 64, 512]))
            try:
                both(f"multi-device seed {seed}", lambda: lifted, lambda: multi.liftover(hp, hw, policy=policy, want=capi.WANT_TEXT, stats=True)["paf_text"], dump)
            finally:
                multi.set_slicing()
    else:
        panics += 1
    both(f"stats seed {seed}", lambda: orc.run_stats(paf_text), lambda: bamstats.run_stats(ctx, paf_text), dump)
    both(f"invert seed {seed}", lambda: orc.run_invert(paf_text), lambda: liftover.run_invert(ctx, paf_text), dump)
    ms = rng.choice([0, 1, 3, 10, 100])
    both(f"break-paf {ms} seed {seed}", lambda: orc.run_break_paf(paf_text, ms, policy), lambda: liftover.run_break_paf(ctx, paf_text, ms, policy), dump)
    if style == "eqx" and rng.random() < 0.5:
        qlens = {}
        for ln in paf_text.splitlines():
            f = ln.split(b"\t")
            if f[1].isdigit():  # (a record broken in column 2 above has no usable query length: its line is skipped anyway)
                qlens[f[0].decode()] = int(f[1])
        qbed = gen.tiling_bed(qlens, rng.choice([3, 11, 50]))
        dump["in.qbed"] = qbed
        both(f"qbed seed {seed}", lambda: orc.run_liftover(paf_text, qbed, qbed=True, policy=policy, threads=2),
             lambda: liftover.run_liftover(ctx, paf_text, qbed, policy=policy, qbed=True), dump)
    # rb trim-paf: piles of query-overlapping records, random scores, sometimes one broken record
    tp = gen.random_trim_paf(seed, n_names=rng.randint(1, 6), recs_per_name=rng.randint(1, 8), max_ops=rng.choice([3, 20, 120, 1500]),
                             style=style, canonical=canonical, allow_zero=(not canonical and rng.random() < 0.5),
                             lead_trail=rng.random() < 0.7, span=rng.choice([5, 80, 2000]), big=rng.choice([0, 0, 1]))
    if rng.random() < 0.1:
        lines = tp.splitlines()
        k = rng.randrange(len(lines))
        f = lines[k].split(b"\t")
        cg = [i for i, x in enumerate(f) if x.startswith(b"cg:Z:")][0]
        how = rng.choice(["span_t", "span_q", "lead_d", "bad_op", "no_len", "all_indel"])
        if how == "span_t": f[8] = str(int(f[8]) + rng.choice([1, 5])).encode()
        elif how == "span_q": f[3] = str(int(f[3]) + 1).encode()
        elif how == "lead_d": f[cg] = b"cg:Z:3D" + f[cg][5:]; f[8] = str(int(f[8]) + 3).encode()
        elif how == "bad_op": f[cg] = f[cg][:-1] + rng.choice([b"Z", b"m", b"*"])
        elif how == "no_len": f[cg] = b"cg:Z:=" + f[cg][5:]
        elif how == "all_indel": f[cg] = b"cg:Z:2I3D1I"; f[8] = str(int(f[7]) + 3).encode(); f[3] = str(int(f[2]) + 3).encode()
        lines[k] = b"\t".join(f)
        tp = b"\n".join(lines) + b"\n"
    scores = rng.choice([(1, 1, 1), (1, 1, 1), (rng.randint(1, 4), rng.randint(0, 4), rng.randint(0, 4))])
    rc_flag = rng.random() < 0.5
    dump["trim.paf"] = tp
    dump["info.txt"] += f"trim scores {scores} remove_contained {rc_flag}\n".encode()
    tpol = rng.randint(0, 1)
    dump["info.txt"] += f"trim policy {tpol}\n".encode()
    trimmed = both(f"trim-paf seed {seed}", lambda: orc.run_trim_paf(tp, *scores, rc_flag, policy=tpol),
                   lambda: liftover.run_trim_paf(ctx, tp, *scores, rc_flag, policy=tpol), dump)
    n_trim_panics += trimmed == "PANIC"
    n += 1
    seed += 1
print(f"trim-paf: {n} piles, {n_trim_panics} reference panics reproduced")
print(f"fuzz ok: {n} cases (seeds {seed0}..{seed - 1}), {panics} reference panics reproduced, {budget:.0f} s")
