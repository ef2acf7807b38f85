#!/usr/bin/env python
"""bench.py — PAF liftover + stats hot path on B200 (the metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic input: `rb liftover --bed
<1 kb tiling windows>` over a synthetic whole-genome-scale eqx PAF followed by the per-row
`rb stats --paf` counters (config C4 of BASELINE.json at N = 1).  For N > 1 the job is N
haplotypes vs the same reference, records partitioned across the GPUs by target contig (LPT on
CIGAR bytes), no collective on the data path: per-GPU work stays ~constant ("weak").

One JSON line on stdout (rank 0).  `value` = lifted window-records/s with inputs resident in HBM
(kernel sequence only, CUDA events); `e2e` = the same through the C-ABI call rb_liftover() with
pinned HOST buffers in and pinned host buffers out (H2D + kernels + D2H timed).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "lifted window-records/s"
UNIT = "records/s"
CONTIG_LEN = [248387328, 242696752, 201105948, 193574945, 182045439, 172126628, 160567428, 146259331, 150617247, 134758134,
              135127769, 133324548, 113566686, 101161492, 99753195, 96330374, 84276897, 80542538, 61707364, 66210255, 45090682,
              51324926, 154259566, 62460029, 16569]
CONTIG_NAME = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]
WINDOW = 1000
CPU_SAMPLE_CONTIGS = ["chr21", "chr22", "chrM"]  # --impl reference, per step: ~96 Mbp of the ~3.1 Gbp workload
CPU_BASELINE_CONTIGS = ["chr16", "chr17", "chr18", "chr19", "chr20", "chr21", "chr22", "chrM"]  # cpu_baseline: ~0.5 Gbp, one pass


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def cpu_reference_sample(threads, contigs=None):
    """The reference's CPU path (restated: oracle/, literal per-base algorithm) on a bounded sample of
    the same workload: all records of `contigs` + their 1 kb windows -> liftover -> stats."""
    contigs = contigs or CPU_SAMPLE_CONTIGS
    import orc
    from rustybam_b200 import hostlib
    mask_scale = 1.0
    paf = hostlib.HostPaf.synth(scale=mask_scale, n_hap=1, threads=min(8, os.cpu_count() or 1))
    texts, beds, nrec = [], [], 0
    for nm in contigs:
        tid = paf.find_name(nm)
        t, n = paf.text_of_contig(tid)
        texts.append(t)
        nrec += n
        beds.append(paf.tiling_bed_text(WINDOW, tid))
    paf_text, bed_text = b"".join(texts), b"".join(beds)
    paf.close()
    return paf_text, bed_text, nrec


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  The Rust
    binary cannot be built in this image (no cargo/rustc, no network), so this is the oracle port."""
    if rank != 0:
        return
    import orc
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    threads = os.cpu_count() or 8  # `rb -t N`: every host core (the reference's default is -t 8, cli.rs:17-19)
    paf_text, bed_text, nrec = cpu_reference_sample(threads)
    times, rows = [], 0
    for i in range(args.warmup + args.steps):
        r = orc.bench_pipeline(paf_text, bed_text, threads=threads)
        if i >= args.warmup:
            times.append(r["secs_liftover"] + r["secs_stats"])
        rows = r["rows"]
    sec = sum(times) / len(times)
    value = rows / sec
    sample = (f"records of {'+'.join(CPU_SAMPLE_CONTIGS)} ({nrec} records, {len(paf_text) / 1e6:.1f} MB PAF text) x their 1 kb windows "
              f"-> {rows} rows; liftover+stats incl. PAF parse and printing; restated reference CPU path, not the rb binary")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32/u64 (+f32 identities)", "data": "synthetic",
            "config": {"workload": "C4: rb liftover --bed <1 kb tiling windows> over synthetic HG002-vs-CHM13-scale eqx PAF + per-row rb stats --paf",
                       "window_bp": WINDOW, "sample_per_step": "+".join(CPU_SAMPLE_CONTIGS), "host_cores": os.cpu_count()},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Host side of the e2e path: run this rank's threads (and so first-touch its page-locked buffers) on the NUMA
    node the GPU's PCIe root hangs off, so that the 0.8 GB of DMA per step does not cross the socket interconnect.
    Returns (node, n_cpus, previous affinity) or None when the topology cannot be read (single-node hosts: no-op)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)  # CUDA's numbering (honours CUDA_VISIBLE_DEVICES), unlike NVML's
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        before = os.sched_getaffinity(0)
        cpus &= before
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus), before
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale (1.0 = ~3.1 Gbp, the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from rustybam_b200 import build, capi, hostlib

    build.build_all()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if not os.environ.get("RB_BENCH_NO_NUMA") else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- this rank's shard: `world` haplotypes, contigs of LPT bin `rank` (rustybam_b200/shard.py) ----
    from rustybam_b200 import shard as sharding
    n_hap = world
    t0 = time.time()
    shard, shard_info = sharding.make_shard(rank, world, scale=args.scale, threads=max(1, (os.cpu_count() or 8) // world))
    wins = shard.tiling_windows(WINDOW)
    gen_s = time.time() - t0

    lib = capi.load()
    # page-lock the big input buffers so that the e2e H2D copies are DMA from pinned memory
    pinned = []
    ids_bytes = int(wins.c.ids_off[wins.n_win]) if wins.c.ids_off else 0  # 3-column BED: ids are formatted on the GPU
    for ptr, nbytes in ((shard.c.cigar, shard.cigar_nbytes), (wins.c.st, wins.n_win * 8), (wins.c.en, wins.n_win * 8),
                        (wins.c.ids, ids_bytes), (wins.c.ids_off, (wins.n_win + 1) * 8 if ids_bytes else 0),
                        (wins.c.bed_row, wins.n_win * 4), (wins.c.t_id, wins.n_win * 4)):
        addr = C.cast(ptr, C.c_void_p).value
        if addr and nbytes and lib.rb_host_register(C.c_void_p(addr), nbytes) == 0:
            pinned.append(addr)

    ctx = capi.Context(local_rank)
    # a real (non-default) stream: the library launches on it and the CUDA events below are recorded on it.  (Handing
    # the library torch's default stream, handle 0, means "use your own stream" in the C ABI: events recorded on the
    # default stream would then not bracket the kernels.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB): flushed between timed steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident (kernel-only) measurement ----
    b = ctx.upload(shard, wins)
    summ = None
    for _ in range(args.warmup):
        summ = ctx.batch_liftover(b, with_stats=True, want=capi.WANT_TEXT)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        summ = ctx.batch_liftover(b, with_stats=True, want=capi.WANT_TEXT)
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_resident = time.perf_counter() - wall0
    ms_resident = sum(step_ms) / len(step_ms)

    # ---- per-kernel times for the roofline of the dominant kernel (separate pass, events around each launch) ----
    ctx.set_profiling(True)
    ctx.kernel_times(reset=True)
    n_prof = max(3, args.steps // 4)
    prof_ms = []
    for _ in range(n_prof):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.batch_liftover(b, with_stats=True, want=capi.WANT_TEXT)
        e1.record(stream)
        e1.synchronize()
        prof_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    ktimes = ctx.kernel_times(reset=True)
    ctx.set_profiling(False)
    ctx.batch_free(b)

    # ---- end to end through the C ABI: pinned host buffers in, pinned host buffers out ----
    for _ in range(args.warmup):
        ctx.liftover(shard, wins, want=capi.WANT_TEXT, stats=True, copy=False)
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.liftover(shard, wins, want=capi.WANT_TEXT, stats=True, copy=False)
        e1.record(stream)
        e1.synchronize()
        e2e_ms.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop()  # sampled across the resident, per-kernel and end-to-end timed loops
    ms_e2e = sum(e2e_ms) / len(e2e_ms)

    n_out, n_pairs, n_ops = summ["n_out"], summ["n_pairs"], summ["n_ops"]
    out_bytes, cigar_bytes = summ["out_bytes"], summ["cigar_bytes"]
    h2d = (cigar_bytes + shard.n_rec * (8 * 8 + 1 + 8) + wins.n_win * (8 + 8 + 4) + ((wins.n_win + 1) * 8 + ids_bytes if ids_bytes else 0))
    d2h = out_bytes + (n_out + 1) * 8 + n_out * 40

    # ---- reduce over ranks: time = max, units = sum ----
    tot = torch.tensor([float(n_out), float(cigar_bytes), float(out_bytes), float(n_pairs), float(n_ops), float(h2d), float(d2h)],
                       dtype=torch.float64, device="cuda")
    tmax = torch.tensor([ms_resident, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot, tmax = tot.tolist(), tmax.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic (compulsory) bytes per launch of each kernel, rank 0's shard — DESIGN.md §5
        smp_bytes = (n_ops // 32) * 4 * 48  # 4 counters blocks (1 sample + 3 sub-samples) of 48 B per 32-op chunk
        alg = {
            "k_tokenise": cigar_bytes + 4 * n_ops,                       # text in, one 4-byte op word out
            "k_samples": 4 * n_ops + smp_bytes,                          # op words in, samples out
            "k_scan_lift": 4 * n_ops + smp_bytes + 16 * wins.n_win + 128 * n_pairs,
            "k_combine": n_pairs * (128 + 16 + 112 + 4),
            # windows (16 B) in, PairRes (112 B) + line size (4 B) out per pair; every op word and sample block read once
            "k_lift": n_pairs * (16 + 112 + 4) + 4 * n_ops + smp_bytes,
            # PairRes + line offset in, every output byte + the stats row + the line offset out, copied CIGAR text in
            "k_serialise": n_pairs * (112 + 8) + out_bytes + n_out * (40 + 8) + cigar_bytes,
            "k_scan_lines": n_pairs * (4 + 16),
        }
        total_k = sum(ms for _, ms in ktimes.values()) or 1.0
        dom = max((k for k in ktimes if k in alg), key=lambda k: ktimes[k][1])
        launches, ms_sum = ktimes[dom]
        dom_ms = ms_sum / max(launches, 1)
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": tot[0] / (tmax[0] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tmax[0], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32/u64 (+f32 identities)", "data": "synthetic",
            "config": {"workload": "C4: rb liftover --bed <1 kb tiling windows> over synthetic HG002-vs-CHM13-scale eqx PAF + per-row rb stats --paf",
                       "haplotypes": n_hap, "window_bp": WINDOW, "scale": args.scale, "records": int(shard.n_rec) if world == 1 else None,
                       "bed_rows_rank0": wins.n_win, "pairs": int(tot[3]), "cigar_ops": int(tot[4]), "cigar_bytes": int(tot[1]),
                       "out_bytes": int(tot[2]), "partition": "target contig, LPT on CIGAR bytes, no collective", "lpt_balance": shard_info.get("balance"),
                       "l2": "256 MiB buffer written between timed steps (L2 flush); inputs+outputs > L2"},
            "cigar_gb_per_s": tot[1] / (tmax[0] * 1e-3) / 1e9,
            "e2e": {"value": tot[0] / (tmax[1] * 1e-3), "unit": UNIT, "ms_per_step": tmax[1], "h2d_bytes_per_step": int(tot[5]),
                    "d2h_bytes_per_step": int(tot[6]), "cigar_gb_per_s": tot[1] / (tmax[1] * 1e-3) / 1e9},
            "gpu_launches": int(sum(v[0] for v in ktimes.values()) // n_prof) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg[dom]),
                         "kernel_ms": dom_ms, "kernel_share_of_step": ms_sum / total_k,
                         "step_ms_with_per_kernel_events": sum(prof_ms) / len(prof_ms),
                         "all_kernels_ms": {k: v[1] / max(v[0], 1) for k, v in ktimes.items()},
                         "all_kernels_frac": {k: (alg[k] / (v[1] / max(v[0], 1) * 1e-3) / 1e9) / peak for k, v in ktimes.items() if k in alg}},
            "wall_s_resident_loop": wall_resident, "gen_s": gen_s,
        }
        line["numa"] = {"node": numa[0], "cpus": numa[1]} if numa else None
        if numa:
            os.sched_setaffinity(0, numa[2])  # the CPU baseline gets every host core back
        if not args.no_cpu_baseline:
            import orc
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
            threads = os.cpu_count() or 8
            paf_text, bed_text, nrec = cpu_reference_sample(threads, CPU_BASELINE_CONTIGS)
            r = orc.bench_pipeline(paf_text, bed_text, threads=threads)
            sec = r["secs_liftover"] + r["secs_stats"]
            line["cpu_baseline"] = {
                "value": r["rows"] / sec, "unit": UNIT, "cores": threads, "kind": "port", "host_cores": os.cpu_count(),
                "seconds": sec, "sample": f"records of {'+'.join(CPU_BASELINE_CONTIGS)} ({nrec} records) x their 1 kb windows -> {r['rows']} rows; "
                                          "restated reference CPU path (oracle/), not the rb binary; PAF parse + liftover + print + stats"}
        print(json.dumps(line))
    for addr in pinned:
        lib.rb_host_unregister(C.c_void_p(addr))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
