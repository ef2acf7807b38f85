"""Full-size checks of the CUDA path (BASELINE.json configs C2-C4 shapes): exact parity against the
oracle where the oracle finishes in seconds (2 % genome scale), and size-independent properties at
the full ~3.1 Gbp / ~50 M-op scale (independent numpy tokeniser, conservation of matched bases over
tiling windows, idempotence of lifting the lifted rows, offsets/sortedness)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import orc
from rustybam_b200 import bamstats, capi, hostlib
from rustybam_b200.paf import Paf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def full():
    return hostlib.HostPaf.synth(scale=1.0)


@pytest.mark.parametrize("width", [1000, 100_000, 2_000_000])  # 2 Mb windows: ~80 KB rows, their verbatim runs leave through k_copy_mid
def test_synth_2pct_exact_parity(ctx, width):
    paf = hostlib.HostPaf.synth(scale=0.02)
    wins = paf.tiling_windows(width)
    res = ctx.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    paf_text, bed_text = paf.text(), paf.tiling_bed_text(width)
    want = orc.run_liftover(paf_text, bed_text, threads=8)
    assert res["paf_text"] == want
    out = Paf.from_text(res["paf_text"])
    st = (bamstats.print_cigar_stats_header() + bamstats.stats_rows(out, res["stats"])).encode()
    assert st == orc.run_stats(want)
    assert bamstats.run_stats(ctx, paf_text) == orc.run_stats(paf_text)


def numpy_record_stats(paf):
    """Independent tokeniser: per-record (=, X, I, D bases; I, D events) straight from the CIGAR bytes."""
    n = paf.n_rec
    off = np.ctypeslib.as_array(paf.c.cigar_off, shape=(n + 1,)).astype(np.int64)
    cig = np.ctypeslib.as_array(paf.c.cigar, shape=(paf.cigar_nbytes,))
    out = np.zeros((n, 6), dtype=np.int64)
    r0 = 0
    while r0 < n:  # slices of ~32 MB of text, aligned to record boundaries
        r1 = r0 + 1
        while r1 < n and off[r1 + 1] - off[r0] < (32 << 20):
            r1 += 1
        b = cig[off[r0]:off[r1]]
        is_op = (b < 48) | (b > 57)
        idx = np.nonzero(is_op)[0]
        before = np.cumsum(is_op) - is_op          # ops strictly before each byte
        nxt = idx[np.minimum(before, len(idx) - 1)]  # position of the op character that ends this byte's number
        expo = (nxt - np.arange(len(b)) - 1).astype(np.int64)
        digit = np.where(is_op, 0, b - 48).astype(np.float64)
        lens = np.bincount(before[~is_op], weights=(digit * np.power(10.0, np.maximum(expo, 0)))[~is_op], minlength=len(idx))
        lens = np.rint(lens).astype(np.int64)
        codes = b[idx]
        rec_of_op = np.searchsorted(off[r0:r1 + 1] - off[r0], idx, side="right") - 1
        for col, ch in enumerate(b"=XID"):
            out[r0:r1, col] = np.bincount(rec_of_op[codes == ch], weights=lens[codes == ch], minlength=r1 - r0)
        out[r0:r1, 4] = np.bincount(rec_of_op[codes == ord("I")], minlength=r1 - r0)
        out[r0:r1, 5] = np.bincount(rec_of_op[codes == ord("D")], minlength=r1 - r0)
        r0 = r1
    return out


def test_full_scale_stats_vs_numpy_tokeniser(ctx, full):
    st = ctx.stats(full)
    ref = numpy_record_stats(full)
    assert st["n"] == full.n_rec
    assert (st["equal"].astype(np.int64) == ref[:, 0]).all()
    assert (st["diff"].astype(np.int64) == ref[:, 1]).all()
    assert (st["ins"].astype(np.int64) == ref[:, 2]).all()
    assert (st["del"].astype(np.int64) == ref[:, 3]).all()
    assert (st["ins_events"].astype(np.int64) == ref[:, 4]).all()
    assert (st["del_events"].astype(np.int64) == ref[:, 5]).all()
    # bamstats.rs:138-142 recomputed in numpy f32: 0 ULP
    eq, tot = st["equal"].astype(np.float32), (st["equal"] + st["diff"] + st["del"] + st["ins"]).astype(np.float32)
    assert (st["id_by_all"] == (np.float32(100.0) * eq) / tot).all()


def test_full_scale_liftover_properties(ctx, full):
    wins = full.tiling_windows(1000)
    res = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    n_out, text, off = res["n_out"], res["paf_text"], res["line_off"]
    assert n_out > 3_000_000 and res["n_pairs"] >= n_out
    assert off[0] == 0 and off[-1] == len(text) and (np.diff(off.astype(np.int64)) > 0).all()
    ends = np.frombuffer(text, dtype=np.uint8)[off[1:].astype(np.int64) - 1]
    assert (ends == 10).all()
    # conservation: tiling windows partition every record's match columns; the only bases that may go missing
    # are the ones the reference itself loses (Q2: a window that starts on the last base before an insertion
    # slides past the insertion, and the previous window ended one base earlier) -> at most one per row
    rec_stats = ctx.stats(full)
    for key in ("equal", "diff"):
        lost = int(rec_stats[key].astype(np.int64).sum()) - int(res["stats"][key].astype(np.int64).sum())
        assert 0 <= lost <= n_out // 100, (key, lost)
    # rows are in emission order: record-major (file order inside a contig), window-minor
    assert (np.diff(res["rec_idx"].astype(np.int64)) >= 0).all()
    # every row lies inside its window and the numeric mirror matches the text (sample of rows)
    rng = np.random.default_rng(1)
    for i in rng.integers(0, n_out, 2000):
        f = text[int(off[i]):int(off[i + 1]) - 1].split(b"\t")
        assert (int(f[2]), int(f[3]), int(f[7]), int(f[8]), int(f[9]), int(f[10])) == (
            int(res["q_st"][i]), int(res["q_en"][i]), int(res["t_st"][i]), int(res["t_en"][i]), int(res["nmatch"][i]), int(res["aln_len"][i]))
        w = int(res["win_idx"][i])
        assert f[12] == b"id:Z:" + f[5] + b":" + str(int(f[7]) // 1000 * 1000 + 1).encode() + b"-" + f[12].split(b"-")[-1]
    # idempotence: lifting the lifted rows of one contig over the same windows returns the same rows; a row that
    # now sits strictly inside its window takes the reference's early return (Q3) and prints its own, empty id
    lo, hi = int(off[0]), int(off[200_000])
    sub = hostlib.HostPaf.from_text(text[lo:hi])
    again = ctx.liftover(sub, sub.windows_from_bed_text(full.tiling_bed_text(1000)), want=capi.WANT_TEXT, stats=False)
    first, second = text[lo:hi].split(b"\n"), again["paf_text"].split(b"\n")
    assert len(first) == len(second)
    n_early = 0
    for x, y in zip(first, second):
        if x != y:
            fx, fy = x.split(b"\t"), y.split(b"\t")
            assert fy[12] == b"id:Z:" and fx[:12] == fy[:12] and fx[13:] == fy[13:]
            assert int(fx[7]) % 1000 != 0 and int(fx[8]) % 1000 != 0  # strictly inside its window
            n_early += 1
    assert n_early < len(first) // 50


# ---------------------------------------------------------------- byte parity at the BASELINE sizes
# The literal oracle needs 24 B per alignment column, so it cannot run the whole ~3.1 Gbp input in seconds — but rows are
# emitted contig by contig, record-major (liftover.rs:151-164), so the rows of one contig in the FULL-SIZE call are exactly
# what `rb liftover` prints for that contig's records and windows alone: the GPU rows of a few contigs of the full-scale
# call are compared byte for byte with the oracle run on those contigs (paf.rs:379-430 parse -> liftover.rs:107-167 ->
# paf.rs:923-943 print -> bamstats.rs:91-154,225-270).
PARITY_CONTIGS = ["chr20", "chr21", "chr22", "chrM"]


def _rows_of_records(res, rec_lo, rec_hi):
    """Row range of records [rec_lo, rec_hi) (consecutive in file order, one target): rows are record-major."""
    idx = res["rec_idx"].astype(np.int64)
    rows = np.flatnonzero((idx >= rec_lo) & (idx < rec_hi))
    if len(rows) == 0:
        return 0, 0
    assert rows[-1] - rows[0] + 1 == len(rows)  # one contiguous run of rows
    return int(rows[0]), int(rows[-1]) + 1


def _record_runs(paf, tid):
    """Runs of consecutive records whose target is `tid` (one run per haplotype in a haplotype-major PAF)."""
    t = np.ctypeslib.as_array(paf.c.t_id, shape=(paf.n_rec,))
    r = np.flatnonzero(t == tid)
    cuts = np.flatnonzero(np.diff(r) != 1) + 1
    return [(int(x[0]), int(x[-1]) + 1) for x in np.split(r, cuts)]


@pytest.mark.parametrize("width", [1000, 100_000])  # C4 and C3 at their stated size
def test_full_scale_liftover_bytes_equal_oracle_on_contigs(ctx, full, width):
    wins = full.tiling_windows(width)
    res = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    off, checked = res["line_off"].astype(np.int64), 0
    for nm in PARITY_CONTIGS:
        tid = full.find_name(nm)
        (lo, hi), = _record_runs(full, tid)
        want = orc.bench_pipeline_keep(full.text(lo, hi), full.tiling_bed_text(width, tid), threads=os.cpu_count() or 8)
        r0, r1 = _rows_of_records(res, lo, hi)
        got = res["paf_text"][off[r0]:off[r1]]
        assert r1 - r0 == want["rows"] and got == want["lifted"], nm
        assert hostlib.HostPaf.from_text(got).stats_text(res["stats"], row0=r0) == want["stats"], nm  # fused per-row stats
        checked += r1 - r0
    assert checked > (150_000 if width == 1000 else 1_500)


def test_full_scale_stats_text_mode_equals_oracle_on_contigs(ctx, full):
    # RB_WANT_STATS_TEXT at C4's size: the rows `rb liftover | rb stats --paf` prints, formatted on the device, sliced pipeline
    wins = full.tiling_windows(1000)
    res = ctx.liftover(full, wins, want=capi.WANT_STATS_TEXT | capi.WANT_NUMERIC, stats=True)
    off = res["line_off"].astype(np.int64)
    assert res["n_out"] > 3_000_000 and off[-1] == len(res["paf_text"]) < 140 * res["n_out"]
    hdr = bamstats.print_cigar_stats_header().encode()
    for nm in PARITY_CONTIGS[1:]:
        tid = full.find_name(nm)
        (lo, hi), = _record_runs(full, tid)
        want = orc.bench_pipeline_keep(full.text(lo, hi), full.tiling_bed_text(1000, tid), threads=os.cpu_count() or 8)
        r0, r1 = _rows_of_records(res, lo, hi)
        assert hdr + res["paf_text"][off[r0]:off[r1]] == want["stats"], nm


def test_full_scale_stats_bytes_equal_oracle_on_contigs(ctx, full):
    # C2: `rb stats --paf` over the whole synthetic PAF; rows of the parity contigs against the oracle's printout
    st = ctx.stats(full)
    for nm in PARITY_CONTIGS:
        (lo, hi), = _record_runs(full, full.find_name(nm))
        text = full.text(lo, hi)
        assert hostlib.HostPaf.from_text(text).stats_text(st, row0=lo) == orc.run_stats(text), nm


def test_c5_shape_beyond_4gib_bytes_equal_oracle_on_contigs(ctx):
    # C5-shaped: 33 haplotypes concatenated haplotype-major (4.4 GB of CIGAR text: byte offsets cross 2^32 in the input and in
    # the output, slices are gathered in emission order, pinned ranges are split at GiB boundaries), 10 kb windows.
    # Compared with the oracle: chr21 of the first / a middle / the last haplotype and chrM of every haplotype.
    n_hap, width = 33, 10_000
    paf = hostlib.HostPaf.synth(scale=1.0, n_hap=n_hap, threads=os.cpu_count() or 8)
    assert paf.cigar_nbytes > 1 << 32
    lib = capi.load()
    addr = C.cast(paf.c.cigar, C.c_void_p).value
    pinned = lib.rb_host_register(C.c_void_p(addr), paf.cigar_nbytes) == 0
    try:
        wins = paf.tiling_windows(width)
        res = ctx.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    finally:
        if pinned:
            lib.rb_host_unregister(C.c_void_p(addr))
    off = res["line_off"].astype(np.int64)
    assert int(off[-1]) == len(res["paf_text"]) > 1 << 32 and (np.diff(off) > 0).all()
    checked = 0
    for nm, haps in (("chr21", (0, n_hap // 2, n_hap - 1)), ("chrM", range(n_hap))):
        tid = paf.find_name(nm)
        runs = _record_runs(paf, tid)
        assert len(runs) == n_hap
        bed_text = paf.tiling_bed_text(width, tid)
        for h in haps:
            lo, hi = runs[h]
            want = orc.bench_pipeline_keep(paf.text(lo, hi), bed_text, threads=os.cpu_count() or 8)
            r0, r1 = _rows_of_records(res, lo, hi)
            got = res["paf_text"][off[r0]:off[r1]]
            assert r1 - r0 == want["rows"] and got == want["lifted"], (nm, h)
            assert hostlib.HostPaf.from_text(got).stats_text(res["stats"], row0=r0) == want["stats"], (nm, h)
            checked += r1 - r0
    assert checked > 10_000
    # the last rows of the call (beyond 4 GiB of output text) belong to the last contig in emission order
    assert int(off[_rows_of_records(res, *_record_runs(paf, paf.find_name("chrM"))[-1])[0]]) > 1 << 32


@pytest.mark.parametrize("width", [1000, 10_000, 100_000])
def test_full_scale_streaming_path_equals_search_path(ctx, full, width):
    """k_scan_lift + k_combine (boundaries resolved while scanning) against k_samples + k_lift (one search per
    pair, the path the small-case oracle parity also pins): identical bytes, counters and mirror at full size."""
    wins = full.tiling_windows(width)
    a = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    ctx.set_lift_mode(capi.LIFT_STREAM)
    try:
        b = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    finally:
        ctx.set_lift_mode(capi.LIFT_SEARCH)
    assert a["n_out"] == b["n_out"] and a["n_pairs"] == b["n_pairs"]
    assert a["paf_text"] == b["paf_text"]
    for k in ("q_st", "q_en", "t_st", "t_en", "nmatch", "aln_len", "rec_idx", "win_idx"):
        assert (a[k] == b[k]).all(), k
    for k in ("equal", "diff", "ins", "del", "ins_events", "del_events", "matches"):
        assert (a["stats"][k] == b["stats"][k]).all(), k
    for k in ("id_by_matches", "id_by_events", "id_by_all"):
        assert (a["stats"][k].view(np.uint32) == b["stats"][k].view(np.uint32)).all(), k


def test_full_scale_sliced_call_equals_unsliced(ctx, full):
    """rb_liftover in overlapped slices (the default for large calls) returns the bytes of the single-batch call."""
    wins = full.tiling_windows(1000)
    a = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    ctx.set_slicing(0)
    try:
        b = ctx.liftover(full, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    finally:
        ctx.set_slicing()
    assert a["n_out"] == b["n_out"] and a["n_pairs"] == b["n_pairs"] and a["paf_text"] == b["paf_text"]
    for k in ("line_off", "q_st", "q_en", "t_st", "t_en", "nmatch", "aln_len", "rec_idx", "win_idx"):
        assert (a[k] == b[k]).all(), k
    for k in ("equal", "diff", "ins", "del", "ins_events", "del_events", "matches"):
        assert (a["stats"][k] == b["stats"][k]).all(), k
    for k in ("id_by_matches", "id_by_events", "id_by_all"):
        assert (a["stats"][k].view(np.uint32) == b["stats"][k].view(np.uint32)).all(), k
    # text only / numeric only
    c = ctx.liftover(full, wins, want=capi.WANT_TEXT, stats=False)
    assert c["paf_text"] == a["paf_text"]
    d = ctx.liftover(full, wins, want=capi.WANT_NUMERIC, stats=True)
    assert (d["t_st"] == a["t_st"]).all() and (d["stats"]["equal"] == a["stats"]["equal"]).all()


def test_full_scale_multi_device_context_equals_single(ctx, full):
    """rb_liftover / rb_stats on a multi-device rb_ctx (three contexts on GPU 0 standing in for three GPUs; with more GPUs
    visible the real ones are used): partition in C++ on the packed arrays, one host thread per device, rows merged in
    emission order into ONE output — the bytes, mirror and counters of the single-device call."""
    import torch
    n_gpu = torch.cuda.device_count()
    ids = list(range(n_gpu)) if n_gpu >= 2 else [0, 0, 0]
    multi = capi.Context(devices=ids)
    try:
        for paf, width in ((full, 1000), (hostlib.HostPaf.synth(scale=0.25, n_hap=3), 10_000)):  # (the second: haplotype-major file order)
            wins = paf.tiling_windows(width)
            a = ctx.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
            b = multi.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
            assert a["n_out"] == b["n_out"] and a["n_pairs"] == b["n_pairs"] and a["paf_text"] == b["paf_text"]
            for k in ("line_off", "q_st", "q_en", "t_st", "t_en", "nmatch", "aln_len", "rec_idx", "win_idx"):
                assert (a[k] == b[k]).all(), k
            for k in ("equal", "diff", "ins", "del", "ins_events", "del_events", "matches"):
                assert (a["stats"][k] == b["stats"][k]).all(), k
            for k in ("id_by_matches", "id_by_events", "id_by_all"):
                assert (a["stats"][k].view(np.uint32) == b["stats"][k].view(np.uint32)).all(), k
            s1, s2 = ctx.stats(paf), multi.stats(paf)
            assert all((s1[k] == s2[k]).all() for k in ("equal", "diff", "ins", "del", "ins_events", "del_events", "matches"))
            assert (s1["id_by_all"].view(np.uint32) == s2["id_by_all"].view(np.uint32)).all()
    finally:
        multi.close()


def test_two_haplotypes_gathered_slices_equal_unsliced(ctx):
    """A multi-haplotype PAF (file order: haplotype-major, so contigs interleave) is sliced in emission order with
    gathered uploads; rows, rec_idx and counters equal the single-batch call."""
    paf = hostlib.HostPaf.synth(scale=0.25, n_hap=3)
    wins = paf.tiling_windows(10_000)
    ctx.set_slicing(4 << 20)
    try:
        a = ctx.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
        ctx.set_slicing(0)
        b = ctx.liftover(paf, wins, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    finally:
        ctx.set_slicing()
    assert a["n_out"] == b["n_out"] > 100_000 and a["paf_text"] == b["paf_text"]
    for k in ("line_off", "q_st", "t_en", "rec_idx", "win_idx"):
        assert (a[k] == b[k]).all(), k
    assert (a["stats"]["equal"] == b["stats"]["equal"]).all()
    assert (np.diff(a["rec_idx"].astype(np.int64)) < 0).any()  # emission order is not file order here


def test_contig_shards_merge_to_the_single_call_output(ctx):
    """The multi-GPU recipe on one device: records split by target contig (LPT on CIGAR bytes), every shard lifted on
    its own, outputs concatenated in emission order == the output of the unsharded call (and of the oracle)."""
    from rustybam_b200 import shard
    paf = hostlib.HostPaf.synth(scale=0.02, n_hap=2)
    bed_text = paf.tiling_bed_text(10_000)
    whole = ctx.liftover(paf, paf.windows_from_bed_text(bed_text), want=capi.WANT_TEXT, stats=False)["paf_text"]
    weights = shard.contig_bytes(paf)
    bins, loads = shard.lpt_bins(weights, 3)
    outs = []
    for tids in bins:
        part = shard.take_contigs(paf, tids)
        outs.append(ctx.liftover(part, part.windows_from_bed_text(bed_text), want=capi.WANT_TEXT, stats=False)["paf_text"])
    order = []
    for ln in paf.text().splitlines():
        t = ln.split(b"\t")[5]
        if t not in order:
            order.append(t)
    assert shard.merge_outputs(order, outs) == whole
    assert whole == orc.run_liftover(paf.text(), bed_text, threads=8)


def test_rb_cli_stats_on_many_rows(ctx, full):
    """`rb liftover | rb stats --paf` at volume: 100 k lifted rows through the CLI (parallel host parse, GPU counters,
    parallel row formatting) equal the oracle's `rb stats --paf` of the same rows."""
    wins = full.tiling_windows(1000)
    res = ctx.liftover(full, wins, want=capi.WANT_TEXT, stats=False)
    sub = res["paf_text"][:int(res["line_off"][100_000])]
    rb = os.path.join(ROOT, "rustybam_b200", "rb")
    got = subprocess.run([rb, "stats", "--paf", "-"], input=sub, capture_output=True, check=True).stdout
    assert got == orc.run_stats(sub)


def test_full_scale_invert_properties(ctx, full):
    # rb invert over the whole ~50 M-op PAF: one row per record, in file order; the swap is an involution on the
    # 12 columns + CIGAR, and a fixed point of invert . invert is reached after one application (canonical spelling)
    res = ctx.invert(full, want=capi.WANT_TEXT | capi.WANT_NUMERIC)
    assert res["n_out"] == full.n_rec and res["rec_idx"].tolist() == list(range(full.n_rec))
    c = full.c
    n = full.n_rec
    assert res["q_st"].tolist() == [c.t_st[i] for i in range(n)] and res["t_en"].tolist() == [c.q_en[i] for i in range(n)]
    rec_stats = ctx.stats(full)
    assert (res["nmatch"] == rec_stats["equal"].astype(np.uint64) + rec_stats["diff"]).all()
    once = res["paf_text"]
    twice = ctx.invert(hostlib.HostPaf.from_text(once), want=capi.WANT_TEXT)["paf_text"]
    thrice = ctx.invert(hostlib.HostPaf.from_text(twice), want=capi.WANT_TEXT)["paf_text"]
    assert thrice == once and twice != once
    # twice == the input records (same 9 leading columns, mapq and CIGAR; nmatch / aln_len re-inferred)
    orig = full.text().split(b"\n")
    back = twice.split(b"\n")
    assert len(orig) == len(back)
    for a, b in zip(orig[:-1], back[:-1]):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        assert fa[:9] == fb[:9] and fa[11] == fb[11] and fa[-1] == fb[-1]
    # the stats of an inverted record are the record's with ins <-> del
    inv_stats = ctx.stats(hostlib.HostPaf.from_text(once))
    assert (inv_stats["ins"] == rec_stats["del"]).all() and (inv_stats["del_events"] == rec_stats["ins_events"]).all()
    assert (inv_stats["equal"] == rec_stats["equal"]).all()


def test_rb_cli_matches_oracle(tmp_path):
    rb = os.path.join(ROOT, "rustybam_b200", "rb")
    paf_gz = os.path.join(ROOT, "tests", "golden", "asm_small.paf.gz")
    bed = os.path.join(ROOT, "tests", "golden", "asm_small.bed")
    lifted = subprocess.run([rb, "-t", "4", "liftover", "--bed", bed, paf_gz], capture_output=True, check=True).stdout
    assert lifted == orc.run_liftover(orc.golden_paf(), orc.golden_bed())
    st = subprocess.run([rb, "stats", "--paf", "-"], input=lifted, capture_output=True, check=True).stdout
    assert st == orc.run_stats(lifted)
    big = subprocess.run([rb, "liftover", "--largest", "--bed", bed, paf_gz], capture_output=True, check=True).stdout
    assert big == orc.run_liftover(orc.golden_paf(), orc.golden_bed(), largest=True)
    broken = subprocess.run([rb, "break-paf", "--max-size", "100", paf_gz], capture_output=True, check=True).stdout
    assert broken == orc.run_break_paf(orc.golden_paf(), 100)
    inverted = subprocess.run([rb, "invert", paf_gz], capture_output=True, check=True).stdout
    assert inverted == orc.run_invert(orc.golden_paf())
    bad = tmp_path / "bad.paf"
    bad.write_bytes(b"Q\t10\t0\t8\t+\tT\t20\t0\t8\t0\t0\t60\tcg:Z:3D5=\n")
    r = subprocess.run([rb, "liftover", "--bed", bed, str(bad)], capture_output=True)
    assert r.returncode == 101 and r.stdout == b""


# ---------------------------------------------------------------- rb trim-paf at scale
def overlapped_on_query(text: bytes, shift=300) -> bytes:
    """The synthetic PAF's records tile their query without overlap; moving the i-th record of a query name down by
    300 * i bases makes every record overlap its predecessor by 101..300 bases (no containment)."""
    seen, out = {}, []
    for ln in text.split(b"\n"):
        if not ln:
            continue
        f = ln.split(b"\t", 4)
        i = seen.get(f[0], 0)
        seen[f[0]] = i + 1
        f[2], f[3] = str(int(f[2]) - shift * i).encode(), str(int(f[3]) - shift * i).encode()
        out.append(b"\t".join(f))
    return b"\n".join(out) + b"\n"


def test_trim_paf_synth_2pct_exact_parity(ctx):
    text = overlapped_on_query(hostlib.HostPaf.synth(scale=0.02).text())
    want = orc.run_trim_paf(text)
    hp = hostlib.HostPaf.from_text(text)
    res = ctx.trim_paf(hp, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    assert res["paf_text"] == want
    assert sum(a != b for a, b in zip(sorted(text.splitlines()), sorted(want.splitlines()))) > 100  # most records were cut
    out = Paf.from_text(want)
    assert (bamstats.print_cigar_stats_header() + bamstats.stats_rows(out, res["stats"])).encode() == orc.run_stats(want)


def test_full_scale_trim_paf_properties(ctx, full):
    # ~50 M ops, 761 records on 25 query names, every record overlapping its predecessor: structural invariants of the result
    text = overlapped_on_query(full.text())
    hp = hostlib.HostPaf.from_text(text)
    n = hp.n_rec
    res = ctx.trim_paf(hp, want=capi.WANT_TEXT | capi.WANT_NUMERIC, stats=True)
    assert res["n_out"] == n
    idx = res["rec_idx"].astype(np.int64)
    assert sorted(idx.tolist()) == list(range(n))
    c = hp.c
    q_st_in = np.array([c.q_st[i] for i in range(n)], dtype=np.uint64)[idx]
    q_en_in = np.array([c.q_en[i] for i in range(n)], dtype=np.uint64)[idx]
    t_st_in = np.array([c.t_st[i] for i in range(n)], dtype=np.uint64)[idx]
    t_en_in = np.array([c.t_en[i] for i in range(n)], dtype=np.uint64)[idx]
    q_id = np.array([c.q_id[i] for i in range(n)], dtype=np.int64)[idx]
    # every row is a sub-interval of its record, on both sequences
    assert (res["q_st"] >= q_st_in).all() and (res["q_en"] <= q_en_in).all() and (res["q_st"] < res["q_en"]).all()
    assert (res["t_st"] >= t_st_in).all() and (res["t_en"] <= t_en_in).all() and (res["t_st"] < res["t_en"]).all()
    # rows are grouped by query name (stable sort: file order inside a name), and no two neighbours overlap any more
    same = q_id[1:] == q_id[:-1]
    starts = np.flatnonzero(np.r_[True, ~same])
    assert len(starts) == len(set(q_id[starts].tolist())) == 25
    assert (res["q_en"][:-1][same] <= res["q_st"][1:][same]).all()
    assert ((q_en_in[:-1][same] > q_st_in[1:][same])).all()  # ... while every such pair did overlap in the input
    # at most the overlap was given up: what a cut removes from the two records together is bounded by their overlap + the
    # indel columns slid over (a few bases)
    st = res["stats"]
    assert (res["nmatch"] == st["equal"].astype(np.uint64) + st["diff"]).all()
    assert (res["aln_len"] == st["equal"].astype(np.uint64) + st["diff"] + st["ins"] + st["del"]).all()
    assert (res["q_en"] - res["q_st"] == st["equal"].astype(np.uint64) + st["diff"] + st["ins"]).all()
    assert (res["t_en"] - res["t_st"] == st["equal"].astype(np.uint64) + st["diff"] + st["del"]).all()
    # the output is a valid PAF (spans match the CIGARs: the loader's check_integrity passes) and a fixed point of the command
    again = ctx.trim_paf(hostlib.HostPaf.from_text(res["paf_text"]), want=capi.WANT_TEXT, stats=False)
    assert again["paf_text"] == res["paf_text"]
