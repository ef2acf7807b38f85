"""ctypes face of the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import gzip
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
RIGHTMOST, EARLY_EXIT = 0, 1


class ReferencePanic(Exception):
    """The reference would panic (exit status 101) on this input."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liborc.so"))
        _lib.orc_free.argtypes = [C.c_void_p]
    return _lib


def _take(p, n):
    s = C.string_at(p, n.value)
    lib().orc_free(p)
    return s


def run_liftover(paf: bytes, bed: bytes, qbed=False, largest=False, policy=RIGHTMOST, threads=1) -> bytes:
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_run_liftover(paf, C.c_size_t(len(paf)), bed, C.c_size_t(len(bed)), int(qbed), int(largest),
                                policy, threads, C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0
    return _take(out, n)


def run_break_paf(paf: bytes, max_size=100, policy=RIGHTMOST) -> bytes:
    """`rb break-paf --max-size N` (main.rs:271-281) on PAF text."""
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_run_break_paf(paf, C.c_size_t(len(paf)), C.c_uint32(max_size), policy, C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0
    return _take(out, n)


def run_invert(paf: bytes) -> bytes:
    """`rb invert` (main.rs:176-182) on PAF text."""
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_run_invert(paf, C.c_size_t(len(paf)), C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0
    return _take(out, n)


def run_trim_paf(paf: bytes, match_score=1, diff_score=1, indel_score=1, remove_contained=False, policy=RIGHTMOST) -> bytes:
    """`rb trim-paf` (main.rs:218-230, paf.rs:210-305, trim_overlap.rs:36-86) on PAF text."""
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_run_trim_paf(paf, C.c_size_t(len(paf)), int(match_score), int(diff_score), int(indel_score), int(remove_contained),
                                policy, C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0
    return _take(out, n)


class TrimSteps:
    """`rb trim-paf` one recursion level at a time (paf.rs:210-287): begin -> round() until nobody waits -> end()."""

    def __init__(self, paf: bytes, match_score=1, diff_score=1, indel_score=1, policy=RIGHTMOST):
        err = C.create_string_buffer(512)
        lib().orc_trim_begin.restype = C.c_void_p
        self.h = lib().orc_trim_begin(paf, C.c_size_t(len(paf)), err, C.c_size_t(512))
        if not self.h:
            raise ReferencePanic(err.value.decode())
        self.args = (int(match_score), int(diff_score), int(indel_score), policy)

    def round(self) -> bool:
        w, err = C.c_int(), C.create_string_buffer(512)
        rc = lib().orc_trim_round(C.c_void_p(self.h), *self.args, C.byref(w), err, C.c_size_t(512))
        if rc == 101:
            raise ReferencePanic(err.value.decode())
        return bool(w.value)

    def end(self, remove_contained=False) -> bytes:
        out, n = C.c_void_p(), C.c_size_t()
        lib().orc_trim_end(C.c_void_p(self.h), int(remove_contained), C.byref(out), C.byref(n))
        self.h = None
        return _take(out, n)


def trim_pair(left: str, right: str, match_score=1, diff_score=1, indel_score=1, policy=RIGHTMOST):
    """aligned_pairs on both + trim_overlapping_pafs (trim_overlap.rs:36-86); returns the two output lines."""
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_trim_pair(left.encode(), right.encode(), int(match_score), int(diff_score), int(indel_score), policy,
                             C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0, rc
    return _take(out, n).decode().splitlines()


def run_stats(paf: bytes, qbed=False) -> bytes:
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_run_stats(paf, C.c_size_t(len(paf)), int(qbed), C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    assert rc == 0
    return _take(out, n)


def bench_pipeline(paf: bytes, bed: bytes, policy=RIGHTMOST, threads=8):
    sl, ss, rows, ob = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
    err = C.create_string_buffer(512)
    rc = lib().orc_bench_pipeline(paf, C.c_size_t(len(paf)), bed, C.c_size_t(len(bed)), policy, threads,
                                  C.byref(sl), C.byref(ss), C.byref(rows), C.byref(ob), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    return dict(secs_liftover=sl.value, secs_stats=ss.value, rows=rows.value, out_bytes=ob.value)


def bench_pipeline_keep(paf: bytes, bed: bytes, policy=RIGHTMOST, threads=8):
    """bench_pipeline that also returns what `rb liftover` and `rb stats --paf` print (for byte comparisons)."""
    sl, ss, rows = C.c_double(), C.c_double(), C.c_uint64()
    lo, ln, so, sn = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
    err = C.create_string_buffer(512)
    rc = lib().orc_bench_pipeline_keep(paf, C.c_size_t(len(paf)), bed, C.c_size_t(len(bed)), policy, threads, C.byref(sl), C.byref(ss),
                                       C.byref(rows), C.byref(lo), C.byref(ln), C.byref(so), C.byref(sn), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    return dict(secs_liftover=sl.value, secs_stats=ss.value, rows=rows.value, lifted=_take(lo, ln), stats=_take(so, sn))


def trim_line(line: str, name: str, st: int, en: int, rid: str = "", policy=RIGHTMOST):
    """aligned_pairs + trim_paf_rec_to_rgn on one record line; returns the output line or None."""
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_trim_line(line.encode(), name.encode(), C.c_uint64(st), C.c_uint64(en), rid.encode(), policy,
                             C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    if rc == 1:
        return None
    assert rc == 0, rc
    return _take(out, n).decode()


def break_paf(line: str, break_length: int, policy=RIGHTMOST):
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_break_paf(line.encode(), C.c_uint32(break_length), policy, C.byref(out), C.byref(n), err,
                             C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    return _take(out, n).decode().splitlines()


def aligned_pairs_line(line: str) -> str:
    out, n, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(512)
    rc = lib().orc_aligned_pairs_cigar(line.encode(), C.byref(out), C.byref(n), err, C.c_size_t(512))
    if rc == 101:
        raise ReferencePanic(err.value.decode())
    return _take(out, n).decode()


def parse_cigar(s: str):
    cap = len(s) + 1
    lens, ops, n = (C.c_uint32 * cap)(), (C.c_uint8 * cap)(), C.c_size_t()
    rc = lib().orc_parse_cigar(s.encode(), C.c_size_t(len(s)), lens, ops, C.c_size_t(cap), C.byref(n))
    if rc == 101:
        raise ReferencePanic(s[:40])
    return [(lens[i], chr(ops[i])) for i in range(n.value)]


def cigar_stats(s: str):
    counts, ids = (C.c_uint32 * 7)(), (C.c_float * 3)()
    rc = lib().orc_cigar_stats(s.encode(), C.c_size_t(len(s)), counts, ids)
    if rc == 101:
        raise ReferencePanic(s[:40])
    keys = ["equal", "diff", "ins", "del", "matches", "ins_events", "del_events"]
    d = dict(zip(keys, list(counts)))
    d.update(id_by_matches=ids[0], id_by_events=ids[1], id_by_all=ids[2])
    return d


def fmt_f32(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_fmt_f32(C.c_float(v), buf, C.c_size_t(64))
    return buf.value.decode()


def parse_bed(bed: bytes):
    out, n = C.c_void_p(), C.c_size_t()
    lib().orc_parse_bed(bed, C.c_size_t(len(bed)), C.byref(out), C.byref(n))
    rows = []
    for ln in _take(out, n).decode().splitlines():
        a = ln.split("\t")
        rows.append((a[0], int(a[1]), int(a[2]), a[3]))
    return rows


def golden_paf() -> bytes:
    with gzip.open(os.path.join(GOLDEN, "asm_small.paf.gz"), "rb") as f:
        return f.read()


def golden_bed() -> bytes:
    with open(os.path.join(GOLDEN, "asm_small.bed"), "rb") as f:
        return f.read()
