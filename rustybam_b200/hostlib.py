"""ctypes binding of librbhost.so (C++ host helpers: packed PAF/BED builders, synthetic generator).

The objects returned here own host buffers laid out exactly as rb_records / rb_windows expect;
`.c` is the C struct to pass to capi.Context calls."""
import ctypes as C
import os

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librbhost.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m rustybam_b200.build`")
    lib = C.CDLL(LIB_PATH)
    lib.rbh_synth_paf.restype = C.c_void_p
    lib.rbh_synth_paf.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_int]
    lib.rbh_synth_paf_mask.restype = C.c_void_p
    lib.rbh_synth_paf_mask.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_uint32]
    lib.rbh_paf_from_text.restype = C.c_void_p
    lib.rbh_paf_from_text.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.rbh_paf_from_file.restype = C.c_void_p
    lib.rbh_paf_from_file.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    lib.rbh_paf_view.argtypes = [C.c_void_p, C.POINTER(capi.RbRecords)]
    lib.rbh_paf_size.restype = C.c_uint64
    lib.rbh_paf_size.argtypes = [C.c_void_p]
    lib.rbh_paf_skipped.restype = C.c_uint64
    lib.rbh_paf_skipped.argtypes = [C.c_void_p]
    lib.rbh_paf_free.argtypes = [C.c_void_p]
    lib.rbh_paf_find_name.restype = C.c_int64
    lib.rbh_paf_find_name.argtypes = [C.c_void_p, C.c_char_p]
    lib.rbh_paf_text.restype = C.c_void_p
    lib.rbh_paf_text.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_size_t)]
    lib.rbh_paf_text_of_contig.restype = C.c_void_p
    lib.rbh_paf_text_of_contig.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    lib.rbh_tiling_windows.restype = C.c_void_p
    lib.rbh_tiling_windows.argtypes = [C.c_void_p, C.c_uint64]
    lib.rbh_windows_from_bed_text.restype = C.c_void_p
    lib.rbh_windows_from_bed_text.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.rbh_windows_from_bed_text_slow.restype = C.c_void_p
    lib.rbh_windows_from_bed_text_slow.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.rbh_windows_view.argtypes = [C.c_void_p, C.POINTER(capi.RbWindows)]
    lib.rbh_windows_free.argtypes = [C.c_void_p]
    lib.rbh_tiling_bed_text.restype = C.c_void_p
    lib.rbh_tiling_bed_text.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(C.c_size_t)]
    lib.rbh_read_all.restype = C.c_void_p
    lib.rbh_read_all.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
    lib.rbh_free_str.argtypes = [C.c_void_p]
    lib.rbh_fmt_f32.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
    lib.rbh_stats_text.restype = C.c_void_p
    lib.rbh_stats_text.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    _lib = lib
    return lib


def _take(p, n):
    s = C.string_at(p, n.value)
    load().rbh_free_str(p)
    return s


def read_all(path: str) -> bytes:
    """File contents through the host reader (plain, .gz, .bgz — BGZF blocks are inflated on all host threads)."""
    n = C.c_size_t()
    p = load().rbh_read_all(path.encode(), C.byref(n))
    if not p:
        raise OSError(f"cannot read {path}")
    return _take(p, n)


class HostPanic(Exception):
    """The reference's parser would panic on this text."""


class HostPaf:
    """Packed records owned by the C++ host (rbh::Paf)."""

    def __init__(self, handle):
        self.lib = load()
        self.h = handle
        self.c = capi.RbRecords()
        self.lib.rbh_paf_view(self.h, C.byref(self.c))
        self.n_rec = int(self.c.n_rec)
        self.cigar_nbytes = int(self.c.cigar_nbytes)

    @staticmethod
    def synth(seed=20261017, scale=1.0, n_hap=1, threads=8, contigs=None):
        """contigs: indices into the CHM13-like table (0 = chr1 .. 21 = chr22, 22 = chrX, 23 = chrY, 24 = chrM) to generate; the
        streams are independent per (haplotype, contig), so a subset equals those records of the full set."""
        if contigs is None:
            return HostPaf(load().rbh_synth_paf(seed, scale, n_hap, threads))
        mask = 0
        for c in contigs:
            mask |= 1 << int(c)
        return HostPaf(load().rbh_synth_paf_mask(seed, scale, n_hap, threads, mask))

    @staticmethod
    def from_text(text: bytes):
        err = C.create_string_buffer(256)
        h = load().rbh_paf_from_text(text, len(text), err, 256)
        if not h:
            raise HostPanic(err.value.decode())
        return HostPaf(h)

    @staticmethod
    def from_file(path: str):
        """Paf::from_file (paf.rs:62-78): what the `rb` CLI does with its input argument."""
        err = C.create_string_buffer(256)
        h = load().rbh_paf_from_file(path.encode(), err, 256)
        if not h:
            raise HostPanic(err.value.decode())
        return HostPaf(h)

    @property
    def skipped(self):
        return int(self.lib.rbh_paf_skipped(self.h))

    def find_name(self, name: str) -> int:
        return int(self.lib.rbh_paf_find_name(self.h, name.encode()))

    def text(self, lo=0, hi=None) -> bytes:
        n = C.c_size_t()
        return _take(self.lib.rbh_paf_text(self.h, lo, self.n_rec if hi is None else hi, C.byref(n)), n)

    def text_of_contig(self, tid: int):
        n, cnt = C.c_size_t(), C.c_uint64()
        s = _take(self.lib.rbh_paf_text_of_contig(self.h, tid, C.byref(n), C.byref(cnt)), n)
        return s, int(cnt.value)

    def tiling_bed_text(self, width: int, tid: int = -1) -> bytes:
        n = C.c_size_t()
        return _take(self.lib.rbh_tiling_bed_text(self.h, width, tid, C.byref(n)), n)

    def tiling_windows(self, width: int) -> "HostWindows":
        return HostWindows(self.lib.rbh_tiling_windows(self.h, width))

    def windows_from_bed_text(self, bed: bytes) -> "HostWindows":
        return HostWindows(self.lib.rbh_windows_from_bed_text(self.h, bed, len(bed)))

    def stats_text(self, st: dict, row0=0, qbed=False, header=True) -> bytes:
        """What `rb stats --paf` prints for these records given the GPU counters `st` (rows row0 .. row0 + n_rec of its arrays)."""
        import numpy as np
        cols = [np.ascontiguousarray(st[k], dtype=np.uint32) for k in ("equal", "diff", "ins", "del", "ins_events", "del_events", "matches")]
        ids = [np.ascontiguousarray(st[k], dtype=np.float32) for k in ("id_by_matches", "id_by_events", "id_by_all")]
        assert all(len(c) >= row0 + self.n_rec for c in cols + ids)
        pc = (C.c_void_p * 7)(*[c.ctypes.data for c in cols])
        pi = (C.c_void_p * 3)(*[c.ctypes.data for c in ids])
        n = C.c_size_t()
        return _take(self.lib.rbh_stats_text(self.h, pc, pi, row0, int(qbed), int(header), C.byref(n)), n)

    def windows_from_bed_text_slow(self, bed: bytes) -> "HostWindows":
        """The two-step form (one Region with two strings per row, then pack): what windows_from_bed_text must equal."""
        return HostWindows(self.lib.rbh_windows_from_bed_text_slow(self.h, bed, len(bed)))

    def close(self):
        if self.h:
            self.lib.rbh_paf_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostWindows:
    def __init__(self, handle):
        self.lib = load()
        self.h = handle
        self.c = capi.RbWindows()
        self.lib.rbh_windows_view(self.h, C.byref(self.c))
        self.n_win = int(self.c.n_win)

    def close(self):
        if self.h:
            self.lib.rbh_windows_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fmt_f32(v: float) -> str:
    buf = C.create_string_buffer(64)
    load().rbh_fmt_f32(C.c_float(v), buf, 64)
    return buf.value.decode()
