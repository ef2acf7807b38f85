"""Host-side PAF record model (mirror of src/paf.rs:24-78, 346-430 for the test-suite / bench).

Text stays on the host: this module only splits columns and packs the cg:Z: payloads into the
SoA buffers of include/rbcuda.h.  CIGAR tokenising, integrity checks and everything downstream
run on the GPU behind the C ABI."""
import gzip
import re

import numpy as np

from .capi import Records


class ReferencePanic(Exception):
    """The reference would panic here (exit status 101)."""


_WS = b" \t\n\x0c\r"


def _find_tag(tok: bytes):
    # regex "(..):(.):(.*)" — unanchored, leftmost (paf.rs:21,387-390)
    for p in range(0, len(tok) - 4):
        if tok[p + 2] == 0x3A and tok[p + 4] == 0x3A:
            return tok[p:p + 2], tok[p + 5:]
    return None


def _parse_u64(t: bytes):
    if t[:1] == b"+":
        t = t[1:]
    if not t.isdigit() or int(t) > 0xFFFFFFFFFFFFFFFF:
        return None
    return int(t)


_CIGAR_RE = re.compile(rb"(?:[0-9]+[MIDNSHP=X])*\Z")


def _cigar_syntax_ok(cg: bytes) -> bool:
    """rust-htslib CigarString::try_from, syntax only: lengths fit u32; H only first / last; S only at the ends or next to them over H."""
    if not _CIGAR_RE.match(cg):
        return False
    ops = re.findall(rb"([0-9]+)([MIDNSHP=X])", cg)
    if any(int(n) > 0xFFFFFFFF for n, _ in ops):
        return False
    kinds = [o for _, o in ops]
    for i, o in enumerate(kinds):
        if o == b"H" and i not in (0, len(kinds) - 1):
            return False
        if o == b"S" and i not in (0, len(kinds) - 1) and kinds[i - 1] != b"H" and any(k != b"H" for k in kinds[i + 1:]):
            return False
    return True


class Paf:
    """Paf::from_file (paf.rs:62-78): parsed columns of every record, CIGAR payloads still text."""

    def __init__(self):
        self.cols = []      # (q_name, q_len, q_st, q_en, strand, t_name, t_len, t_st, t_en, mapq)
        self.cigars = []    # bytes per record
        self.skipped = 0

    @staticmethod
    def from_text(text: bytes) -> "Paf":
        paf = Paf()
        lines = text.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        for line in lines:
            t = line.split()  # bytes.split() == ASCII whitespace (plus \x0b, never present in PAF)
            if len(t) < 12:
                raise ReferencePanic("assertion failed: t.len() >= 12")
            cigar = b""
            for tok in t[12:]:
                m = _find_tag(tok)
                if m is None:
                    raise ReferencePanic("assertion failed: PAF_TAG.is_match(token)")
                if m[0] == b"cg" and cigar == b"":
                    cigar = m[1]
            nums = [_parse_u64(t[i]) for i in (1, 2, 3, 6, 7, 8, 9, 10, 11)]
            if any(v is None for v in nums) or len(t[4]) != 1:
                # the reference parses the cg tag BEFORE the numeric columns (paf.rs:386-399 vs 401-417): a malformed CIGAR on a
                # line that would be skipped still panics (kept records get this check on the GPU)
                if cigar and not _cigar_syntax_ok(cigar):
                    raise ReferencePanic("Unable to parse cigar string.")
                paf.skipped += 1  # "Unable to parse PAF record. Skipping line" (paf.rs:73)
                continue
            paf.cols.append((t[0], nums[0], nums[1], nums[2], t[4], t[5], nums[3], nums[4], nums[5], nums[8]))
            paf.cigars.append(cigar)
        return paf

    @staticmethod
    def from_file(path: str) -> "Paf":
        opener = gzip.open if path.endswith((".gz", ".bgz")) else open
        with opener(path, "rb") as f:
            return Paf.from_text(f.read())

    def __len__(self):
        return len(self.cols)

    def pack(self) -> Records:
        names, index = [], {}

        def nid(nm):
            if nm not in index:
                index[nm] = len(names)
                names.append(nm)
            return index[nm]

        n = len(self.cols)
        q_id = np.array([nid(c[0]) for c in self.cols], dtype=np.uint32)
        t_id = np.array([nid(c[5]) for c in self.cols], dtype=np.uint32)
        col = lambda i: np.array([c[i] for c in self.cols], dtype=np.uint64)
        blob = b"".join(self.cigars)
        off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            off[1:] = np.cumsum([len(c) for c in self.cigars], dtype=np.uint64)
        strand = np.frombuffer(b"".join(c[4] for c in self.cols), dtype=np.uint8) if n else np.zeros(0, np.uint8)
        recs = Records(np.frombuffer(blob, dtype=np.uint8), off, col(1), col(2), col(3), col(6), col(7), col(8), col(9), strand,
                       q_id, t_id, names)
        recs.name_index = index
        return recs
