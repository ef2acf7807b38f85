"""Host-side BED regions (mirror of src/bed.rs:14-45, 140-194; bio 1.6.0 bed::Reader semantics:
tab-delimited csv, '#' comment lines, rows whose field count differs from the first row are
skipped with a warning, Region.id = column 4 or "{chrom}:{st+1}-{en}")."""
import gzip

import numpy as np

from .capi import Windows


class Region:
    __slots__ = ("name", "st", "en", "id", "default_id")

    def __init__(self, name, st, en, id_, default_id=False):
        self.name, self.st, self.en, self.id, self.default_id = name, st, en, id_, default_id


def parse_bed_text(text: bytes):
    out, nf0 = [], None
    for line in text.replace(b"\r", b"\n").split(b"\n"):
        if not line or line[:1] == b"#":
            continue
        f = line.split(b"\t")
        if nf0 is None:
            nf0 = len(f)
        elif len(f) != nf0:
            continue
        if len(f) < 3 or not f[1].isdigit() or not f[2].isdigit():
            continue
        st, en = int(f[1]), int(f[2])
        if st > 0xFFFFFFFFFFFFFFFF or en > 0xFFFFFFFFFFFFFFFF:
            continue
        rid = f[3] if len(f) > 3 else f[0] + b":" + str(st + 1).encode() + b"-" + str(en).encode()
        out.append(Region(f[0], st, en, rid, default_id=len(f) <= 3))
    return out


def parse_bed(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        return parse_bed_text(f.read())


def pack_windows(rgns, name_index: dict) -> Windows:
    """BED rows on contigs absent from the PAF are ignored (Q5); bed_row keeps the FILE order."""
    keep = [(i, r) for i, r in enumerate(rgns) if r.name in name_index]
    t_id = np.array([name_index[r.name] for _, r in keep], dtype=np.uint32)
    st = np.array([r.st for _, r in keep], dtype=np.uint64)
    en = np.array([r.en for _, r in keep], dtype=np.uint64)
    # 3-column BED: ids stay on the device side (rb_windows.ids == NULL -> "{chrom}:{st+1}-{en}" is formatted by the GPU)
    ids = None if keep and all(r.default_id for _, r in keep) else [r.id for _, r in keep]
    return Windows(t_id, st, en, ids)
