"""`rb stats --paf` (src/bamstats.rs:91-154, 225-270; src/main.rs:50-58): counters and f32 identities
come from the GPU, this module only prints them the way the reference does."""
import numpy as np

from .capi import RbError, REF_PANIC_CODES
from .paf import Paf, ReferencePanic


def fmt_f32(v) -> str:
    """Rust `{}` for f32: shortest round-trip digits, positional, no trailing '.0'; an exact tie between the two shortest
    candidates rounds UP (flt2dec's Dragon rule) where numpy / printf round to even — so this goes through the C++ host's
    formatter (host/f32_fast.hpp), e.g. 16.0078125 -> "16.007813"."""
    from . import hostlib
    return hostlib.fmt_f32(float(np.float32(v)))


def print_cigar_stats_header(qbed=False) -> str:
    if qbed:
        s = "#query_name\tquery_start\tquery_end\tquery_length\tstrand\treference_name\treference_start\treference_end\treference_length\t"
    else:
        s = "#reference_name\treference_start\treference_end\treference_length\tstrand\tquery_name\tquery_start\tquery_end\tquery_length\t"
    return s + "perID_by_matches\tperID_by_events\tperID_by_all\tmatches\tmismatches\tdeletion_events\tinsertion_events\tdeletions\tinsertions\n"


def stats_rows(paf: Paf, st: dict, qbed=False) -> str:
    rows = []
    for i, c in enumerate(paf.cols):
        q = f"{c[0].decode()}\t{c[2]}\t{c[3]}\t{c[1]}\t"
        r = f"{c[5].decode()}\t{c[7]}\t{c[8]}\t{c[6]}\t"
        lead = (q + c[4].decode() + "\t" + r) if qbed else (r + c[4].decode() + "\t" + q)
        rows.append(lead + "\t".join([fmt_f32(st["id_by_matches"][i]), fmt_f32(st["id_by_events"][i]), fmt_f32(st["id_by_all"][i]),
                                      str(st["equal"][i]), str(st["diff"][i]), str(st["del_events"][i]), str(st["ins_events"][i]),
                                      str(st["del"][i]), str(st["ins"][i])]) + "\n")
    return "".join(rows)


def stats_from_paf(ctx, paf: Paf) -> dict:
    try:
        return ctx.stats(paf.pack())
    except RbError as e:
        if e.code in REF_PANIC_CODES:
            raise ReferencePanic(str(e)) from e
        raise


def run_stats(ctx, paf_text: bytes, qbed=False) -> bytes:
    """`rb stats --paf FILE`: stdout bytes (header first, main.rs:51)."""
    paf = Paf.from_text(paf_text)
    st = stats_from_paf(ctx, paf)
    return (print_cigar_stats_header(qbed) + stats_rows(paf, st, qbed)).encode()
