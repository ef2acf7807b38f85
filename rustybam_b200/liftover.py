"""liftover::trim_paf_by_rgns (src/liftover.rs:134-167) + Display (src/paf.rs:923-943) on the GPU."""
from . import bed as _bed
from .capi import POLICY_RIGHTMOST, WANT_NUMERIC, WANT_QBED, WANT_TEXT, RbError, REF_PANIC_CODES
from .paf import Paf, ReferencePanic


def trim_paf_by_rgns(ctx, rgns, paf: Paf, invert_query=False, policy=POLICY_RIGHTMOST, stats=True, want=WANT_TEXT | WANT_NUMERIC):
    """Returns the dict of rb_lift_out arrays; res["paf_text"] is what `rb liftover` prints."""
    recs = paf.pack()
    wins = _bed.pack_windows(rgns, recs.name_index)  # with --qbed the BED names are QUERY names (same name table)
    try:
        return ctx.liftover(recs, wins, policy=policy, want=want | (WANT_QBED if invert_query else 0), stats=stats)
    except RbError as e:
        if e.code in REF_PANIC_CODES:
            raise ReferencePanic(str(e)) from e
        raise


def largest_rows(res) -> bytes:
    """`--largest` (main.rs:200-208): rows stably sorted by id; per id the LAST row of maximal target span."""
    text, off = res["paf_text"], res["line_off"]
    rows = []
    for i in range(res["n_out"]):
        line = text[int(off[i]):int(off[i + 1])]
        rows.append((line.split(b"\t")[12][5:], int(res["t_en"][i]) - int(res["t_st"][i]), line))
    rows.sort(key=lambda r: r[0])  # stable
    out, i = [], 0
    while i < len(rows):
        j, best = i, i
        while j < len(rows) and rows[j][0] == rows[i][0]:
            if rows[j][1] >= rows[best][1]:
                best = j
            j += 1
        out.append(rows[best][2])
        i = j
    return b"".join(out)


def run_liftover(ctx, paf_text: bytes, bed_text: bytes, policy=POLICY_RIGHTMOST, qbed=False, largest=False) -> bytes:
    """`rb liftover --bed BED [--qbed] [--largest] PAF` (main.rs:186-214): stdout bytes."""
    rgns = _bed.parse_bed_text(bed_text)
    paf = Paf.from_text(paf_text)
    res = trim_paf_by_rgns(ctx, rgns, paf, invert_query=qbed, policy=policy, stats=False,
                           want=WANT_TEXT | (WANT_NUMERIC if largest else 0))
    return largest_rows(res) if largest else res["paf_text"]


def break_paf_on_indels(ctx, paf: Paf, max_size=100, policy=POLICY_RIGHTMOST, stats=True, want=WANT_TEXT | WANT_NUMERIC):
    """liftover::break_paf_on_indels for every record (src/liftover.rs:182-226, driver src/main.rs:271-281) on the GPU:
    res["paf_text"] is what `rb break-paf --max-size N` prints."""
    try:
        return ctx.break_paf(paf.pack(), max_size=max_size, policy=policy, want=want, stats=stats)
    except RbError as e:
        if e.code in REF_PANIC_CODES:
            raise ReferencePanic(str(e)) from e
        raise


def paf_swap_query_and_target(ctx, paf: Paf, want=WANT_TEXT | WANT_NUMERIC):
    """paf::paf_swap_query_and_target for every record (src/paf.rs:1050-1094, driver src/main.rs:176-182) on the GPU:
    res["paf_text"] is what `rb invert` prints."""
    try:
        return ctx.invert(paf.pack(), want=want)
    except RbError as e:
        if e.code in REF_PANIC_CODES:
            raise ReferencePanic(str(e)) from e
        raise


def overlapping_paf_recs(ctx, paf: Paf, match_score=1, diff_score=1, indel_score=1, remove_contained=False, policy=POLICY_RIGHTMOST,
                         stats=False, want=WANT_TEXT | WANT_NUMERIC):
    """Paf::overlapping_paf_recs + the print loop of `rb trim-paf` (src/paf.rs:210-305, src/trim_overlap.rs:36-86, driver
    src/main.rs:218-230) on the GPU: res["paf_text"] is what `rb trim-paf` prints (rows ordered by query name)."""
    try:
        return ctx.trim_paf(paf.pack(), match_score, diff_score, indel_score, remove_contained, policy=policy, want=want, stats=stats)
    except RbError as e:
        if e.code in REF_PANIC_CODES:
            raise ReferencePanic(str(e)) from e
        raise


def run_trim_paf(ctx, paf_text: bytes, match_score=1, diff_score=1, indel_score=1, remove_contained=False, policy=POLICY_RIGHTMOST) -> bytes:
    """`rb trim-paf [-m M] [-d D] [-i I] [-r] PAF`: stdout bytes."""
    return overlapping_paf_recs(ctx, Paf.from_text(paf_text), match_score, diff_score, indel_score, remove_contained, policy=policy,
                                want=WANT_TEXT)["paf_text"]


def run_invert(ctx, paf_text: bytes) -> bytes:
    """`rb invert PAF`: stdout bytes."""
    return paf_swap_query_and_target(ctx, Paf.from_text(paf_text), want=WANT_TEXT)["paf_text"]


def run_break_paf(ctx, paf_text: bytes, max_size=100, policy=POLICY_RIGHTMOST) -> bytes:
    """`rb break-paf --max-size N PAF`: stdout bytes."""
    return break_paf_on_indels(ctx, Paf.from_text(paf_text), max_size, policy, stats=False, want=WANT_TEXT)["paf_text"]
