"""Multi-GPU sharding of the liftover + stats path: records are partitioned by TARGET
CONTIG, balanced by longest-processing-time-first on CIGAR bytes; every GPU runs the whole kernel
sequence on its own contigs; there is no collective on the data path (a (window, record) pair only
needs that record's CIGAR and that window — liftover.rs:155-164 already treats contigs independently).
Outputs are concatenated on the host in emission order (contig first-appearance, record, BED row)."""
import numpy as np

from . import hostlib


def lpt_bins(weights, n_bins):
    """weights: {key: weight}.  Returns (bins: list[list[key]], loads)."""
    loads, bins = [0] * n_bins, [[] for _ in range(n_bins)]
    for k in sorted(weights, key=lambda k: (-weights[k], k)):
        b = loads.index(min(loads))
        bins[b].append(k)
        loads[b] += weights[k]
    return bins, loads


def contig_bytes(paf: "hostlib.HostPaf"):
    """CIGAR bytes per target name id."""
    n = paf.n_rec
    t_id = np.ctypeslib.as_array(paf.c.t_id, shape=(n,))
    off = np.ctypeslib.as_array(paf.c.cigar_off, shape=(n + 1,))
    sizes = (off[1:] - off[:-1]).astype(np.int64)
    out = {}
    for t in np.unique(t_id):
        out[int(t)] = int(sizes[t_id == t].sum())
    return out


def take_contigs(paf: "hostlib.HostPaf", tids) -> "hostlib.HostPaf":
    """A new packed PAF holding only the records whose target name id is in `tids` (file order kept)."""
    n = paf.n_rec
    t_id = np.ctypeslib.as_array(paf.c.t_id, shape=(n,))
    keep = np.nonzero(np.isin(t_id, np.fromiter(tids, dtype=np.uint32)))[0]
    parts, i = [], 0
    while i < len(keep):  # contiguous runs -> text -> re-pack (host side, never inside a timed region)
        j = i
        while j + 1 < len(keep) and keep[j + 1] == keep[j] + 1:
            j += 1
        parts.append(paf.text(int(keep[i]), int(keep[j]) + 1))
        i = j + 1
    return hostlib.HostPaf.from_text(b"".join(parts))


def make_shard(rank, world, scale=1.0, seed=20261017, threads=8):
    """The bench workload of one rank: `world` haplotypes vs the CHM13-like reference, contigs of LPT bin `rank`."""
    full = hostlib.HostPaf.synth(seed=seed, scale=scale, n_hap=world, threads=threads)
    if world == 1:
        return full, {"bins": 1, "balance": 1.0}
    w = contig_bytes(full)
    bins, loads = lpt_bins(w, world)
    shard = take_contigs(full, bins[rank])
    full.close()
    return shard, {"bins": world, "balance": max(loads) / (sum(loads) / world), "contigs": len(bins[rank])}


def split_by_target(paf_text: bytes):
    """Rows of one rank's `rb liftover` output grouped by target name: {t_name: bytes}.  A rank emits its contigs one after
    the other (liftover.rs:155-164), so each target is one contiguous run of lines."""
    out, order = {}, []
    start, cur = 0, None
    pos = 0
    n = len(paf_text)
    while pos < n:
        end = paf_text.index(b"\n", pos) + 1
        f = paf_text[pos:end].split(b"\t", 6)
        t = f[5]
        if t != cur:
            if cur is not None:
                out[cur] = out.get(cur, b"") + paf_text[start:pos]
            if t not in out:
                order.append(t)
            cur, start = t, pos
        pos = end
    if cur is not None:
        out[cur] = out.get(cur, b"") + paf_text[start:n]
    return out


def merge_outputs(contig_order, per_rank_text):
    """Concatenates the ranks' outputs in the reference's emission order: contigs by first appearance of t_name in the
    whole PAF (liftover.rs:151), each contig's rows taken from the rank that owns it.  contig_order: list of t_name bytes."""
    by_target = {}
    for text in per_rank_text:
        for t, rows in split_by_target(text).items():
            assert t not in by_target, "a target contig belongs to exactly one rank"
            by_target[t] = rows
    return b"".join(by_target.get(t, b"") for t in contig_order)


# ---- rb trim-paf: records interact only with records of the SAME QUERY NAME (paf.rs:229-239), so the set
# shards by query name — LPT on CIGAR bytes again.  But the reference's decision to run another round is GLOBAL (`if unseen > 0`
# counts the waiting pairs of all names, paf.rs:283-285) and a name's result depends on it (a record that was contained when its
# name's last pair was cut may overlap the cut records afterwards; it is only trimmed if another name forces one more round).
# So this sub-command has one real exchange step: after every round the ranks OR one flag.  Every rank's rows come out ordered
# by query name and the host merges the name groups back into the reference's order (stable sort by q_name, paf.rs:224).
def split_by_query(paf_text: bytes):
    """Lines of a PAF grouped by query name, file order kept inside a name: {q_name: bytes}."""
    out = {}
    for ln in paf_text.splitlines(keepends=True):
        if ln.strip():
            out.setdefault(ln.split(None, 1)[0], []).append(ln)
    return {k: b"".join(v) for k, v in out.items()}


def shard_by_query(paf_text: bytes, world: int):
    """Per-rank PAF texts for `rb trim-paf`: whole query names, balanced by bytes.  Returns (texts, balance)."""
    groups = split_by_query(paf_text)
    bins, loads = lpt_bins({k: len(v) for k, v in groups.items()}, world)
    owner = {k: r for r, b in enumerate(bins) for k in b}
    texts = [[] for _ in range(world)]
    for ln in paf_text.splitlines(keepends=True):  # file order is kept inside every rank (it breaks ties of the stable sort)
        if ln.strip():
            texts[owner[ln.split(None, 1)[0]]].append(ln)
    mean = sum(loads) / world if world else 0
    return [b"".join(t) for t in texts], (max(loads) / mean if mean else 1.0)


def trim_rounds_lockstep(round_fn, any_waiting):
    """The round loop of one rank: `round_fn()` runs one round on this rank's names and says whether one of its pairs had to
    wait; `any_waiting(flag)` ORs the flag over all ranks (torch.distributed all_reduce MAX; identity on one rank).  Every
    rank runs the same number of rounds — the whole set's."""
    rounds = 1
    while any_waiting(round_fn()):
        rounds += 1
    return rounds


def merge_trim_outputs(per_rank_text):
    """Concatenates the ranks' `rb trim-paf` outputs in the reference's order: query names ascending (byte-wise); a name's
    rows all come from the rank that owns it, already in their final order."""
    by_query = {}
    for text in per_rank_text:
        for q, rows in split_by_query(text).items():
            assert q not in by_query, "a query name belongs to exactly one rank"
            by_query[q] = rows
    return b"".join(by_query[q] for q in sorted(by_query))
