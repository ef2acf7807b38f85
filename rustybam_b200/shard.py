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
