"""The multi-GPU path shards by target contig with no data-path collective; the host-side logic
(LPT partition, shard extraction, totals) is exercised here with world_size = 2 over gloo on CPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rustybam_b200 import hostlib, shard
    sh, info = shard.make_shard(rank, world, scale=0.004, threads=2)
    n = sh.n_rec
    t_id = np.ctypeslib.as_array(sh.c.t_id, shape=(n,))
    names = {bytes(np.ctypeslib.as_array(sh.c.names, shape=(int(sh.c.names_off[sh.c.n_names]),))[int(sh.c.names_off[t]):int(sh.c.names_off[t + 1])]) for t in set(t_id.tolist())}
    mine = torch.tensor([n, sh.cigar_nbytes], dtype=torch.int64)
    tot = mine.clone()
    dist.all_reduce(tot)                       # sizes only: the data path itself has no collective
    gathered = [None] * world
    dist.all_gather_object(gathered, sorted(names))
    if rank == 0:
        full = hostlib.HostPaf.synth(scale=0.004, n_hap=world, threads=2)
        q.put((tot.tolist(), [full.n_rec, full.cigar_nbytes], gathered, info["balance"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_everything_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, full, gathered, balance = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot == full                                   # every record lands on exactly one rank
    assert not (set(gathered[0]) & set(gathered[1]))     # contigs are disjoint
    assert len(set(gathered[0]) | set(gathered[1])) == 25
    assert balance < 1.2


def _merge_worker(rank, world, port, q):
    """Each rank lifts the records of its own contigs (here with the CPU oracle: no GPU in this test), rank 0 gathers the
    outputs and concatenates them in the reference's emission order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gen
    import orc
    from rustybam_b200 import shard
    paf_text, contigs = gen.random_paf(5, n_contigs=5, recs_per_contig=7)   # contigs interleave in the file
    bed_text = gen.tiling_bed(contigs, 13)
    lines = paf_text.splitlines(keepends=True)
    order = []
    for ln in lines:
        t = ln.split(b"\t")[5]
        if t not in order:
            order.append(t)
    weights = {t: sum(len(ln) for ln in lines if ln.split(b"\t")[5] == t) for t in order}
    bins, _ = shard.lpt_bins(weights, world)
    mine = b"".join(ln for ln in lines if ln.split(b"\t")[5] in bins[rank])
    out = orc.run_liftover(mine, bed_text) if mine else b""
    gathered = [None] * world
    dist.all_gather_object(gathered, out)      # host-side gather of the finished rows; no collective on the data path
    if rank == 0:
        q.put((shard.merge_outputs(order, gathered), orc.run_liftover(paf_text, bed_text)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_outputs_concatenate_in_emission_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, whole = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged == whole and len(whole) > 0


# a record set where query name X's result depends on name Y: C is contained when X's only pair (A, B) is cut, overlaps both
# cut records afterwards, and is trimmed only because Y's second pair forces another round (paf.rs:283-285 counts globally)
def _rec(q, st, en, t0, cg=None):
    n = en - st
    return f"{q}\t1000\t{st}\t{en}\t+\tT\t100000\t{t0}\t{t0 + n}\t0\t0\t60\tcg:Z:{cg or str(n) + '='}\n".encode()


COUPLED = (_rec("X", 0, 100, 1000, "75=25X") + _rec("Y", 0, 100, 4000) + _rec("X", 50, 150, 2000, "25X75=") + _rec("Y", 50, 150, 5000) +
           _rec("X", 60, 90, 3000) + _rec("Y", 120, 220, 6000))


def _trim_worker(rank, world, port, q):
    """`rb trim-paf` shards by QUERY NAME with one flag OR-ed over the ranks after every round: each rank steps the records of
    its own names (here with the CPU oracle: no GPU in this test), rank 0 gathers the rows and merges the name groups."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gen
    import orc
    from rustybam_b200 import shard

    def any_waiting(flag):
        t = torch.tensor([int(flag)], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the sub-command's one exchange step: one integer per round
        return bool(t.item())

    results = []
    for paf_text in (COUPLED, gen.random_trim_paf(21, n_names=9, recs_per_name=5, max_ops=40, lead_trail=False)):
        texts, balance = shard.shard_by_query(paf_text, world)
        steps = orc.TrimSteps(texts[rank], 2, 1, 1)
        rounds = shard.trim_rounds_lockstep(steps.round, any_waiting)
        out = steps.end(True)
        alone = orc.run_trim_paf(texts[rank], 2, 1, 1, True)  # what the rank would print without the exchange
        gathered, independent = [None] * world, [None] * world
        dist.all_gather_object(gathered, out)      # host-side gather of the finished rows
        dist.all_gather_object(independent, alone)
        results.append((shard.merge_trim_outputs(gathered), shard.merge_trim_outputs(independent), orc.run_trim_paf(paf_text, 2, 1, 1, True),
                        rounds, balance))
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_trim_paf_shards_by_query_name():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_trim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (merged, independent, whole, rounds, _), (merged2, _, whole2, _, balance2) = results
    assert merged == whole and rounds >= 2          # lockstep rounds: byte-identical to the single set
    assert independent != whole                      # ... while ranks that stop on their own are NOT (X keeps its contained C)
    assert merged2 == whole2 and len(whole2) > 0 and balance2 < 1.5


def test_lpt_bins_balance():
    from rustybam_b200 import shard
    lens = [248387328, 242696752, 201105948, 193574945, 182045439, 172126628, 160567428, 146259331, 150617247, 134758134, 135127769,
            133324548, 113566686, 101161492, 99753195, 96330374, 84276897, 80542538, 61707364, 66210255, 45090682, 51324926, 154259566,
            62460029, 16569]
    for n in (2, 4, 8):
        bins, loads = shard.lpt_bins(dict(enumerate(lens)), n)
        assert sorted(sum(bins, [])) == list(range(25))
        assert max(loads) / (sum(loads) / n) < 1.1
