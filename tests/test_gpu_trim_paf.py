"""`rb trim-paf` through the C ABI (rb_trim_paf) against the literal per-base oracle, byte-for-byte.
Needs a real B200 (pytest -m gpu)."""
import os
import subprocess

import numpy as np
import pytest

import gen
import orc
from rustybam_b200 import bamstats, capi, liftover
from rustybam_b200.capi import RbError
from rustybam_b200.paf import Paf, ReferencePanic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def check_trim_against_oracle(ctx, paf_text, scores=(1, 1, 1), remove_contained=False, policy=0):
    """GPU output == oracle output (or both panic); returns the text (None on a panic)."""
    try:
        want = orc.run_trim_paf(paf_text, *scores, remove_contained, policy=policy)
    except orc.ReferencePanic:
        with pytest.raises(ReferencePanic):
            liftover.run_trim_paf(ctx, paf_text, *scores, remove_contained, policy=policy)
        return None
    paf = Paf.from_text(paf_text)
    res = liftover.overlapping_paf_recs(ctx, paf, *scores, remove_contained, policy=policy, stats=True)
    assert res["paf_text"] == want
    rows = [ln.split(b"\t") for ln in want.splitlines()]
    assert res["n_out"] == len(rows)
    # numeric mirror and fused stats of the emitted rows
    assert [int(r[2]) for r in rows] == res["q_st"].tolist() and [int(r[3]) for r in rows] == res["q_en"].tolist()
    assert [int(r[7]) for r in rows] == res["t_st"].tolist() and [int(r[8]) for r in rows] == res["t_en"].tolist()
    assert [int(r[9]) for r in rows] == res["nmatch"].tolist() and [int(r[10]) for r in rows] == res["aln_len"].tolist()
    # rec_idx = the caller's record of every row; the rows are ordered by query name (stable)
    names = [ln.split(b"\t")[0] for ln in paf_text.splitlines()]
    idx = res["rec_idx"].tolist()
    assert [names[i] for i in idx] == [r[0] for r in rows]
    assert [names[i] for i in idx] == sorted(names[i] for i in idx) and len(set(idx)) == len(idx)
    out_paf = Paf.from_text(want)
    st = (bamstats.print_cigar_stats_header() + bamstats.stats_rows(out_paf, res["stats"])).encode()
    assert st == orc.run_stats(want)
    return want


def test_trim_reference_vectors(ctx):
    # trim_overlap.rs:137-170 (test_inversion_trimming): left / center on the reverse strand / right
    paf = (b"Q\t20\t0\t10\t+\tT\t20\t0\t10\t3\t9\t60\tcg:Z:7=1X2=\n"
           b"Q\t20\t4\t15\t-\tT\t20\t5\t16\t3\t9\t60\tcg:Z:3=1X3=1M1X2=\n"
           b"Q\t20\t10\t20\t+\tT\t20\t10\t20\t3\t9\t60\tcz:Z:10=\tcg:Z:2=2X2=2X2=\n")
    want = check_trim_against_oracle(ctx, paf)
    assert [ln.split(b"\t")[-1] for ln in want.splitlines()] == [b"cg:Z:7=", b"cg:Z:2=1X3=1M", b"cg:Z:2=2X2="]
    # trim_overlap.rs:22-34 (doctest) as a set: the right record is contained in the left one -> untouched, or removed with -r
    paf = b"Q\t10\t0\t10\t+\tT\t20\t0\t10\t3\t9\t60\tcg:Z:7=1X2=\nQ\t10\t5\t10\t-\tT\t20\t10\t15\t3\t9\t60\tcg:Z:3=1X1=\n"
    assert check_trim_against_oracle(ctx, paf).count(b"\n") == 2
    assert check_trim_against_oracle(ctx, paf, remove_contained=True).count(b"\n") == 1
    # the doctest's pair made partial (the right record reaches past the left one): both get cut at the best split
    paf = b"Q\t12\t0\t10\t+\tT\t20\t0\t10\t3\t9\t60\tcg:Z:7=1X2=\nQ\t12\t5\t11\t-\tT\t20\t10\t16\t3\t9\t60\tcg:Z:1=3=1X1=\n"
    check_trim_against_oracle(ctx, paf)
    assert liftover.run_trim_paf(ctx, b"") == b""
    one = b"Q\t10\t0\t10\t+\tT\t20\t0\t8\t3\t9\t60\tcg:Z:2I5=1X2=\n"  # a single record: stripped, not trimmed
    assert check_trim_against_oracle(ctx, one) == b"Q\t10\t2\t10\t+\tT\t20\t0\t8\t8\t8\t60\tid:Z:_TO.2I.\tcg:Z:5=1X2=\n"


@pytest.mark.parametrize("seed", range(24))
def test_trim_random(ctx, seed):
    kw = dict(style="all" if seed % 2 else "eqx", canonical=(seed % 4 < 2), allow_zero=(seed % 8 >= 6), big=(2 if seed % 5 == 0 else 0),
              n_names=3 + seed % 4, recs_per_name=3 + seed % 3, span=80 if seed % 3 else 2000)
    paf_text = gen.random_trim_paf(seed, **kw)
    n_ok = 0
    for scores, rc in (((1, 1, 1), False), ((2, 3, 1), True), ((1, 0, 0), False)):
        n_ok += check_trim_against_oracle(ctx, paf_text, scores, rc) is not None
        # the early-exit binary_search of Rust 1.52 ..= 1.81 (SURVEY Q2 in query space): scores of the positions in front of
        # deletions follow the record's current truncation, the device re-scans the records it cut (k_trim_rescan)
        check_trim_against_oracle(ctx, paf_text, scores, rc, policy=orc.EARLY_EXIT)
    assert n_ok or seed in ()  # (panics are rare; a seed where every score set panics would test nothing)


def test_trim_many_names_and_rounds(ctx):
    # hundreds of query names in one call (one block per selected pair and round), piles of records per name (many rounds)
    paf_text = gen.random_trim_paf(77, n_names=300, recs_per_name=3, max_ops=40)
    want = check_trim_against_oracle(ctx, paf_text)
    assert want is not None and want.count(b"\n") == paf_text.count(b"\n")
    check_trim_against_oracle(ctx, paf_text, policy=orc.EARLY_EXIT)
    pile = gen.random_trim_paf(78, n_names=2, recs_per_name=40, max_ops=30, span=600, lead_trail=False)
    check_trim_against_oracle(ctx, pile, (1, 1, 1), True, policy=orc.EARLY_EXIT)
    want = check_trim_against_oracle(ctx, pile, (1, 1, 1), True)
    if want is not None:
        assert liftover.run_trim_paf(ctx, want) == want  # nothing left to trim: a second pass changes nothing


def test_trim_long_records(ctx):
    # records of ~10^4-10^5 ops: many scan tiles per record, thousands of split-point candidates per pair and thread
    paf_text = gen.random_trim_paf(5, n_names=2, recs_per_name=4, max_ops=4000, span=30000, big=8)
    assert check_trim_against_oracle(ctx, paf_text) is not None
    check_trim_against_oracle(ctx, paf_text, (3, 2, 1), True)
    paf_text = gen.random_trim_paf(6, n_names=4, recs_per_name=3, max_ops=2000, span=20000, big=12)
    assert check_trim_against_oracle(ctx, paf_text, (1, 1, 1)) is not None
    check_trim_against_oracle(ctx, paf_text, (2, 1, 3), False, policy=orc.EARLY_EXIT)


def test_trim_bundled_fixture(ctx):
    # .test/asm_small.paf: 249 records on 5 query names (up to 84 per name) -> 166 trimmed pairs over 56 rounds
    paf_text = orc.golden_paf()
    want = check_trim_against_oracle(ctx, paf_text)
    assert want.count(b"\n") == 249
    dropped = check_trim_against_oracle(ctx, paf_text, (3, 2, 1), True)
    assert dropped.count(b"\n") == 241
    early = check_trim_against_oracle(ctx, paf_text, policy=orc.EARLY_EXIT)
    assert early.count(b"\n") == 249
    changed = sum(a.split(b"\t")[2:4] != b.split(b"\t")[2:4] for a, b in zip(sorted(paf_text.splitlines()), sorted(want.splitlines())))
    assert changed > 100


def test_trim_shards_by_query_name(ctx):
    # the multi-GPU form (shard.shard_by_query + rb_trim_paf_begin / _round / _end): whole query names per rank, one flag
    # OR-ed over the ranks after every round.  Three "ranks" = three contexts on this GPU, stepped in lockstep.
    from rustybam_b200 import shard
    from test_shard_gloo import COUPLED
    for paf_text in (COUPLED, gen.random_trim_paf(31, n_names=12, recs_per_name=5, max_ops=60, lead_trail=False)):
        want = orc.run_trim_paf(paf_text, 1, 1, 1, True)
        assert liftover.run_trim_paf(ctx, paf_text, 1, 1, 1, True) == want      # the single call counts waiting pairs globally
        texts, balance = shard.shard_by_query(paf_text, 3)
        ranks = [capi.Context(0) for _ in texts]
        try:
            for c, t in zip(ranks, texts):
                c.trim_paf_begin(Paf.from_text(t).pack(), 1, 1, 1)
            rounds = 1
            while any([c.trim_paf_round() for c in ranks]):   # (a list: every rank runs the round before the flags are OR-ed)
                rounds += 1
            outs = [c.trim_paf_end(True, want=capi.WANT_TEXT, stats=False)["paf_text"] for c in ranks]
        finally:
            for c in ranks:
                c.close()
        assert shard.merge_trim_outputs(outs) == want
        assert sum(t.count(b"\n") for t in texts) == paf_text.count(b"\n")
    # the coupled set: name X alone stops one round earlier and keeps its contained record untrimmed
    x_only = b"".join(ln for ln in COUPLED.splitlines(keepends=True) if ln.startswith(b"X"))
    alone = liftover.run_trim_paf(ctx, x_only, 1, 1, 1, False)
    assert alone == orc.run_trim_paf(x_only, 1, 1, 1, False)
    whole = liftover.run_trim_paf(ctx, COUPLED, 1, 1, 1, False)
    assert [ln for ln in whole.splitlines() if ln.startswith(b"X")] != alone.splitlines()


def test_trim_errors(ctx):
    paf = Paf.from_text(gen.random_trim_paf(3))
    with pytest.raises(RbError) as e:  # neither of the two binary_search policies
        ctx.trim_paf(paf.pack(), policy=7)
    assert e.value.code == capi.RB_ERR_BAD_ARG
    # a record that does not start on M/=/X after the strip is outside the documented domain: loud, not wrong
    odd = Paf.from_text(b"Q\t30\t0\t10\t+\tT\t40\t5\t20\t0\t0\t60\tcg:Z:5N5=5=\nQ\t30\t5\t15\t+\tT\t40\t50\t60\t0\t0\t60\tcg:Z:10=\n")
    with pytest.raises(RbError) as e:
        ctx.trim_paf(odd.pack())
    assert e.value.code == -8
    # reference panics at load are reported as such (paf.rs:70), and the context stays usable
    bad = b"Q\t30\t0\t13\t+\tT\t40\t5\t16\t0\t0\t60\tcg:Z:12=\n"
    with pytest.raises(orc.ReferencePanic):
        orc.run_trim_paf(bad)
    with pytest.raises(ReferencePanic):
        liftover.run_trim_paf(ctx, bad)
    check_trim_against_oracle(ctx, gen.random_trim_paf(4))


def test_trim_cli(tmp_path):
    rb = os.path.join(ROOT, "rustybam_b200", "rb")
    paf_gz = os.path.join(ROOT, "tests", "golden", "asm_small.paf.gz")
    got = subprocess.run([rb, "trim-paf", paf_gz], capture_output=True, check=True).stdout
    assert got == orc.run_trim_paf(orc.golden_paf())
    got = subprocess.run([rb, "tp", "-m", "3", "-d", "2", "-i", "1", "-r", paf_gz], capture_output=True, check=True).stdout
    assert got == orc.run_trim_paf(orc.golden_paf(), 3, 2, 1, True)
