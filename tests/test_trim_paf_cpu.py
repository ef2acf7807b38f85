"""`rb trim-paf` on the CPU side: the literal oracle against the reference's own vectors, and the product's closed-form
query-space core (trim_core.cuh + trim_rounds.hpp — the code the GPU kernels run) fuzzed against that oracle."""
import os
import subprocess

import pytest

import gen
import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_trim_overlapping_pafs_doctest():
    # trim_overlap.rs:22-34
    left, right = orc.trim_pair("Q 10 0 10 + T 20 0 10 3 9 60 cg:Z:7=1X2=", "Q 10 5 10 - T 20 10 15 3 9 60 cg:Z:3=1X1=")
    assert left.split("\t")[-1] == "cg:Z:7=" and right.split("\t")[-1] == "cg:Z:3="
    assert left.split("\t")[2:4] == ["0", "7"] and right.split("\t")[2:4] == ["7", "10"]


@pytest.mark.parametrize("policy", [orc.RIGHTMOST, orc.EARLY_EXIT])
def test_oracle_inversion_trimming(policy):
    # trim_overlap.rs:137-170 (test_inversion_trimming): left / center (reverse strand) / right
    paf = (b"Q 20 0 10 + T 20 0 10 3 9 60 cg:Z:7=1X2=\n"
           b"Q 20 4 15 - T 20 5 16 3 9 60 cg:Z:3=1X3=1M1X2=\n"
           b"Q 20 10 20 + T 20 10 20 3 9 60 cz:Z:10= cg:Z:2=2X2=2X2=\n")
    rows = orc.run_trim_paf(paf, 1, 1, 1, False, policy).decode().splitlines()
    assert [r.split("\t")[-1] for r in rows] == ["cg:Z:7=", "cg:Z:2=1X3=1M", "cg:Z:2=2X2="]


def test_oracle_contained_records():
    # a record whose query span lies inside another's is never trimmed; --remove-contained drops it (paf.rs:243-248, 290-300)
    paf = b"Q 10 0 10 + T 20 0 10 3 9 60 cg:Z:7=1X2=\nQ 10 5 10 - T 20 10 15 3 9 60 cg:Z:3=1X1=\n"
    assert orc.run_trim_paf(paf).count(b"\n") == 2
    kept = orc.run_trim_paf(paf, remove_contained=True)
    assert kept.count(b"\n") == 1 and b"7=1X2=" in kept


def test_oracle_is_idempotent_on_random_sets():
    for seed in range(6):
        paf = gen.random_trim_paf(seed, lead_trail=False)  # (a stripped record's "_TO.." id is not read back from id:Z:)
        try:
            once = orc.run_trim_paf(paf)
        except orc.ReferencePanic:
            continue
        assert orc.run_trim_paf(once) == once  # nothing overlaps (without containment) after a run
        assert once.count(b"\n") == paf.count(b"\n")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("native") / "trim_core_check")
    subprocess.check_call(
        ["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "rustybam_b200", "csrc"), "-I",
         os.path.join(ROOT, "oracle"), "-o", exe, os.path.join(ROOT, "tests", "native", "trim_core_check.cpp"),
         os.path.join(ROOT, "oracle", "rb_oracle.cpp")])
    return exe


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_closed_form_trim_matches_literal_oracle(harness, seed):
    # (without a third argument the harness alternates the two binary_search policies from group to group)
    r = subprocess.run([harness, str(seed), "1500"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "FAIL=0" in r.stdout


@pytest.mark.parametrize("policy", [0, 1])
def test_closed_form_trim_each_search_policy(harness, policy):
    """Right-most (Rust < 1.52 / >= 1.82) and early-exit (1.52 ..= 1.81) core::slice::binary_search over the per-column query
    positions (paf.rs:564-574): the closed form replays the early-exit probe sequence over the duplicate columns of every
    position in front of a non-query run, for the record's CURRENT truncation (trim_core.cuh, trim_probe_op)."""
    r = subprocess.run([harness, str(11 + policy), "1200", str(policy)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "FAIL=0" in r.stdout


def test_f32_display_restatements_agree(tmp_path):
    """Rust's f32 `Display` (shortest round-trip digits, positional; bamstats.rs:262-265) is restated twice, independently:
    the oracle tries precisions with printf + strtof (+ an exact-expansion tie test), the product's csrc/f32_fmt.cuh (host
    formatter today, device-ready) generates Burger-Dybvig free-format digits; std::to_chars is a third opinion that may only
    differ at exact ties.  No reference test pins the digits ("parity unpinned"), so they are compared over a strided sweep of
    every f32 in [0, 100] — the identities' range — plus the neighbourhood of every power of two (where the rounding interval
    is asymmetric).  The exhaustive sweep (1.12 G values, no difference) is recorded in profiles/r01z_f32_fmt.txt."""
    exe = str(tmp_path / "f32_fmt_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "rustybam_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "f32_fmt_check.cpp"), os.path.join(ROOT, "oracle", "rb_oracle.cpp")])
    r = subprocess.run([exe, "0", "100", "4099"], capture_output=True, text=True)
    assert r.returncode == 0 and "DIFF=0" in r.stdout, r.stdout + r.stderr
    for k in range(-20, 7):  # 2^k - 64 ulp .. 2^k + 64 ulp
        lo, hi = 2.0 ** k * (1 - 2.0 ** -18), 2.0 ** k * (1 + 2.0 ** -17)
        r = subprocess.run([exe, repr(lo), repr(hi), "1"], capture_output=True, text=True)
        assert r.returncode == 0 and "DIFF=0" in r.stdout, r.stdout + r.stderr
