"""Closed-form pair resolution (the __host__ __device__ code k_lift runs on the GPU) fuzzed on the
CPU against the literal per-base oracle — both binary_search policies, ~100k pairs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("native") / "lift_core_check")
    subprocess.check_call(
        ["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "rustybam_b200", "csrc"), "-I",
         os.path.join(ROOT, "oracle"), "-o", exe, os.path.join(ROOT, "tests", "native", "lift_core_check.cpp"),
         os.path.join(ROOT, "oracle", "rb_oracle.cpp")])
    return exe


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_closed_form_matches_literal_oracle(harness, seed):
    r = subprocess.run([harness, str(seed), "4000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "FAIL=0" in r.stdout
