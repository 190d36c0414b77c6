"""The C++ stand-in driver (harness/): builds against the in-tree library, parses the reference's inFlow.dat format,
and refuses to compute without a GPU (no CPU path anywhere)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "harness", "fsilbm_harness")
SAMPLE = os.path.join(ROOT, "tests", "golden", "inFlow_two_blocks.dat")


@pytest.fixture(scope="module")
def harness():
    from fsilbm3d_b200.build import build
    build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "harness")], check=True, capture_output=True)
    return HARNESS


def test_parse_inflow(harness):
    r = subprocess.run([harness, "--parse-only", SAMPLE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    p = json.loads(r.stdout.strip().splitlines()[-1])   # the reference's start-up messages come first
    assert p["npsize"] == 2 and p["nblocks"] == 2 and p["nFish"] == 0 and p["fluidProbingNum"] == 2
    assert p["uvwIn"] == [0.04, 0.0, 0.0] and p["Re"] == 100.0            # '100.0d0' Fortran exponent
    # calculate_reference_params (Solidbody.f90:219-284): Uref = |uvwIn(1)|, Tref = Lref/Uref, nu = Uref*Lref/Re
    assert p["Uref"] == 0.04 and p["Lref"] == 8.0 and p["Tref"] == 8.0 / 0.04 and p["nu"] == 0.04 * 8.0 / 100.0
    assert p["blocks"][1]["BndConds"] == [0] * 6 and p["blocks"][1]["dh"] == 0.5 and p["dtolLBM"] == 1e-8


def test_missing_section_uses_the_reference_message(harness, tmp_path):
    text = open(SAMPLE).read().replace("ProbingSolid", "ProbeSolid")
    f = tmp_path / "inFlow.dat"
    f.write_text(text)
    r = subprocess.run([harness, "--parse-only", str(f)], capture_output=True, text=True)
    assert r.returncode == 1 and "probingSolid is not found in inFlow.dat" in r.stdout


def test_keyword_match_lowercases_only_the_first_letter(harness, tmp_path):
    """to_lowercase declares `character:: string` (length 1), Util.f90:78-87: 'flowCondition' matches, 'FLOWCONDITION' does not."""
    f = tmp_path / "inFlow.dat"
    f.write_text(open(SAMPLE).read().replace("FlowCondition", "flowCondition"))
    assert subprocess.run([harness, "--parse-only", str(f)], capture_output=True).returncode == 0
    f.write_text(open(SAMPLE).read().replace("FlowCondition", "FLOWCONDITION"))
    assert subprocess.run([harness, "--parse-only", str(f)], capture_output=True).returncode == 1


def test_no_gpu_no_compute(harness, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([harness, SAMPLE], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 1 and "no CUDA device" in r.stdout
