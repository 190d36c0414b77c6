"""End to end through the C++ stand-in driver: inFlow.dat in, DatFlow / DatContinue / DatInfo files and FIELDSTAT out,
compared byte for byte with what the oracle's state gives (two blocks with 2:1 refinement, 100 root steps)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "harness", "fsilbm_harness")
SAMPLE = os.path.join(ROOT, "tests", "golden", "inFlow_two_blocks.dat")


@pytest.fixture(scope="module")
def harness():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fsilbm3d_b200.build import build
    build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "harness")], check=True, capture_output=True)
    return HARNESS


def oracle_run(O, nsteps):
    fl = O.Flow(nu=0.04 * 8.0 / 100.0, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, volumeForceIn=(1e-6, 0.0, 0.0), ntolLBM=3, dtolLBM=1e-8)
    Fb = O.LBMBlock(24, 16, 16, dh=1.0, BndConds=(101, 104, 301, 301, 301, 301), flow=fl)
    Sb = O.LBMBlock(17, 13, 13, dh=0.5, xmin=6.0, ymin=4.0, zmin=4.0, BndConds=(0,) * 6, flow=fl)
    Fb.initialise(0.0); Sb.initialise(0.0)
    root = O.TreeNode(Fb); root.add_son(O.TreeNode(Sb), 1)
    for b in (Fb, Sb):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for n in range(1, nsteps + 1):
        O.set_blktime_all(root, float(n))
        O.tree_collision_streaming_IBM_FEM(root)
    Fb.calculate_macro_quantities(); Sb.calculate_macro_quantities()
    return Fb, Sb


def flow_bytes(b, ID, o):
    sl = (slice(o, b.xDim - o), slice(o, b.yDim - o), slice(o, b.zDim - o))
    out = np.array([b.xDim - 2 * o, b.yDim - 2 * o, b.zDim - 2 * o, ID], np.int32).tobytes()
    out += np.array([b.xmin + o * b.dh, b.ymin + o * b.dh, b.zmin + o * b.dh, b.dh]).tobytes()
    out += ((1.0 / 3.0) * (b.den[sl] - 1.0)).astype(np.float32).tobytes()
    for k in range(3):
        out += (b.uuu[(k,) + sl] * (1.0 / 0.04)).astype(np.float32).tobytes()
    return out


def test_two_block_run_files(oracle, harness, tmp_path):
    wd = str(tmp_path)
    shutil.copy(SAMPLE, os.path.join(wd, "inFlow.dat"))
    r = subprocess.run([harness, "inFlow.dat"], capture_output=True, text=True, cwd=wd, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    # Tref = 200, dt = 1: timeSimTotal 0.5 -> 100 steps; flow output at t/Tref = 0, 0.25, 0.5; continue at 0.25, 0.5
    names = sorted(os.listdir(os.path.join(wd, "DatFlow")))
    assert names == ["Flow0000000000_b001", "Flow0000000000_b002", "Flow0000025000_b001", "Flow0000025000_b002", "Flow0000050000_b001", "Flow0000050000_b002"]
    assert sorted(os.listdir(os.path.join(wd, "DatContinue"))) == ["continue0000025000", "continue0000050000"]
    Fb, Sb = oracle_run(oracle, 100)
    assert open(os.path.join(wd, "DatFlow", "Flow0000050000_b001"), "rb").read() == flow_bytes(Fb, 1, 0)
    assert open(os.path.join(wd, "DatFlow", "Flow0000050000_b002"), "rb").read() == flow_bytes(Sb, 2, 1)
    raw = open(os.path.join(wd, "DatContinue", "continue0000050000"), "rb").read()
    assert np.frombuffer(raw, np.int32, 2).tolist() == [2, 100] and np.frombuffer(raw, np.float64, 1, 8)[0] == 0.5
    off = 16 + 32 + 12
    assert np.array_equal(np.frombuffer(raw, np.float64, Fb.fIn.size, off).reshape(Fb.fIn.shape), Fb.fIn)
    off += Fb.fIn.size * 8 + 32 + 12
    assert np.array_equal(np.frombuffer(raw, np.float64, Sb.fIn.size, off).reshape(Sb.fIn.shape), Sb.fIn)
    # FIELDSTAT of both blocks, format (A,F18.12)
    st = [b.ComputeFieldStat() for b in (Fb, Sb)]
    lines = [l for l in r.stdout.splitlines() if "FIELDSTAT" in l]
    assert len(lines) == 12
    assert lines[0] == f" FIELDSTAT L2 u {st[0][0]:18.12f}" and lines[9] == f" FIELDSTAT Linfinity u {st[1][3]:18.12f}"
    # DatInfo: title line (write_information_titles, FlowCondition.f90:177-187) + 4 cadence points (0.125, 0.25, 0.375, 0.5)
    flux = open(os.path.join(wd, "DatInfo", "FluidFlux.dat")).read().splitlines()
    assert len(flux) == 5 and flux[0] == ' VARIABLES = "t"  "inlet"  "middle"  "outlet"'
    assert len(open(os.path.join(wd, "DatInfo", "FluidProbes_0002.dat")).read().splitlines()) == 5

    # restart: continue file of t/Tref = 0.25 -> run on to 0.5.  check_is_continue (FluidDomain.f90:166-224) gives every
    # node the populations of the FINEST saved block containing it, so the father's nodes under the son take the son's
    # values (coincident nodes: weights 0/1, exact); the oracle run below does the same by hand.
    wd2 = os.path.join(wd, "restart")
    os.makedirs(os.path.join(wd2, "DatContinue"))
    shutil.copy(os.path.join(wd, "DatContinue", "continue0000025000"), os.path.join(wd2, "DatContinue", "continue"))
    text = open(SAMPLE).read().replace("# isConCmpt numsubstep\n0 1", "# isConCmpt numsubstep\n1 1")
    open(os.path.join(wd2, "inFlow.dat"), "w").write(text)
    r2 = subprocess.run([harness, "inFlow.dat"], capture_output=True, text=True, cwd=wd2, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:]
    assert "Continue computing" in r2.stdout
    O = oracle
    F1, S1 = oracle_run(O, 50)
    F1.fIn[:, 6:15, 4:11, 4:11] = S1.fIn[:, ::2, ::2, ::2]
    root = O.TreeNode(F1); root.add_son(O.TreeNode(S1), 1)
    for b in (F1, S1):     # a restarted run is a new run: main.f90:62-64 again (buffers and half-way stashes start empty)
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for n in range(51, 101):
        O.set_blktime_all(root, float(n))
        O.tree_collision_streaming_IBM_FEM(root)
    F1.calculate_macro_quantities(); S1.calculate_macro_quantities()
    assert open(os.path.join(wd2, "DatFlow", "Flow0000050000_b001"), "rb").read() == flow_bytes(F1, 1, 0)
    assert open(os.path.join(wd2, "DatFlow", "Flow0000050000_b002"), "rb").read() == flow_bytes(S1, 2, 1)
    st1 = [b.ComputeFieldStat() for b in (F1, S1)]
    lines2 = [l for l in r2.stdout.splitlines() if "FIELDSTAT" in l]
    assert lines2[0] == f" FIELDSTAT L2 u {st1[0][0]:18.12f}" and lines2[6] == f" FIELDSTAT L2 u {st1[1][0]:18.12f}"
    # (a restarted two-block run is NOT the uninterrupted run: the father's footprint planes now carry the son's rescaled
    #  boundary values -- reference behaviour, reproduced above bit for bit)
    assert lines2 != lines
