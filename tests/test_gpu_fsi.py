"""Flexible-plate fluid-structure coupling on the CUDA path (BASELINE configs[0] and [3] in
miniature).  The structural side is the C++ restatement of SolidSolver.f90 on both sides of the comparison (it is host
code in the reference as well); what is compared is the library's fluid + IBM path against the CPU oracle inside the
closed FSI loop, where every marker force feeds back into the plate position of the next step.
Tolerances are north_star's: <= 1e-12 relative on density and velocity, <= 1e-10 relative on body forces and plate
positions; IBM iteration counts must be equal."""
import os
import subprocess

import numpy as np
import pytest

from tests import fsi_cases as C
from tests.common import rel_err

pytestmark = pytest.mark.gpu
TOL_FLUID, TOL_FORCE = 1e-12, 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "harness", "fsilbm_harness")


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fsilbm3d_b200 as F
    F.lib()
    return F


def run_gpu_coupled(F, case, sb, steps):
    fk = C.flow_kwargs(case, sb)
    X, Y, Z = case["dims"]
    gb = F.LBMBlock(X, Y, Z, dh=1.0, BndConds=case["BndConds"], flow=F.FlowCondType(**fk))
    gb.initialise(0.0)
    gb.update_volume_force(); gb.set_boundary_conditions()
    its = []
    for k in range(1, steps + 1):
        its.append(F.tree_collision_streaming_IBM_FEM(gb, sb.plates, time=float(k)))
    return gb, its


@pytest.mark.parametrize("case,steps", [(C.FLAG, 60), (C.HEAVE, 60)], ids=["flag_uniform_inflow", "heaving_pitching"])
def test_flexible_plate_parity(oracle, F, case, steps, tmp_path):
    sb_o = C.open_structure_cpp(case, str(tmp_path / "o"))
    ob, ov, its_o = C.run_oracle_coupled(oracle, case, sb_o, steps)
    sb_g = C.open_structure_cpp(case, str(tmp_path / "g"))
    gb, its_g = run_gpu_coupled(F, case, sb_g, steps)
    assert its_o == its_g
    bo, bg = sb_o.VBodies[0], sb_g.VBodies[0]
    den, uuu = gb.download_macro()
    assert rel_err(den, ob.den) <= TOL_FLUID and rel_err(uuu, ob.uuu) <= TOL_FLUID
    assert rel_err(bg.v_Eforce, np.array(ov.v_Eforce)) <= TOL_FORCE
    assert rel_err(bg.pos[:, 0:3], bo.pos[:, 0:3]) <= TOL_FORCE
    assert rel_err(bg.lodFlow, bo.lodFlow) <= TOL_FORCE
    # the ordered IBM mode keeps the reference's summation order, so on one GPU the closed loop is in fact bit-identical
    assert np.array_equal(uuu, ob.uuu) and np.array_equal(bg.v_Eforce, np.array(ov.v_Eforce)) and np.array_equal(bg.pos, bo.pos)
    assert np.abs(bo.dsp[-1, 0:3]).max() > 1e-3          # the plate really moved
    assert bg.FishInfo[3] == bo.FishInfo[3]               # same number of CG iterations in the beam solver
    gb.close()


def test_harness_flexible_plate_end_to_end(oracle, F, tmp_path):
    """inFlow.dat + plate.dat -> the C++ stand-in driver (beam solver + libfsilbm_b200.so) -> DatBody / DatBodySpan /
    DatInfo files; the end state of the plate is compared with the oracle-fluid run of the same loop."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "harness")], check=True, capture_output=True)
    case = C.FLAG
    wd = str(tmp_path / "run")
    steps = 40
    sb = C.open_structure_cpp(case, wd, timeSimTotal=steps / 160.0, timeBodyDelta=0.125, timeInfoDelta=0.0625, timeFlowDelta=0.25, solidProbes=(3, 9))
    r = subprocess.run([HARNESS, "inFlow.dat"], capture_output=True, text=True, cwd=wd, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert f" IBM iterations (total): {3 * steps}" in r.stdout
    ob, ov, its = C.run_oracle_coupled(oracle, case, sb, steps)
    b = sb.VBodies[0]
    rows = [l.split() for l in open(os.path.join(wd, "DatInfo", "BodiesFinal.txt")) if l.startswith("NODE")]
    got = np.array([[float(v) for v in row[2:]] for row in rows])
    assert got.shape == (b.nND, 18)
    assert rel_err(got[:, 0:3], b.pos[:, 0:3]) <= TOL_FORCE
    assert rel_err(got[:, 12:18], b.lodFlow) <= TOL_FORCE
    # files of main.f90:85-87,124-128,143 with the reference's names and zone headers
    assert sorted(os.listdir(os.path.join(wd, "DatBody"))) == ["Bodies_0000000000.dat", "Bodies_0000012500.dat", "Bodies_0000025000.dat"]
    assert sorted(os.listdir(os.path.join(wd, "DatBodySpan")))[0] == "BodiesVirtual_0000000000.dat"
    body = open(os.path.join(wd, "DatBody", "Bodies_0000025000.dat")).read().splitlines()
    assert body[2] == ' ZONE T = "fish0001"' and body[4] == f" Nodes={b.nND:8d}, Elements={b.nEL:8d}, ZONETYPE=FELINESEG"
    assert len(body) == 7 + b.nND + b.nEL
    x_tip = float(body[7 + b.nND - 1].split()[0]) * sb.Lref
    assert x_tip == pytest.approx(b.pos[-1, 0], rel=1e-9)
    forces = open(os.path.join(wd, "DatInfo", "Group001_forces.dat")).read().splitlines()
    assert forces[0].strip().startswith("VARIABLES") and len(forces) == 1 + 2 * 4   # title + (zone, row) at 4 cadence points
    Fx = float(forces[-1].split()[3]) * sb.Fref
    assert Fx == pytest.approx(b.lodFlow[:, 0].sum(), rel=1e-8)
    assert os.path.exists(os.path.join(wd, "DatInfo", "Group001_solidProbes_0002.dat"))
    assert os.path.exists(os.path.join(wd, "DatInfo", "Group001_energy.dat"))
    assert "nFish = 0001" in open(os.path.join(wd, "Check.dat")).read()
