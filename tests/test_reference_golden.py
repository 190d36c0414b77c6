"""The C oracle against the REFERENCE ITSELF: tests/golden/ref_*.npz hold the state the reference's unmodified main program
leaves on the cases of tests/reference_cases.py (its Fortran sources executed by oracle/ftn/, see make_reference_golden.py).
The oracle must reproduce the populations of every block bit for bit, and so must the marker forces, markers and nodal beam state
of the body cases (they go through the C++ structural side, harness/libfsilbm_solid.so, as well)."""
import os

import numpy as np
import pytest

from tests import reference_cases as RC
from tests.common import rel_err

NAMES = [n for n in RC.CASES if os.path.exists(RC.golden_path(n))]


def test_every_case_has_a_golden_file():
    assert sorted(NAMES) == sorted(RC.CASES)


@pytest.mark.parametrize("name", [n for n in NAMES if not RC.has_body(RC.CASES[n])])
def test_oracle_reproduces_the_reference_fluid(oracle, name):
    case, g = RC.load(name)
    assert case == RC.CASES[name] or case == __import__("json").loads(__import__("json").dumps(RC.CASES[name]))
    blocks, _, _ = RC.run_oracle(oracle, RC.CASES[name])
    for k, b in enumerate(blocks):
        ref = g[f"fIn{k}"]
        assert ref.shape == b.fIn.shape
        assert np.array_equal(b.fIn, ref), f"block {k}: {int((b.fIn != ref).sum())} of {ref.size} populations differ, max rel {rel_err(b.fIn, ref):.3e}"
        assert np.array_equal(b.den, g[f"den{k}"]) and np.array_equal(b.uuu, g[f"uuu{k}"])
    # FIELDSTAT (FluidDomain.f90:1739-1789, format (A,F18.12)) as the reference printed it
    st = blocks[0].ComputeFieldStat()
    lines = [str(x) for x in g["fieldstat"]]
    assert lines[0] == f" FIELDSTAT L2 u {st[0]:18.12f}"


@pytest.mark.parametrize("name", [n for n in NAMES if RC.has_body(RC.CASES[n])])
def test_oracle_and_cpp_structure_reproduce_the_reference_body_case(oracle, name, tmp_path):
    from fsilbm3d_b200 import solid_solver as S
    case, g = RC.CASES[name], RC.load(name)[1]
    wd = str(tmp_path)
    RC.write_inputs(case, wd)
    sb = S.SolidBodies("inFlow.dat", case["bc"], cwd=wd)
    blocks, ov, its = RC.run_oracle(oracle, case, sb)
    ovs = ov if isinstance(ov, list) else [ov]
    flexible = False      # flexible plates are bit-exact as well since the C++ structural side keeps the association of Uref**2
    # the reference entered PenaltyForce_ sum(iterLBM) x bodies times: same iteration counts, step by step in total
    assert sum(its) * len(ovs) == int(g["penalty_calls"])
    if case["dtolLBM"] > 1e-20:
        assert min(its) < case["ntolLBM"], "the tolerance never ended the iteration early: the case does not test the loop control"
    e_f = rel_err(blocks[0].fIn, g["fIn0"])
    exact = np.array_equal(blocks[0].fIn, g["fIn0"])
    worst = dict(F=0.0, x=0.0, p=0.0, v=0.0)
    for k, (body, o) in enumerate(zip(sb.VBodies, ovs)):
        worst["F"] = max(worst["F"], rel_err(np.array(o.v_Eforce), g[f"body{k}_v_Eforce"]))
        worst["x"] = max(worst["x"], rel_err(np.array(o.v_Exyz), g[f"body{k}_v_Exyz"]))
        worst["p"] = max(worst["p"], rel_err(body.pos, g[f"body{k}_pos"]))
        worst["v"] = max(worst["v"], rel_err(body.vel, g[f"body{k}_vel"]))
        if not flexible:
            assert np.array_equal(np.array(o.v_Eforce), g[f"body{k}_v_Eforce"]) and np.array_equal(np.array(o.v_Exyz), g[f"body{k}_v_Exyz"])
    print(f"{name}: rel err fIn {e_f:.2e} marker force {worst['F']:.2e} markers {worst['x']:.2e} beam pos {worst['p']:.2e} vel {worst['v']:.2e}; "
          f"fIn bit-exact: {exact}; iterations {its}")
    if not flexible:
        for k, b in enumerate(blocks):      # (a plate carried by a refined son: both blocks)
            assert np.array_equal(b.fIn, g[f"fIn{k}"]) and np.array_equal(b.den, g[f"den{k}"]) and np.array_equal(b.uuu, g[f"uuu{k}"]), k
    assert e_f <= 1e-12 and worst["x"] <= 1e-12 and worst["p"] <= 1e-12
    assert worst["F"] == 0.0 and worst["v"] == 0.0 and worst["p"] == 0.0
    sb.close()


@pytest.mark.parametrize("name", [n for n in NAMES if RC.CASES[n].get("outputs")])
def test_output_files_of_the_reference(oracle, name):
    """The files the reference wrote (main.f90:124-143), byte for byte, against what the oracle's end state and the product's
    host-side formatters (fsilbm3d_b200.flow_io) give: flow fields as real(4) (FluidDomain.f90:1640-1723), flux and probe lines in
    E20.10 (FluidDomain.f90:2051-2054, FlowCondition.f90:218-219)."""
    from fsilbm3d_b200 import flow_io
    case, g = RC.CASES[name], RC.load(name)[1]
    files = {k[5:]: bytes(g[k]) for k in g.files if k.startswith("file:")}
    blocks, _, _ = RC.run_oracle(oracle, case)
    Uref, Tref = case["Uref"], case["Lref"] / case["Uref"]
    tname = flow_io._name10(case["steps"] / Tref)
    for k, b in enumerate(blocks):
        o = 0 if k == 0 else 1                                    # offsetOutput of the son block in these cases
        sl = (slice(o, b.xDim - o), slice(o, b.yDim - o), slice(o, b.zDim - o))
        want = np.array([b.xDim - 2 * o, b.yDim - 2 * o, b.zDim - 2 * o, k + 1], np.int32).tobytes()
        want += np.array([b.xmin + o * b.dh, b.ymin + o * b.dh, b.zmin + o * b.dh, b.dh]).tobytes()
        want += ((1.0 / 3.0) * (b.den[sl] - 1.0)).astype(np.float32).tobytes()
        for c in range(3):
            want += (b.uuu[(c,) + sl] * (1.0 / Uref)).astype(np.float32).tobytes()
        assert files[f"DatFlow/Flow{tname}_b{k + 1:03d}"] == want, f"flow file of block {k + 1}"
    # flux through the inlet, middle and outlet planes of the root block, trapezoid weights on y and z
    b = blocks[0]
    wy = np.ones(b.yDim); wy[0] = wy[-1] = 0.5
    wz = np.ones(b.zDim); wz[0] = wz[-1] = 0.5
    per = lambda lo, hi: 1.0 if (lo == 301 and hi == 301) else 0.0
    Yref = (b.yDim - 1) * b.dh + per(case["bc"][2], case["bc"][3]) * b.dh
    Zref = (b.zDim - 1) * b.dh + per(case["bc"][4], case["bc"][5]) * b.dh
    vals = []
    for ix in (0, (b.xDim + 1) // 2 - 1, b.xDim - 1):
        acc = 0.0
        for kz in range(b.zDim):                                   # the reference's loop order: k outer, j inner
            for jy in range(b.yDim):
                acc = acc + b.uuu[0, ix, jy, kz] * b.den[ix, jy, kz] * b.dh * b.dh * wy[jy] * wz[kz]
        vals.append(acc / (1.0 * Uref * Zref * Yref))
    line = "".join(flow_io._e20_10(v) for v in (case["steps"] / Tref, *vals))
    got = files["DatInfo/FluidFlux.dat"].decode().splitlines()
    assert got[0] == ' VARIABLES = "t"  "inlet"  "middle"  "outlet"' and got[1] == line
    assert len([k for k in files if "FluidProbes" in k]) == len(case["probes"])
    if case.get("outputtype", 1) >= 2:
        # running means of calculate_turbulent_statistic_ (FluidDomain.f90:1147-1172; invStep = 1/real(n) in real(4)), start step 0
        # (main.f90:76-80 with timeWriteBegin = 0), from the oracle's velocity after every step
        ave = np.zeros((9,) + blocks[0].den.shape)
        for step in range(1, case["steps"] + 1):
            bk, _, _ = RC.run_oracle(oracle, dict(case, steps=step))
            u = bk[0].uuu
            inv = float(np.float32(1) / np.float32(step - 0 + 1))
            for c in range(3):
                ave[c] = ave[c] * (1.0 - inv) + inv * u[c]
            for c in range(3):
                ave[3 + c] = ave[3 + c] * (1.0 - inv) + inv * (u[c] - ave[c]) * (u[c] - ave[c])
            for n, (a, c) in enumerate(((0, 1), (0, 2), (1, 2))):
                ave[6 + n] = ave[6 + n] * (1.0 - inv) + inv * (u[a] - ave[a]) * (u[c] - ave[c])
        b = blocks[0]
        want = np.array([b.xDim, b.yDim, b.zDim, 1], np.int32).tobytes() + np.array([b.xmin, b.ymin, b.zmin, b.dh]).tobytes()
        want += ((1.0 / 3.0) * (b.den - 1.0)).astype(np.float32).tobytes()
        for c in range(3):
            want += (ave[c] * (1.0 / Uref)).astype(np.float32).tobytes()
        for c in range(3, 9):
            want += (ave[c] * (1.0 / Uref / Uref)).astype(np.float32).tobytes()
        assert files["DatFlow/MeanFlow_b001"] == want
