"""Cross-checks of the two restatements of the reference's structural side -- harness/beam_solver.cpp +
solid_body.cpp (what the stand-in driver runs) against oracle/beam_restatement.py (numpy, written separately) -- on
forced beams and on the coupled fluid + IBM + beam loop over the CPU oracle.  CPU only.

Tolerances: both follow the reference's algorithm step for step, including the CG solver's absolute residual stop
at 1e-6 (SolidSolver.f90:1976), so iteration counts must be EQUAL; what differs is summation order (element-by-element
products there, one assembled matrix here), which the ill-conditioned Newmark operator amplifies to ~1e-10 in
positions.  Asserted: positions 1e-8 absolute, marker forces 1e-5 relative."""
import os

import numpy as np
import pytest

from tests.beam_cases import chain, open_cpp, open_numpy
from tests import fsi_cases as C

MAT = (1.0e4, 1.0e4 / 2.6, 0.01, 1.0, 0.0, 1.0e-5, 8.333e-6, 2.0e-5)
KIN = dict(freq=0.8, XYZAmpl=(0.0, 0.0, 0.1), XYZPhi=(0.0, 0.0, 30.0), AoAo=(0.0, 5.0, 0.0), AoAAmpl=(0.0, 20.0, 0.0), AoAPhi=(0.0, 90.0, 0.0),
           firstXYZ=(0.3, 0.2, 0.1), initXYZVel=(0.01, 0.0, 0.0))


@pytest.mark.parametrize("model", [2, 1], ids=["elastic", "rigid"])
def test_forced_beam_cpp_vs_numpy(model, tmp_path):
    """A 8-element plate whose leading node follows a prescribed heave + pitch + drift, loaded by seeded random marker
    forces through the nodal-load half of FluidVolumeForce_, two structural sub-steps per step, 100 steps."""
    n = 9
    grp = dict(KIN, iBodyModel=model)
    kw = dict(dampM=0.5, dampK=1e-4, dtolFEM=1e-14, ntolFEM=20)
    mesh = dict(Lspan=0.05, Rspan=0.07, dirc=(0.0, 1.0, 0.2), Nspan=3, material=MAT, group=grp)
    sb = open_cpp(str(tmp_path), chain(n), **mesh, **kw)
    b = sb.VBodies[0]
    pb = open_numpy(chain(n), **mesh, **kw)
    assert np.array_equal(b.pos, pb.pos) and np.array_equal(b.vel, pb.vel)
    X, V, Aa = pb.markers()
    assert np.allclose(b.v_Exyz, X, rtol=0, atol=1e-15) and np.allclose(b.v_Evel, V, rtol=0, atol=1e-15) and np.allclose(b.v_Ea, Aa, rtol=1e-15)
    assert np.allclose(b.mss, pb.mss, rtol=1e-14)
    dt = 0.01
    rng = np.random.default_rng(C_SEED)
    for k in range(1, 101):
        F = rng.normal(size=(b.v_nelmts, 3)) * 1e-3
        b.v_Eforce[...] = F
        b.FluidLoads()
        pb.fluid_loads(X, F)
        for s in (1, 2):
            b.structure(k * dt, s, dt, dt / 2)
            pb.structure(k * dt, s, dt, dt / 2)
        b.UpdatePosVelArea()
        X, V, Aa = pb.markers()
    assert np.allclose(b.lodFlow.reshape(-1), pb.lodFlow, rtol=0, atol=1e-14)
    assert np.abs(b.pos - pb.pos).max() <= 1e-8
    assert np.abs(b.vel - pb.vel).max() <= 1e-5 * max(1.0, np.abs(pb.vel).max())
    assert np.abs(b.v_Exyz - X).max() <= 1e-8 and np.abs(b.v_Evel - V).max() <= 1e-5
    if model == 2:
        info = b.FishInfo
        assert int(info[1]) == pb.iterNR and int(info[3]) == pb.cg_iterations and pb.cg_iterations > 1000
        se = b.strainEnergy
        st, bt = pb.strain_energy()
        assert np.allclose(se[:, 0], st, rtol=1e-6, atol=1e-16) and np.allclose(se[:, 1], bt, rtol=1e-6, atol=1e-16)
        tr = b.triads                                    # nodal and element triads stay orthonormal
        for which in range(3):
            T = tr[:, which]
            assert np.allclose(np.einsum("eki,ekj->eij", T, T), np.eye(3)[None], atol=1e-12)
        assert np.allclose(tr[:, 0], pb.Te, atol=1e-8)


C_SEED = 20261017


def test_material_from_EmR_tcR(tmp_path):
    """isKB = 0: section properties derived from EmR, tcR, denR, psR (SolidSolver.f90:1579-1595) agree between the
    restatements, and KB/KS reported back follow :1593-1594."""
    grp = dict(iBodyModel=2, EmR=4.0e4, tcR=0.02, denR=1.5, psR=0.3)
    kw = dict(isKB=0, Lref=1.0, UrefType=9, Uref=0.5, denIn=1.2)
    mesh = dict(Lspan=0.2, Rspan=0.3, dirc=(0.0, 0.0, 1.0), Nspan=4, group=grp)
    sb = open_cpp(str(tmp_path), chain(7), **mesh, **kw)
    pb = open_numpy(chain(7), **mesh, **kw)
    got = sb.VBodies[0].m_property
    got[:, 4] = 0.0   # gamma (unused) comes from the file on the C++ side
    assert np.allclose(got, pb.prop, rtol=1e-14)
    E, A, Iy = got[0, 0], got[0, 2], got[0, 6]
    assert E == pytest.approx(4.0e4 * 1.2 * 0.25) and A == pytest.approx(0.5 * 0.02) and Iy == pytest.approx(0.5 * 0.02 ** 3 / 12.0)


@pytest.mark.parametrize("case", [C.FLAG, C.HEAVE], ids=["flag", "heaving"])
def test_coupled_loop_on_oracle_cpp_vs_numpy(oracle, case, tmp_path):
    """main.f90's loop (markers -> IBM on the oracle fluid -> nodal loads -> numsubstep beam sub-steps) for 40 steps,
    once with the C++ beam, once with the numpy beam."""
    sb = C.open_structure_cpp(case, str(tmp_path / "a"))
    ob, ov, its = C.run_oracle_coupled(oracle, case, sb, 40)
    sb2 = C.open_structure_cpp(case, str(tmp_path / "b"))
    beam = C.open_structure_numpy(case, sb2)
    ob2, ov2, its2 = C.run_oracle_coupled(oracle, case, sb2, 40, structure=beam)
    b = sb.VBodies[0]
    assert its == its2
    assert int(b.FishInfo[3]) == beam.cg_iterations
    assert np.abs(b.pos - beam.pos).max() <= 1e-8
    assert np.abs(ov.v_Eforce - ov2.v_Eforce).max() <= 1e-5 * np.abs(ov.v_Eforce).max()
    assert np.abs(ob.uuu - ob2.uuu).max() <= 1e-8 * np.abs(ob.uuu).max()
    tip_motion = np.abs(b.dsp[-1, 0:3]).max()
    assert tip_motion > 1e-3, "the case must actually deform the plate"
    # force conservation of the spread: sum over cells of force*dh^3 = -sum of marker forces is checked in test_oracle_kat;
    # here: the nodal loads carry the whole marker force (Solidbody.f90:964-965)
    assert np.allclose(b.lodFlow[:, 0:3].sum(0), ov.v_Eforce.sum(0), rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize("threads", ["1", "3"])
def test_advance_equals_call_by_call(tmp_path, monkeypatch, threads):
    """SolidBodies.advance() -- nodal loads, the structural sub-steps and next step's markers of all listed bodies in ONE
    threaded call (what the step issues behind the collide-stream launch) -- leaves every body bit-identical to the
    reference's call-by-call sequence FluidVolumeForce_ (host half) -> Solver per sub-step -> UpdatePosVelArea_."""
    monkeypatch.setenv("FSILBM_SOLID_THREADS", threads)
    n = 9
    kw = dict(dampM=0.5, dampK=1e-4, dtolFEM=1e-14, ntolFEM=20)
    groups = [dict(KIN, fishNum=1, mesh="plate.dat", iBodyModel=2, iBodyType=1, isMotionGiven=(1,) * 6, firstXYZ=(0.3 + 0.5 * k, 0.2, 0.1), freq=0.8 + 0.1 * k)
              for k in range(3)]
    runs = []
    for mode in ("calls", "advance"):
        wd = os.path.join(str(tmp_path), mode)
        os.makedirs(wd)
        from fsilbm3d_b200 import solid_solver as S
        from tests.beam_cases import BOX
        S.write_plate_dat(os.path.join(wd, "plate.dat"), chain(n), 0.05, 0.07, (0.0, 1.0, 0.2), Nspan=3, material=MAT)
        with open(os.path.join(wd, "inFlow.dat"), "w") as f:
            f.write(S.inflow_text(UrefType=9, Uref=1.0, LrefType=1, Lref=1.0, isKB=2, blocks=[BOX], groups=groups, numsubstep=2, **kw))
        sb = S.SolidBodies("inFlow.dat", (301,) * 6, cwd=wd)
        assert len(sb.VBodies) == 3
        rng = np.random.default_rng(C_SEED)
        dt = 0.01
        for k in range(1, 31):
            for b in sb.VBodies:
                b.UpdatePosVelArea()                    # a no-op right after advance()
                b.v_Eforce[...] = rng.normal(size=(b.v_nelmts, 3)) * 1e-3
            if mode == "calls":
                for b in sb.VBodies:
                    b.FluidLoads()
                for s in (1, 2):
                    sb.Solver(k * dt, s, dt, dt / 2)
            else:
                sb.advance([0, 1, 2], k * dt, 2, dt)
        for b in sb.VBodies:
            b.UpdatePosVelArea()
        runs.append([(b.pos.copy(), b.vel.copy(), b.lodFlow.copy(), b.v_Exyz.copy(), b.v_Evel.copy(), b.FishInfo.copy()) for b in sb.VBodies])
        sb.close()
    for ba, bb in zip(*runs):
        for x, y in zip(ba, bb):
            assert np.array_equal(x, y)
    assert np.abs(runs[0][0][0] - runs[0][1][0]).max() > 0.1   # the three bodies are different beams
