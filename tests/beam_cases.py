"""Shared builders for the structural-solver tests: the same beam / plate case as (i) the two input files the C++
stand-in driver reads and (ii) constructor arguments of the independent numpy restatement."""
import os

import numpy as np

from fsilbm3d_b200 import solid_solver as S

BOX = dict(dims=(8, 8, 8), BndConds=(301,) * 6)   # a placeholder fluid block for runs that never touch the fluid


def chain(n, length=1.0, origin=(0.0, 0.0, 0.0), axis=(1.0, 0.0, 0.0)):
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    return np.asarray(origin, float)[None, :] + np.linspace(0.0, length, n)[:, None] * a[None, :]


def open_cpp(wd, xyz, *, Lspan=0.05, Rspan=0.05, dirc=(0.0, 1.0, 0.0), Nspan=2, material=None, constraint=None, group=None, rootBC=(301,) * 6,
             blocks=(BOX,), **inflow):
    """Writes wd/plate.dat + wd/inFlow.dat and opens them with the C++ structural side."""
    os.makedirs(wd, exist_ok=True)
    S.write_plate_dat(os.path.join(wd, "plate.dat"), xyz, Lspan, Rspan, dirc, constraint=constraint, Nspan=Nspan,
                      material=material if material is not None else (1.0,) * 8)
    g = dict(fishNum=1, mesh="plate.dat", iBodyModel=2, iBodyType=1, isMotionGiven=(1,) * 6)
    g.update(group or {})
    kw = dict(UrefType=9, Uref=1.0, LrefType=1, Lref=1.0, isKB=2, blocks=list(blocks), groups=[g])
    kw.update(inflow)
    with open(os.path.join(wd, "inFlow.dat"), "w") as f:
        f.write(S.inflow_text(**kw))
    return S.SolidBodies("inFlow.dat", rootBC, cwd=wd)


def open_numpy(xyz, *, Lspan=0.05, Rspan=0.05, dirc=(0.0, 1.0, 0.0), Nspan=2, material=None, constraint=None, group=None, **inflow):
    """The same case on oracle/beam_restatement.Beam."""
    from oracle.beam_restatement import Beam
    g = dict(iBodyModel=2, isMotionGiven=(1,) * 6)
    g.update(group or {})
    n = len(xyz)
    con = constraint if constraint is not None else [[1] * 6] + [[0] * 6] * (n - 1)
    z3 = (0.0, 0.0, 0.0)
    return Beam(xyz, Lspan, Rspan, dirc, con, Nspan, iBodyModel=g["iBodyModel"], isMotionGiven=g["isMotionGiven"], prop=material,
                isKB=inflow.get("isKB", 2), EmR=g.get("EmR", 0.0), tcR=g.get("tcR", 0.0), psR=g.get("psR", 0.3), denR=g.get("denR", 1.0),
                Lref=inflow.get("Lref", 1.0), Uref=inflow.get("Uref", 1.0), denIn=inflow.get("denIn", 1.0), Freq=g.get("freq", 0.0),
                XYZo=g.get("firstXYZ", z3), initXYZVel=g.get("initXYZVel", z3), XYZAmpl=g.get("XYZAmpl", z3), XYZPhi_deg=g.get("XYZPhi", z3),
                AoAo_deg=g.get("AoAo", z3), AoAAmpl_deg=g.get("AoAAmpl", z3), AoAPhi_deg=g.get("AoAPhi", z3),
                dampK=inflow.get("dampK", 0.0), dampM=inflow.get("dampM", 0.0), GeoGamma=inflow.get("GeoGamma", 1.0),
                NewmarkGamma=inflow.get("NewmarkGamma", 0.5), NewmarkBeta=inflow.get("NewmarkBeta", 0.25),
                dtolFEM=inflow.get("dtolFEM", 1e-10), ntolFEM=inflow.get("ntolFEM", 20), IBPenaltyAlpha=inflow.get("IBPenaltyAlpha", 1.0))
