"""Flexible-plate cases shared by the coupled (fluid + IBM + beam) tests: one definition that yields (i) inFlow.dat +
plate.dat for the C++ stand-in driver / its structural library, (ii) the arguments of the numpy beam restatement,
(iii) the fluid block on either backend.  BASELINE configs[0] in miniature: a flexible plate, leading edge held, in
uniform inflow (lattice units, dh = 1)."""
import os

import numpy as np

from fsilbm3d_b200 import solid_solver as S
from tests.beam_cases import chain, open_cpp, open_numpy

FLAG = dict(
    dims=(40, 24, 24), BndConds=(101, 104, 301, 301, 301, 301), uvwIn=(0.05, 0.0, 0.0), Re=80.0,
    nEL=8, chord=8.0, span=6.0, Nspan=6, origin=(12.3, 11.6, 9.2),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=1.0, psR=0.3, KB=0.02, KS=500.0, AoAo=(0.0, 0.0, 12.0), firstXYZ=(12.3, 11.6, 9.2)),
    inflow=dict(isKB=1, LrefType=0, UrefType=0, TrefType=0, numsubstep=2, ntolLBM=3, dtolLBM=1e-30, dtolFEM=1e-12, ntolFEM=20, dampM=0.0, dampK=0.0),
)

HEAVE = dict(   # a flapping flexible plate: prescribed heave + pitch at the leading edge (configs[3] in miniature)
    dims=(40, 24, 24), BndConds=(101, 104, 301, 301, 301, 301), uvwIn=(0.04, 0.0, 0.0), Re=60.0,
    nEL=8, chord=8.0, span=6.0, Nspan=6, origin=(12.3, 11.6, 9.2),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=2.0, psR=0.3, KB=0.05, KS=800.0, freq=0.002, XYZAmpl=(0.0, 1.5, 0.0),
               AoAAmpl=(0.0, 0.0, 10.0), AoAPhi=(0.0, 0.0, 90.0), firstXYZ=(12.3, 11.6, 9.2)),
    inflow=dict(isKB=1, LrefType=0, UrefType=0, TrefType=0, numsubstep=4, ntolLBM=3, dtolLBM=1e-30, dtolFEM=1e-12, ntolFEM=20, dampM=0.0, dampK=0.0),
)


def plate_mesh(case):
    """Reference-frame mesh: the beam runs along x from the origin 0 (XYZo = firstXYZ carries the position), markers
    extend along z from -Lspan to +Rspan."""
    xyz = chain(case["nEL"] + 1, case["chord"])
    return dict(xyz=xyz, Lspan=0.5 * case["span"], Rspan=0.5 * case["span"], dirc=(0.0, 0.0, 1.0), Nspan=case["Nspan"])


def inflow_kwargs(case, **extra):
    kw = dict(case["inflow"])
    kw.update(Re=case["Re"], uvwIn=case["uvwIn"], blocks=[dict(dims=case["dims"], BndConds=case["BndConds"])])
    kw.update(extra)
    return kw


def open_structure_cpp(case, wd, **extra):
    m = plate_mesh(case)
    return open_cpp(wd, m["xyz"], Lspan=m["Lspan"], Rspan=m["Rspan"], dirc=m["dirc"], Nspan=m["Nspan"], group=case["group"],
                    rootBC=case["BndConds"], **inflow_kwargs(case, **extra))


def open_structure_numpy(case, sb):
    """The numpy beam with the reference quantities the C++ side derived (Lref from the chord, Uref from uvwIn)."""
    m = plate_mesh(case)
    kw = dict(case["inflow"])
    kw.update(Lref=sb.Lref, Uref=sb.Uref, denIn=sb.denIn)
    g = dict(case["group"])
    if kw.get("isKB") == 1:   # Beam_calculate_angle_material, SolidSolver.f90:1597-1614, restated here for the numpy side's `prop`
        prop = []
        spanlen = m["Lspan"] + m["Rspan"]
        th = np.sqrt(g["KB"] / g["KS"] * 12.0) * sb.Lref
        Aa = spanlen * th
        Em = g["KS"] * sb.denIn * sb.Uref ** 2 * sb.Lref * spanlen / Aa
        ratio = th / spanlen
        prop = (Em, Em / (2.0 * (1.0 + g["psR"])), Aa, g["denR"] * spanlen * sb.Lref * sb.denIn / Aa, 0.0,
                spanlen * th ** 3 / 3.0 * (1.0 - 0.63 * ratio + 0.052 * ratio ** 5), spanlen * th ** 3 / 12.0, th * spanlen ** 3 / 12.0)
        kw["isKB"] = 2
        return open_numpy(m["xyz"], Lspan=m["Lspan"], Rspan=m["Rspan"], dirc=m["dirc"], Nspan=m["Nspan"], material=prop, group=g, **kw)
    return open_numpy(m["xyz"], Lspan=m["Lspan"], Rspan=m["Rspan"], dirc=m["dirc"], Nspan=m["Nspan"], group=g, **kw)


def flow_kwargs(case, sb):
    return dict(nu=sb.nu, uvwIn=case["uvwIn"], Uref=sb.Uref, ntolLBM=sb.ntolLBM, dtolLBM=sb.dtolLBM, numsubstep=sb.numsubstep)


def run_oracle_coupled(O, case, sb, steps, structure="cpp", record=None):
    """main.f90's loop on the oracle fluid with the beam on `structure` ('cpp' = harness library, or a numpy Beam)."""
    fk = flow_kwargs(case, sb)
    nsub = fk.pop("numsubstep")
    X, Y, Z = case["dims"]
    ob = O.LBMBlock(X, Y, Z, dh=1.0, BndConds=case["BndConds"], flow=O.Flow(**fk))
    ob.initialise(0.0)
    ob.update_volume_force(); ob.set_boundary_conditions(); ob.calculate_macro_quantities()
    body = sb.VBodies[0]
    beam = None if structure == "cpp" else structure
    n = body.v_nelmts
    ov = O.VirtualBody(n, v_move=body.v_move, iBodyModel=body.iBodyModel)
    its = []
    for k in range(1, steps + 1):
        t = float(k)
        ob.set_blktime(t)
        if beam is None:
            body.UpdatePosVelArea()
            ov.v_Exyz[...] = body.v_Exyz; ov.v_Evel[...] = body.v_Evel; ov.v_Ea[...] = body.v_Ea
        else:
            ov.v_Exyz[...], ov.v_Evel[...], ov.v_Ea[...] = beam.markers()
        its.append(ob.step([ov]))
        if beam is None:
            body.v_Eforce[...] = ov.v_Eforce
            body.FluidLoads()
            for isub in range(1, nsub + 1):
                body.structure(t, isub, 1.0, 1.0 / nsub)
        else:
            beam.fluid_loads(np.array(ov.v_Exyz), np.array(ov.v_Eforce))
            for isub in range(1, nsub + 1):
                beam.structure(t, isub, 1.0, 1.0 / nsub)
        if record is not None:
            record.append((np.array(ov.v_Eforce), (body.pos if beam is None else beam.pos).copy()))
    ob.calculate_macro_quantities()
    return ob, ov, its
