"""ctypes front-end of harness/libfsilbm_solid.so -- the CPU-only structural side of the stand-in driver
(harness/beam_solver.cpp = SolidSolver.f90, harness/solid_body.cpp = the host half of Solidbody.f90) -- plus writers
for the two input files it reads (inFlow.dat, the structural line mesh `plate.dat`).

Not part of the drop-in boundary: in the reference these procedures belong to the Fortran driver and stay there.
The tests and bench.py use this module to feed the CUDA path (and the oracle) the markers of a flexible plate and to
advance the beam with the forces that come back.  Arrays are numpy views in C order: markers (n,3), nodes (nND,6).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time as _time
from typing import Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_DIR = os.path.join(ROOT, "harness")
LIB = os.path.join(HARNESS_DIR, "libfsilbm_solid.so")
_lib = None


class SolidError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB):
        r = subprocess.run(["make", "-C", HARNESS_DIR, "libfsilbm_solid.so"] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise SolidError("building libfsilbm_solid.so failed:\n" + r.stdout + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", HARNESS_DIR, "libfsilbm_solid.so"], capture_output=True, text=True)   # rebuild if sources are newer
        build()
        L = ctypes.CDLL(LIB)
        L.fsolid_last_error.restype = ctypes.c_char_p
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.fsolid_open.argtypes = [ctypes.c_char_p, ip, ip]
        L.fsolid_close.argtypes = [ctypes.c_int]
        L.fsolid_flow.argtypes = [ctypes.c_int, dp]
        L.fsolid_nfish.argtypes = [ctypes.c_int]
        L.fsolid_body_info.argtypes = [ctypes.c_int, ctypes.c_int, ip]
        L.fsolid_update_pos_vel_area.argtypes = [ctypes.c_int, ctypes.c_int]
        L.fsolid_markers.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp]
        L.fsolid_set_eforce.argtypes = [ctypes.c_int, ctypes.c_int, dp]
        L.fsolid_fluid_loads.argtypes = [ctypes.c_int, ctypes.c_int]
        L.fsolid_set_lodflow.argtypes = [ctypes.c_int, ctypes.c_int, dp]
        L.fsolid_structure.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.fsolid_solver.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.fsolid_marker_ptrs.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(dp)] * 4
        L.fsolid_advance.argtypes = [ctypes.c_int, ctypes.c_int, ip, ctypes.c_double, ctypes.c_int, ctypes.c_double]
        L.fsolid_get.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        L.fsolid_write.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ---------------------------------------------------------------------------------------------------------------------
# input files
# ---------------------------------------------------------------------------------------------------------------------

def write_plate_dat(path: str, xyz: np.ndarray, Lspan, Rspan, dirc=(0.0, 0.0, 1.0), constraint: Sequence[Sequence[int]] | None = None,
                    Nspan: int | Sequence[int] = 1, material=(1.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0), itype: int = 2) -> None:
    """The structural line mesh in the layout Beam_ReadBuild reads (SolidSolver.f90:1270-1350; sample at
    Solidbody.f90:1150-1168): a chain of nND nodes, element n joining nodes n and n+1."""
    xyz = np.asarray(xyz, float)
    nND = xyz.shape[0]
    nEL = nND - 1
    Lspan = np.broadcast_to(np.asarray(Lspan, float), (nND,))
    Rspan = np.broadcast_to(np.asarray(Rspan, float), (nND,))
    dirc = np.broadcast_to(np.asarray(dirc, float), (nND, 3))
    Nspan = np.broadcast_to(np.asarray(Nspan, int), (nEL,))
    if constraint is None:
        constraint = [[1] * 6] + [[0] * 6] * (nND - 1)   # leading node clamped (takes isMotionGiven, Beam_adjustBC)
    with open(path, "w") as f:
        f.write("Frame3D : point number, element number, material number\n")
        f.write(f"{nND:5d}{nEL:5d}{1:5d}\n")
        f.write("POINT\n")
        f.write(f"{nND:5d} X Y Z Lspan Rspan dirX dirY dirZ\n")
        for i in range(nND):
            f.write(f"{i + 1:5d} " + " ".join(f"{v:.17g}" for v in (*xyz[i], Lspan[i], Rspan[i], *dirc[i])) + "\n")
        f.write("ELEMENT\n")
        f.write(f"{nEL:5d} I J K TYPE MAT Nspan\n")
        for n in range(nEL):
            f.write(f"{n + 1:5d}{n + 1:6d}{n + 2:6d}{n + 2:6d}{itype:6d}{1:6d}{int(Nspan[n]):6d}\n")
        f.write("CONSTRAINT\n")
        f.write(f"{nND:5d} XTRA YTRA ZTRA XROT YROT ZROT\n")
        for i in range(nND):
            f.write(f"{i + 1:5d}" + "".join(f"{int(c):6d}" for c in constraint[i]) + "\n")
        f.write("MATERIAL\n")
        f.write(f"{1:5d} E G A RHO GAMMA JT IY IZ\n")
        f.write(f"{1:5d} " + " ".join(f"{v:.17g}" for v in material) + "\n")
        f.write("END\n")


def inflow_text(*, npsize=1, isConCmpt=0, numsubstep=1, timeSimTotal=1.0, timeContiDelta=1e9, timeWriteBegin=0.0, timeWriteEnd=1e9,
                timeFlowDelta=1e9, timeBodyDelta=1e9, timeInfoDelta=1e9, Re=100.0, denIn=1.0, uvwIn=(0.0, 0.0, 0.0),
                shearRateIn=(0.0, 0.0, 0.0), velocityKind=0, volumeForceIn=(0.0, 0.0, 0.0), volumeForceAmp=0.0, volumeForceFreq=0.0,
                volumeForcePhi=0.0, LrefType=1, Lref=1.0, TrefType=0, Tref=1.0, UrefType=0, Uref=1.0, ntolLBM=3, dtolLBM=1e-8,
                interpolateScheme=1, blocks=(), IBPenaltyAlpha=1.0, GeoGamma=1.0, NewmarkGamma=0.5, NewmarkBeta=0.25, dampK=0.0,
                dampM=0.0, dtolFEM=1e-10, ntolFEM=20, isKB=0, groups=(), fluidProbes=(), inWhichBlock=1, solidProbes=()) -> str:
    """inFlow.dat in the reference's keyword-sectioned format.  `blocks`: dicts with ID,
    iCollidModel, offsetOutput, outputtype, dims, dh, xyzmin, BndConds, params.  `groups`: dicts with the per-group
    lines of the SolidBody section (fishNum, numXYZ, mesh, iBodyModel, iBodyType, isMotionGiven, denR, psR, EmR/tcR
    or KB/KS, freq, St, firstXYZ, deltaXYZ, initXYZVel, XYZAmpl, XYZPhi, AoAo, AoAAmpl, AoAPhi)."""
    def v(a):
        return " ".join(f"{x:.17g}" if isinstance(x, float) else str(x) for x in a)
    L = ["Parallel", str(npsize), "FlowCondition", f"{isConCmpt} {numsubstep}", v([float(timeSimTotal), float(timeContiDelta)]),
         v([float(timeWriteBegin), float(timeWriteEnd)]), v([float(timeFlowDelta), float(timeBodyDelta), float(timeInfoDelta)]),
         v([float(Re), float(denIn)]), v(map(float, uvwIn)), v(list(map(float, shearRateIn)) + [velocityKind]), v(map(float, volumeForceIn)),
         v([float(volumeForceAmp), float(volumeForceFreq), float(volumeForcePhi)]), v([LrefType, float(Lref)]), v([TrefType, float(Tref)]),
         v([UrefType, float(Uref)]), v([ntolLBM, float(dtolLBM)]), str(interpolateScheme), "FluidBlocks", str(len(blocks))]
    for i, b in enumerate(blocks):
        L += [v([b.get("ID", i + 1), b.get("iCollidModel", 1), b.get("offsetOutput", 0), b.get("outputtype", 1)]), v(b["dims"]),
              v([float(b.get("dh", 1.0))] + list(map(float, b.get("xyzmin", (0.0, 0.0, 0.0))))), v(b["BndConds"]),
              v(map(float, b.get("params", (0.0,) * 10)))]
        if i < len(blocks) - 1:
            L.append("====================")
    nFish = sum(g["fishNum"] for g in groups)
    L += ["SolidBody", v([float(IBPenaltyAlpha), float(GeoGamma)]), v([float(NewmarkGamma), float(NewmarkBeta)]), v([float(dampK), float(dampM)]),
          v([float(dtolFEM), ntolFEM]), f"{nFish} {len(groups)} {isKB}"]
    for i, g in enumerate(groups):
        z3 = (0.0, 0.0, 0.0)
        L += [v([g["fishNum"]] + list(g.get("numXYZ", (g["fishNum"], 1, 1)))), g["mesh"], v([g.get("iBodyModel", 1), g.get("iBodyType", 1)]),
              v(g.get("isMotionGiven", (1,) * 6)[:3]), v(g.get("isMotionGiven", (1,) * 6)[3:]), v([float(g.get("denR", 1.0)), float(g.get("psR", 0.3))]),
              v([float(g.get("KB", 0.0)), float(g.get("KS", 0.0))]) if isKB == 1 else v([float(g.get("EmR", 0.0)), float(g.get("tcR", 0.0))]),
              v([float(g.get("freq", 0.0)), float(g.get("St", 0.0))])]
        for key in ("firstXYZ", "deltaXYZ", "initXYZVel", "XYZAmpl", "XYZPhi", "AoAo", "AoAAmpl", "AoAPhi"):
            L.append(v(map(float, g.get(key, z3))))
        if i < len(groups) - 1:
            L.append("====================")
    L += ["ProbingFluid", f"{len(fluidProbes)} {inWhichBlock}"] + [v(map(float, p)) for p in fluidProbes]
    L += ["ProbingSolid", str(len(solidProbes))] + [str(int(p)) for p in solidProbes]
    return "\n".join(L) + "\n"


# ---------------------------------------------------------------------------------------------------------------------
# module SolidBody over the C API
# ---------------------------------------------------------------------------------------------------------------------

class Body:
    """One VBodies(iFish) (Solidbody.f90:25-68) with its rbm (SolidSolver.f90:1160)."""

    def __init__(self, owner: "SolidBodies", index: int):
        self._o, self.index = owner, index
        info = (ctypes.c_int * 8)()
        owner._ck(lib().fsolid_body_info(owner.h, index, info))
        self.nND, self.nEL, self.v_nelmts, self.v_move, self.iBodyModel, self.gEQ, _, self.v_type = list(info)
        self.count_Interp = 0
        n = self.v_nelmts
        # the marker arrays ARE the structural library's own v_Exyz / v_Evel / v_Ea / v_Eforce (no copies either way)
        dp = ctypes.POINTER(ctypes.c_double)
        p = [dp(), dp(), dp(), dp()]
        owner._ck(lib().fsolid_marker_ptrs(owner.h, index, *[ctypes.byref(q) for q in p]))
        self.v_Exyz = np.ctypeslib.as_array(p[0], shape=(n, 3))
        self.v_Evel = np.ctypeslib.as_array(p[1], shape=(n, 3))
        self.v_Ea = np.ctypeslib.as_array(p[2], shape=(n,))
        self.v_Eforce = np.ctypeslib.as_array(p[3], shape=(n, 3))
        self.markers_fresh = False   # True after SolidBodies.advance(): UpdatePosVelArea_ of the coming step is done

    def UpdatePosVelArea(self):
        self._o.flush()
        if self.markers_fresh:
            self.markers_fresh = False
            return
        self._o._ck(lib().fsolid_update_pos_vel_area(self._o.h, self.index))

    def FluidLoads(self):
        """lodFlow = 0 (Solidbody.f90:911) then the nodal-load half of FluidVolumeForce_ (:945-967) from self.v_Eforce."""
        self._o.flush()
        self._o._ck(lib().fsolid_fluid_loads(self._o.h, self.index))
        self.count_Interp = 1

    def set_lodFlow(self, lod: np.ndarray):
        lod = np.ascontiguousarray(lod, float).reshape(-1)
        assert lod.size == self.gEQ
        self._o._ck(lib().fsolid_set_lodflow(self._o.h, self.index, _dp(lod)))

    def structure(self, time: float, isubstep: int, deltat: float, subdeltat: float):
        self._o.flush()
        self._o._ck(lib().fsolid_structure(self._o.h, self.index, float(time), int(isubstep), float(deltat), float(subdeltat)))

    def _get(self, what: int, shape):
        self._o.flush()
        out = np.zeros(shape)
        self._o._ck(lib().fsolid_get(self._o.h, self.index, what, _dp(out)))
        return out

    pos = property(lambda s: s._get(0, (s.nND, 6)))
    dsp = property(lambda s: s._get(1, (s.nND, 6)))
    vel = property(lambda s: s._get(2, (s.nND, 6)))
    acc = property(lambda s: s._get(3, (s.nND, 6)))
    lodFlow = property(lambda s: s._get(4, (s.nND, 6)))
    lodInte = property(lambda s: s._get(5, (s.nND, 6)))
    FishInfo = property(lambda s: s._get(6, (4,)))
    triads = property(lambda s: s._get(7, (s.nEL, 3, 3, 3)))       # [element][ee|n1|n2][i][j]
    mss = property(lambda s: s._get(8, (s.nND, 3)))
    strainEnergy = property(lambda s: s._get(9, (s.nEL, 2)))
    m_property = property(lambda s: s._get(10, (s.nEL, 8)))
    x1 = property(lambda s: s._get(11, (s.nEL, 12)))
    kinematics = property(lambda s: s._get(12, (4, 3)))             # XYZ, AoA, UVW, WWW3
    lodGrav = property(lambda s: s._get(13, (s.nND, 6)))
    lengths = property(lambda s: s._get(14, (s.nEL, 3)))            # len0, len1, geoFRM


class Plate:
    """Adapter with the shape block_comm.tree_collision_streaming_IBM_FEM expects of a carried body (.body,
    UpdatePosVelArea(), FluidVolumeForce(), structure()): one VBodies(iFish) backed by the C++ beam solver."""

    def __init__(self, body: Body, owner: "SolidBodies" = None):
        self.body = body
        self.host_seconds = 0.0   # wall time spent in the structural sub-steps (bench.py reports it)
        self.owner = owner           # block_comm: one call per step does the host work of all carried bodies (threads over bodies)

    def UpdatePosVelArea(self):
        self.body.UpdatePosVelArea()

    def FluidVolumeForce(self):
        self.body.FluidLoads()

    def structure(self, time: float, isubstep: int, deltat: float, subdeltat: float):
        t0 = _time.perf_counter()
        self.body.structure(time, isubstep, deltat, subdeltat)
        self.host_seconds += _time.perf_counter() - t0


class SolidBodies:
    """main.f90:29-48 for the structural side.  Relative mesh names in inFlow.dat resolve against `cwd`."""

    def __init__(self, inflow_path: str, rootBC: Sequence[int], cwd: str | None = None):
        bc = (ctypes.c_int * 6)(*[int(b) for b in rootBC])
        h = ctypes.c_int(0)
        old = os.getcwd()
        try:
            if cwd:
                os.chdir(cwd)
            rc = lib().fsolid_open(os.fsencode(inflow_path), bc, ctypes.byref(h))
        finally:
            os.chdir(old)
        if rc != 0:
            raise SolidError(lib().fsolid_last_error().decode())
        self.h = h.value
        fl = (ctypes.c_double * 16)()
        self._ck(lib().fsolid_flow(self.h, fl))
        (self.Lref, self.Uref, self.Tref, self.Aref, self.Fref, self.Eref, self.Pref, self.nu, self.Asfac, self.Lchod, self.Lspan, self.AR,
         self.denIn) = list(fl)[:13]
        self.ntolLBM, self.dtolLBM, self.numsubstep = int(fl[13]), fl[14], int(fl[15])
        self.VBodies = [Body(self, i) for i in range(lib().fsolid_nfish(self.h))]
        self.plates = [Plate(b, self) for b in self.VBodies]
        self.host_seconds = 0.0
        self._pending = []

    def _ck(self, rc: int):
        if rc != 0:
            raise SolidError(lib().fsolid_last_error().decode())

    def Solver(self, time: float, isubstep: int, deltat: float, subdeltat: float):
        """Solver, Solidbody.f90:386-398: every body, in parallel over the bodies as the reference's OpenMP loop."""
        self.flush()
        t0 = _time.perf_counter()
        self._ck(lib().fsolid_solver(self.h, float(time), int(isubstep), float(deltat), float(subdeltat)))
        self.host_seconds += _time.perf_counter() - t0

    def advance_later(self, bodies: Sequence[int], time: float, numsubstep: int, deltat: float):
        """advance(), postponed until something needs its result (the next UpdatePosVelArea / structure / state query of any
        body, write, close) -- so that whatever device work the caller queues in the meantime (a son's father<->son transfers,
        the father's next collide-stream) runs while the beams are solved.  v_Eforce must stay untouched until then."""
        self._pending.append((list(bodies), float(time), int(numsubstep), float(deltat)))

    def flush(self):
        while self._pending:
            self.advance(*self._pending.pop(0))

    def advance(self, bodies: Sequence[int], time: float, numsubstep: int, deltat: float):
        """The host work of one step for `bodies` (indices), one thread per body: nodal loads from v_Eforce, the structural
        sub-steps, UpdatePosVelArea_ for the next step (its call at the top of the next step then returns at once).
        Note for writers: v_Exyz / v_Evel then already hold the COMING step's markers, whereas the reference's
        Write_solid_v_bodies, called after the step, still sees the markers the step used (they lag the beam by one step
        there).  The stand-in driver (harness/), which produces the reference's files, keeps the reference's order."""
        t0 = _time.perf_counter()
        ids = (ctypes.c_int * len(bodies))(*bodies)
        self._ck(lib().fsolid_advance(self.h, len(bodies), ids, float(time), int(numsubstep), float(deltat)))
        for i in bodies:
            self.VBodies[i].count_Interp = 1
            self.VBodies[i].markers_fresh = True
        self.host_seconds += _time.perf_counter() - t0

    def write(self, what: int, time: float, cwd: str):
        self.flush()
        old = os.getcwd()
        try:
            os.chdir(cwd)
            self._ck(lib().fsolid_write(self.h, what, float(time)))
        finally:
            os.chdir(old)

    def close(self):
        self._pending = []
        if self.h:
            lib().fsolid_close(self.h)
            self.h = 0
