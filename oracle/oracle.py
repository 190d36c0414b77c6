"""ctypes front-end of the CPU oracle (oracle/fsilbm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product package fsilbm3d_b200 never does.  Method names follow the reference's
type-bound procedures (FluidDomain.f90:40-55, Solidbody.f90:51-67) so a parity test reads like a
call sequence of the reference.

PARITY UNPINNED: see the header of fsilbm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfsilbm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc; seconds)."""
    src = os.path.join(_HERE, "fsilbm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libfsilbm_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Flow(C.Structure):
    _fields_ = [
        ("nu", C.c_double), ("denIn", C.c_double),
        ("uvwIn", C.c_double * 3), ("shearRateIn", C.c_double * 3),
        ("velocityKind", C.c_int),
        ("volumeForceIn", C.c_double * 3), ("volumeForceAmp", C.c_double),
        ("volumeForceFreq", C.c_double), ("volumeForcePhi", C.c_double),
        ("Uref", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.orc_block_create.restype = vp
    L.orc_block_create.argtypes = [i, i, i, d, d, d, d, pi, i, pd, i, C.POINTER(_Flow)]
    L.orc_block_destroy.argtypes = [vp]
    for name in ("orc_block_fIn", "orc_block_uuu", "orc_block_force", "orc_block_den", "orc_block_volumeForce",
                 "orc_block_M_COLLID", "orc_block_M_FORCE"):
        getattr(L, name).restype = pd
        getattr(L, name).argtypes = [vp]
    L.orc_block_set_blktime.argtypes = [vp, d]
    L.orc_block_get.restype = d
    L.orc_block_get.argtypes = [vp, i]
    L.orc_block_initialise.restype = i
    L.orc_block_initialise.argtypes = [vp, d]
    for name in ("orc_calculate_macro_quantities", "orc_update_volume_force", "orc_add_volume_force",
                 "orc_reset_volume_force", "orc_streaming", "orc_halfway_bc_set"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [vp]
    L.orc_collision.restype = i
    L.orc_collision.argtypes = [vp]
    L.orc_set_boundary_conditions.restype = i
    L.orc_set_boundary_conditions.argtypes = [vp]
    L.orc_compute_field_stat.argtypes = [vp, pd]
    L.orc_body_create.restype = vp
    L.orc_body_create.argtypes = [i, i, i]
    L.orc_body_destroy.argtypes = [vp]
    for name in ("orc_body_Exyz", "orc_body_Evel", "orc_body_Ea", "orc_body_Eforce"):
        getattr(L, name).restype = pd
        getattr(L, name).argtypes = [vp]
    L.orc_body_Ei.restype = C.POINTER(C.c_int16)
    L.orc_body_Ei.argtypes = [vp]
    L.orc_body_Ew.restype = C.POINTER(C.c_float)
    L.orc_body_Ew.argtypes = [vp]
    L.orc_Phi.restype = d
    L.orc_Phi.argtypes = [d]
    L.orc_update_elmt_interp.restype = i
    L.orc_update_elmt_interp.argtypes = [vp, d, d, d, d, i, i, i, pi]
    L.orc_penalty_force.restype = i
    L.orc_penalty_force.argtypes = [vp, d, d, i, i, i, d, pd, pd, pd]
    L.orc_fluid_volume_force.argtypes = [vp, d, i, i, i, pd]
    L.orc_plate_nodal_loads.argtypes = [vp, pi, pi, pi, pd, pd]
    L.orc_calculate_interaction_force.restype = i
    L.orc_calculate_interaction_force.argtypes = [C.POINTER(vp), i, d, d, d, d, d, i, i, i, pd, pd, pi, d, d, i, d]
    L.orc_step.restype = i
    L.orc_step.argtypes = [vp, C.POINTER(vp), i, pi, i, d, pi]
    L.orc_omp_max_threads.restype = i
    L.orc_omp_set_threads.argtypes = [i]
    L.orc_omp_set_threads.restype = None
    L.orc_step_pre.restype = i
    L.orc_step_pre.argtypes = [vp, C.POINTER(vp), i, pi, i, d, pi]
    L.orc_step_post.restype = i
    L.orc_step_post.argtypes = [vp]
    L.orc_pair_create.restype = vp
    L.orc_pair_create.argtypes = [vp, vp, i]
    L.orc_pair_destroy.argtypes = [vp]
    L.orc_pair_get.argtypes = [vp, pi]
    L.orc_pair_extract_layer.argtypes = [vp, i]
    L.orc_pair_father_to_son.argtypes = [vp, i]
    L.orc_pair_son_to_father.argtypes = [vp]
    L.orc_block_tau_all.restype = pd
    L.orc_block_tau_all.argtypes = [vp]
    _lib = L
    return L


@dataclass
class Flow:
    """The slice of FlowCondType the hot path reads (FlowCondition.f90:11-27)."""
    nu: float = 0.1
    denIn: float = 1.0
    uvwIn: Sequence[float] = (0.0, 0.0, 0.0)
    shearRateIn: Sequence[float] = (0.0, 0.0, 0.0)
    velocityKind: int = 0
    volumeForceIn: Sequence[float] = (0.0, 0.0, 0.0)
    volumeForceAmp: float = 0.0
    volumeForceFreq: float = 0.0
    volumeForcePhi: float = 0.0
    Uref: float = 1.0
    ntolLBM: int = 1
    dtolLBM: float = 1e-10

    def _c(self) -> _Flow:
        f = _Flow()
        f.nu, f.denIn = self.nu, self.denIn
        f.uvwIn[:] = list(self.uvwIn)
        f.shearRateIn[:] = list(self.shearRateIn)
        f.velocityKind = self.velocityKind
        f.volumeForceIn[:] = list(self.volumeForceIn)
        f.volumeForceAmp, f.volumeForceFreq, f.volumeForcePhi = self.volumeForceAmp, self.volumeForceFreq, self.volumeForcePhi
        f.Uref = self.Uref
        return f


def _view(ptr, shape, dtype=np.float64):
    n = int(np.prod(shape))
    ct = {np.float64: C.c_double, np.int16: C.c_int16, np.float32: C.c_float}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
    return arr.reshape(shape)


class VirtualBody:
    """Marker state of Solidbody.f90:25-68.  Arrays are numpy views of the oracle's memory:
    v_Exyz/v_Evel/v_Eforce have shape (n,3) (= Fortran (3,n)), v_Ei/v_Ew shape (n,12)."""

    def __init__(self, nelmts: int, v_move: int = 0, iBodyModel: int = 1):
        L = lib()
        self._h = L.orc_body_create(nelmts, v_move, iBodyModel)
        self.v_nelmts = nelmts
        self.v_Exyz = _view(L.orc_body_Exyz(self._h), (nelmts, 3))
        self.v_Evel = _view(L.orc_body_Evel(self._h), (nelmts, 3))
        self.v_Ea = _view(L.orc_body_Ea(self._h), (nelmts,))
        self.v_Eforce = _view(L.orc_body_Eforce(self._h), (nelmts, 3))
        self.v_Ei = _view(L.orc_body_Ei(self._h), (nelmts, 12), np.int16)
        self.v_Ew = _view(L.orc_body_Ew(self._h), (nelmts, 12), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_body_destroy(self._h)
            self._h = None


class LBMBlock:
    """type LBMBlock (FluidDomain.f90:17-56) backed by the oracle.  fIn has shape (19,X,Y,Z)
    (= Fortran fIn(z,y,x,0:18)); uuu/force (3,X,Y,Z); den (X,Y,Z)."""

    def __init__(self, xDim, yDim, zDim, dh=1.0, xmin=0.0, ymin=0.0, zmin=0.0, BndConds=(301,) * 6,
                 iCollidModel=1, params=(0.0,) * 10, flow: Optional[Flow] = None, npsize: Optional[int] = None):
        L = lib()
        self.flow = flow or Flow()
        self.xDim, self.yDim, self.zDim, self.dh = xDim, yDim, zDim, dh
        self.xmin, self.ymin, self.zmin = xmin, ymin, zmin
        self.BndConds = tuple(int(b) for b in BndConds)
        self.iCollidModel = iCollidModel
        bc = (C.c_int * 6)(*self.BndConds)
        pr = (C.c_double * 10)(*params)
        cf = self.flow._c()
        if npsize is None:
            npsize = L.orc_omp_max_threads()
        self.npsize = npsize
        self._h = L.orc_block_create(xDim, yDim, zDim, dh, xmin, ymin, zmin, bc, iCollidModel, pr, npsize, C.byref(cf))
        if not self._h:
            raise ValueError("oracle: block rejected (dims > 32767 or unpaired periodic BC)")
        self.fIn = _view(L.orc_block_fIn(self._h), (19, xDim, yDim, zDim))
        self.uuu = _view(L.orc_block_uuu(self._h), (3, xDim, yDim, zDim))
        self.force = _view(L.orc_block_force(self._h), (3, xDim, yDim, zDim))
        self.den = _view(L.orc_block_den(self._h), (xDim, yDim, zDim))
        self.volumeForce = _view(L.orc_block_volumeForce(self._h), (3,))
        self.tau_all = _view(L.orc_block_tau_all(self._h), (xDim, yDim, zDim))
        self.blktime = 0.0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_block_destroy(self._h)
            self._h = None

    # -- reference procedures -------------------------------------------------------------------
    def initialise(self, time=0.0):
        rc = lib().orc_block_initialise(self._h, time)
        self.blktime = time
        if rc:
            raise ValueError("oracle: velocityKind must be 0 or 2")

    def set_blktime(self, t):
        self.blktime = t
        lib().orc_block_set_blktime(self._h, t)

    @property
    def tau(self): return lib().orc_block_get(self._h, 0)
    @property
    def Omega(self): return lib().orc_block_get(self._h, 1)
    @property
    def Omega2(self): return lib().orc_block_get(self._h, 2)
    @property
    def M_COLLID(self): return _view(lib().orc_block_M_COLLID(self._h), (19, 19))
    @property
    def M_FORCE(self): return _view(lib().orc_block_M_FORCE(self._h), (19, 19))

    def update_volume_force(self): lib().orc_update_volume_force(self._h)
    def calculate_macro_quantities(self): lib().orc_calculate_macro_quantities(self._h)
    def ResetVolumeForce(self): lib().orc_reset_volume_force(self._h)
    def add_volume_force(self): lib().orc_add_volume_force(self._h)

    def collision(self):
        if lib().orc_collision(self._h):
            raise ValueError("oracle: collision model not restated")

    def halfwayBCset(self): lib().orc_halfway_bc_set(self._h)
    def streaming(self): lib().orc_streaming(self._h)

    def set_boundary_conditions(self):
        rc = lib().orc_set_boundary_conditions(self._h)
        if rc:
            raise ValueError(f"oracle: set_boundary_conditions failed rc={rc}")

    def ComputeFieldStat(self):
        out = (C.c_double * 6)()
        lib().orc_compute_field_stat(self._h, out)
        return np.array(out[:])

    # -- IBM (Solidbody.f90:869-918 acting on this block's uuu/force) ------------------------------
    def calculate_interaction_force(self, bodies: List[VirtualBody], rootBC=None) -> int:
        L = lib()
        rootBC = self.BndConds if rootBC is None else rootBC
        arr = (C.c_void_p * max(1, len(bodies)))(*[b._h for b in bodies])
        bc = (C.c_int * 6)(*rootBC)
        it = L.orc_calculate_interaction_force(
            arr, len(bodies), self.dh, self.dh, self.xmin, self.ymin, self.zmin, self.xDim, self.yDim, self.zDim,
            self.uuu.ctypes.data_as(C.POINTER(C.c_double)), self.force.ctypes.data_as(C.POINTER(C.c_double)), bc,
            self.flow.denIn, self.flow.Uref, self.flow.ntolLBM, self.flow.dtolLBM)
        if it < 0:
            raise ValueError("oracle: IBM stopped (stencil out of domain or NaN)")
        return it

    def step(self, bodies: Sequence[VirtualBody] = (), rootBC=None) -> int:
        """LBMBlockComm.f90:283-305 for a block without sons (FEM Solver excluded)."""
        L = lib()
        rootBC = self.BndConds if rootBC is None else rootBC
        arr = (C.c_void_p * max(1, len(bodies)))(*[b._h for b in bodies])
        bc = (C.c_int * 6)(*rootBC)
        it = C.c_int(0)
        rc = L.orc_step(self._h, arr, len(bodies), bc, self.flow.ntolLBM, self.flow.dtolLBM, C.byref(it))
        if rc:
            raise ValueError(f"oracle: step failed rc={rc}")
        return it.value


def Phi(x: float) -> float:
    return lib().orc_Phi(x)


class CommPair:
    """type CommPair (LBMBlockComm.f90:11-18) with the father<->son transfers of :340-979."""

    def __init__(self, father: LBMBlock, son: LBMBlock, interpolateScheme: int = 1):
        self.father, self.son = father, son
        self._h = lib().orc_pair_create(father._h, son._h, interpolateScheme)
        if not self._h:
            raise ValueError("grid points do not match between fluid blocks (LBMBlockComm.f90:537-540)")
        out = (C.c_int * 36)()
        lib().orc_pair_get(self._h, out)
        v = list(out)
        self.sds, self.s, self.f, self.si, self.fi = v[0:6], v[6:12], v[12:18], v[18:24], v[24:30]
        self.dimS, self.dimF = v[30:33], v[33:36]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pair_destroy(self._h)
            self._h = None

    def extract_interpolate_layer(self, time: int): lib().orc_pair_extract_layer(self._h, time)
    def interpolation_father_to_son(self, n_timeStep: int): lib().orc_pair_father_to_son(self._h, n_timeStep)
    def deliver_son_to_father(self): lib().orc_pair_son_to_father(self._h)


class TreeNode:
    """blockTreeNode (LBMBlockComm.f90:19-25): a block, the bodies it carries, its sons with their CommPairs."""

    def __init__(self, block: LBMBlock, bodies: Sequence[VirtualBody] = (), rootBC=None):
        self.block, self.bodies, self.sons, self.comm = block, list(bodies), [], []
        self.rootBC = rootBC

    def add_son(self, node: "TreeNode", interpolateScheme: int = 1):
        self.sons.append(node)
        self.comm.append(CommPair(self.block, node.block, interpolateScheme))
        return node


def tree_collision_streaming_IBM_FEM(node: TreeNode, rootBC=None, iters: Optional[list] = None, before_ibm=None, after_ibm=None):
    """LBMBlockComm.f90:279-318 (FEM Solver excluded), recursion over the block tree.  before_ibm(node) / after_ibm(node): the
    caller's host work around IBM_FEM of each node (:287,320-338): marker update before, nodal loads and structural sub-steps
    (with the node's own blktime and dh) after."""
    L = lib()
    b = node.block
    rootBC = b.BndConds if rootBC is None else rootBC
    if before_ibm is not None:
        before_ibm(node)
    arr = (C.c_void_p * max(1, len(node.bodies)))(*[v._h for v in node.bodies])
    it = C.c_int(0)
    rc = L.orc_step_pre(b._h, arr, len(node.bodies), (C.c_int * 6)(*rootBC), b.flow.ntolLBM, b.flow.dtolLBM, C.byref(it))   # :283-288
    if rc:
        raise ValueError(f"oracle: step_pre failed rc={rc}")
    if iters is not None:
        iters.append(it.value)
    if after_ibm is not None:
        after_ibm(node)
    for pair in node.comm:
        pair.extract_interpolate_layer(1)                                 # :290
    rc = L.orc_step_post(b._h)                                            # :293-303
    if rc:
        raise ValueError(f"oracle: step_post failed rc={rc}")
    for pair in node.comm:
        pair.extract_interpolate_layer(2)                                 # :305
    for son, pair in zip(node.sons, node.comm):                           # :307-317
        for n_timeStep in range(2):
            son.block.set_blktime(son.block.blktime + float(n_timeStep) * son.block.dh)   # :311
            tree_collision_streaming_IBM_FEM(son, rootBC, iters, before_ibm, after_ibm)
            pair.interpolation_father_to_son(n_timeStep)
        pair.deliver_son_to_father()


def set_blktime_all(node: TreeNode, time: float):
    """LBMblks(:)%blktime = time, main.f90:97."""
    node.block.set_blktime(time)
    for s in node.sons:
        set_blktime_all(s, time)
