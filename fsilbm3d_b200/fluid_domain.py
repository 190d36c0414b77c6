"""LBMBlock: host mirror of `type LBMBlock` (FluidDomain.f90:17-56) over the C ABI.

Method names are the reference's type-bound procedures (FluidDomain.f90:40-55).  Two ways to step:
  * the fused path the library is built for: `collide_stream()` (= calculate_macro_quantities +
    ResetVolumeForce + add_volume_force + collision + halfwayBCset + streaming +
    set_boundary_conditions, LBMBlockComm.f90:285-303 minus IBM_FEM);
  * the reference's call granularity, pass by pass (`calculate_macro_quantities()`, `collision()`,
    `streaming()` ...), each an un-fused kernel over device-resident den/uuu/force.
Arrays cross the boundary in the reference's Fortran layout: fIn (19,X,Y,Z) with z fastest.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from ._lib import CFlow, check, ensure_init, lib
from .flow_condition import FlowCondType


def _cflow(fl: FlowCondType) -> CFlow:
    c = CFlow()
    c.nu, c.denIn = fl.nu, fl.denIn
    c.uvwIn[:] = list(fl.uvwIn)
    c.shearRateIn[:] = list(fl.shearRateIn)
    c.velocityKind = fl.velocityKind
    c.volumeForceIn[:] = list(fl.volumeForceIn)
    c.volumeForceAmp, c.volumeForceFreq, c.volumeForcePhi = fl.volumeForceAmp, fl.volumeForceFreq, fl.volumeForcePhi
    c.Uref = fl.Uref
    return c


class LBMBlock:
    def __init__(self, xDim: int, yDim: int, zDim: int, dh: float = 1.0, xmin: float = 0.0, ymin: float = 0.0,
                 zmin: float = 0.0, BndConds: Sequence[int] = (301,) * 6, iCollidModel: int = 1,
                 params: Sequence[float] = (0.0,) * 10, flow: Optional[FlowCondType] = None,
                 xOffset: int = 0, xLocal: Optional[int] = None, device: int = 0):
        ensure_init(device)
        self.flow = flow or FlowCondType()
        self.xDim, self.yDim, self.zDim, self.dh = xDim, yDim, zDim, dh
        self.xmin, self.ymin, self.zmin = xmin, ymin, zmin
        self.BndConds = tuple(int(b) for b in BndConds)
        self.iCollidModel = iCollidModel
        self.xOffset = xOffset
        self.xLocal = xDim if xLocal is None else xLocal
        self.blktime = 0.0
        self.volumeForce = np.zeros(3)
        h = C.c_int(-1)
        cf = _cflow(self.flow)
        check(lib().fsilbm_block_create(xDim, yDim, zDim, xOffset, self.xLocal, dh, xmin, ymin, zmin,
                                        (C.c_int * 6)(*self.BndConds), iCollidModel, (C.c_double * 10)(*params),
                                        C.byref(cf), C.byref(h)))
        self._h = h.value

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().fsilbm_block_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up ---------------------------------------------------------------------------------
    def initialise(self, time: float = 0.0):
        """initialise_, FluidDomain.f90:433."""
        check(lib().fsilbm_block_initialise(self._h, time))
        self.blktime = time

    def _get(self, what: int) -> float:
        v = C.c_double()
        check(lib().fsilbm_block_get(self._h, what, C.byref(v)))
        return v.value

    @property
    def tau(self): return self._get(0)
    @property
    def Omega(self): return self._get(1)
    @property
    def Omega2(self): return self._get(2)

    def set_blktime(self, t: float):
        """LBMblks(:)%blktime = time, main.f90:97."""
        self.blktime = t
        check(lib().fsilbm_block_set_time(self._h, t))

    # ---- data across the boundary ------------------------------------------------------------------
    @property
    def shape(self): return (self.xLocal, self.yDim, self.zDim)

    def upload_fIn(self, fIn: np.ndarray):
        a = np.ascontiguousarray(fIn, dtype=np.float64)
        assert a.shape == (19,) + self.shape, a.shape
        check(lib().fsilbm_block_upload_fIn(self._h, a.ctypes.data))

    def download_fIn(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.empty((19,) + self.shape) if out is None else out
        check(lib().fsilbm_block_download_fIn(self._h, a.ctypes.data))
        return a

    @property
    def fIn(self) -> np.ndarray:
        return self.download_fIn()

    def download_macro(self):
        """den, uuu from the current fIn (calculate_macro_quantities_, FluidDomain.f90:1128)."""
        den = np.empty(self.shape)
        uuu = np.empty((3,) + self.shape)
        check(lib().fsilbm_block_download_macro(self._h, den.ctypes.data, uuu.ctypes.data))
        return den, uuu

    def download_macro_async(self, den: np.ndarray, uuu: np.ndarray):
        """den, uuu into the given (pinned) host arrays without waiting; valid after download_wait() or sync()."""
        assert den.shape == self.shape and uuu.shape == (3,) + self.shape and den.flags.c_contiguous and uuu.flags.c_contiguous
        check(lib().fsilbm_block_download_macro_async(self._h, den.ctypes.data, uuu.ctypes.data))

    def download_wait(self):
        check(lib().fsilbm_block_download_wait(self._h))

    def write_flow_window_async(self, out: np.ndarray, offsetOutput: int = 0, outputtype: int = 1):
        """OUTtmp of write_flow_ (FluidDomain.f90:1640-1699: p,u,v,w as real(4), C [nfields][nx][ny][nz] over the output window)
        into the given (pinned) float32 array without waiting; valid after download_wait() or sync()."""
        assert out.dtype == np.float32 and out.flags.c_contiguous
        # the library fills nfields * (window of the local slab) floats: the array must hold exactly that (see flow_io.flow_window)
        from .flow_io import flow_window
        _, nx, ny, nz = flow_window(self, offsetOutput)
        want = (13 if outputtype >= 2 else 4) * nx * ny * nz
        if out.size != want:
            raise ValueError(f"write_flow_window_async: the staging array holds {out.size} values, the window needs {want}")
        check(lib().fsilbm_block_write_flow_window_async(self._h, offsetOutput, outputtype, out.ctypes.data))

    def download_tau_all(self) -> np.ndarray:
        """tau_all (FluidDomain.f90:51), written by the LES collision models."""
        t = np.empty(self.shape)
        check(lib().fsilbm_block_download_tau_all(self._h, t.ctypes.data))
        return t

    def calculate_turbulent_statistic(self, step: int, step_s: int):
        """calculate_turbulent_statistic_, FluidDomain.f90:1147-1172 (running means stay on the device)."""
        check(lib().fsilbm_block_turbulent_statistic(self._h, step, step_s))

    def ComputeFieldStat(self) -> np.ndarray:
        """ComputeFieldStat_, FluidDomain.f90:1739 (single slab): L2 u,v,w then Linfinity u,v,w."""
        out = (C.c_double * 6)()
        check(lib().fsilbm_block_field_stat(self._h, out))
        r = np.array(out[:])
        r[:3] = np.sqrt(r[:3] / (float(self.xLocal) * float(self.yDim) * float(self.zDim)))
        return r

    # ---- the step ---------------------------------------------------------------------------------
    def update_volume_force(self):
        out = (C.c_double * 3)()
        check(lib().fsilbm_block_update_volume_force(self._h, out))
        self.volumeForce[:] = out[:]

    def set_boundary_conditions(self):
        check(lib().fsilbm_block_set_boundary_conditions(self._h))

    def collide_stream(self):
        check(lib().fsilbm_block_collide_stream(self._h))

    def sync(self):
        check(lib().fsilbm_block_sync(self._h))

    @property
    def halo_transport(self) -> str:
        m = C.c_int(0)
        check(lib().fsilbm_block_halo_transport(self._h, C.byref(m)))
        return {0: "none (single rank)", 1: "nccl send/recv", 2: "peer stores over NVLink (CUDA IPC) fused into the edge kernels"}[m.value]

    @property
    def cuda_stream(self) -> int:
        """cudaStream_t (as an integer) the block's kernels run on, for CUDA-event timing."""
        p = C.c_void_p()
        check(lib().fsilbm_block_stream(self._h, C.byref(p)))
        return p.value or 0

    def calculate_interaction_force_begin(self, bodies, rootBC=None, dt: Optional[float] = None, collective: bool = False) -> None:
        """First half of calculate_interaction_force (Solidbody.f90:869): everything is enqueued on the device, nothing is waited
        for.  collide_stream() may follow at once (it waits for the box fields on the device); the marker forces and iterLBM are
        collected by calculate_interaction_force_wait().  `collective`: slab run in which every rank passes only the bodies near
        its slab (option ibm_force_exchange = 0); the call is then made even with none."""
        n = len(bodies)
        rootBC = self.BndConds if rootBC is None else rootBC
        self._ibm_bodies = list(bodies)
        if n == 0:
            check(lib().fsilbm_ibm_interaction_force_begin(self._h, 0, None, None, None, None, None, self.dh if dt is None else dt,
                                                           self.flow.ntolLBM if collective else 0, self.flow.dtolLBM, (C.c_int * 6)(*rootBC)))
            return
        nel = (C.c_int * n)(*[b.v_nelmts for b in bodies])
        vp = C.c_void_p
        for b in bodies:
            assert b.v_Exyz.flags.c_contiguous and b.v_Evel.flags.c_contiguous and b.v_Ea.flags.c_contiguous and b.v_Eforce.flags.c_contiguous
        ex = (vp * n)(*[b.v_Exyz.ctypes.data for b in bodies])
        ev = (vp * n)(*[b.v_Evel.ctypes.data for b in bodies])
        ea = (vp * n)(*[b.v_Ea.ctypes.data for b in bodies])
        re = (C.c_int * n)(*[1 if (b.v_move == 1 or b.iBodyModel == 2 or b.count_Interp == 0) else 0 for b in bodies])
        check(lib().fsilbm_ibm_interaction_force_begin(self._h, n, nel, ex, ev, ea, re, self.dh if dt is None else dt,
                                                       self.flow.ntolLBM, self.flow.dtolLBM, (C.c_int * 6)(*rootBC)))

    def calculate_interaction_force_wait(self) -> int:
        """Second half: waits for the call enqueued by calculate_interaction_force_begin, fills v_Eforce of its bodies, returns iterLBM."""
        bodies = getattr(self, "_ibm_bodies", [])
        n = len(bodies)
        it = C.c_int(0)
        ef = (C.c_void_p * n)(*[b.v_Eforce.ctypes.data for b in bodies]) if n else None
        check(lib().fsilbm_ibm_interaction_force_wait(self._h, n, ef, C.byref(it)))
        for b in bodies:
            b.count_Interp = 1
        self._ibm_bodies = []
        return it.value

    def calculate_interaction_force(self, bodies, rootBC=None, dt: Optional[float] = None, collective: bool = False) -> int:
        """calculate_interaction_force, Solidbody.f90:869, as one blocking call; returns iterLBM."""
        self.calculate_interaction_force_begin(bodies, rootBC, dt, collective)
        return self.calculate_interaction_force_wait()

    def download_stencil(self, body_index: int, nelmts: int):
        Ei = np.empty((nelmts, 12), dtype=np.int16)
        Ew = np.empty((nelmts, 12), dtype=np.float32)
        check(lib().fsilbm_ibm_download_stencil(self._h, body_index, Ei.ctypes.data, Ew.ctypes.data))
        return Ei, Ew

    def step(self, bodies=(), rootBC=None) -> int:
        """LBMBlockComm.f90:283-305 for a block without sons (FEM Solver excluded)."""
        self.update_volume_force()
        it = self.calculate_interaction_force(list(bodies), rootBC) if len(bodies) else 0
        self.collide_stream()
        return it

    # ---- the reference's pass-by-pass granularity (un-fused kernels) ---------------------------------
    def calculate_macro_quantities(self): check(lib().fsilbm_block_pass_macro(self._h))
    def ResetVolumeForce(self): check(lib().fsilbm_block_pass_reset_volume_force(self._h))
    def add_volume_force(self): check(lib().fsilbm_block_pass_add_volume_force(self._h))
    def collision(self): check(lib().fsilbm_block_pass_collision(self._h))
    def halfwayBCset(self): check(lib().fsilbm_block_pass_halfway_bc_set(self._h))
    def streaming(self): check(lib().fsilbm_block_pass_streaming(self._h))

    def download_fields(self):
        den = np.empty(self.shape)
        uuu = np.empty((3,) + self.shape)
        force = np.empty((3,) + self.shape)
        check(lib().fsilbm_block_download_fields(self._h, den.ctypes.data, uuu.ctypes.data, force.ctypes.data))
        return den, uuu, force

    def upload_fields(self, den=None, uuu=None, force=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (den, uuu, force)]
        check(lib().fsilbm_block_upload_fields(self._h, *[None if a is None else a.ctypes.data for a in arrs]))
