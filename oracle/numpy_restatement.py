"""numpy_restatement.py -- SECOND, independent CPU restatement of the FSILBM3D hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference is Fortran, ships no tests or golden vectors, and cannot be built in this
environment (no Fortran compiler).  This file exists to reduce the risk of a transcription error in the C oracle
(oracle/fsilbm_oracle.c): it was written separately, straight from the Fortran, with a different program
structure (whole-array numpy operations, np.roll streaming, slice-based boundary faces) but the same
floating-point evaluation order expression by expression, so the two restatements are expected to agree
bit for bit on fluid cases.  tests/golden/make_golden.py runs THIS file to produce the committed golden
vectors; tests/test_oracle_golden.py holds the C oracle to them and tests/test_gpu_golden.py the CUDA path.

Citations are file:line into /root/reference/src.  Arrays: f[q, x, y, z] (C order) == Fortran fIn(z,y,x,q).
Only tests/ and tests/golden/make_golden.py import this module.
"""
from __future__ import annotations

import math

import numpy as np

# ConstParams.f90:11-20
EE = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
               [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1],
               [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1]], dtype=np.int64)
OPPO = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])
POSITIVEDIRS = np.array([1, 3, 5, 7, 8, 11, 12, 15, 16])
NEGATIVEDIRS = np.array([2, 4, 6, 10, 9, 14, 13, 18, 17])
# ConstParams.f90:22-25
WT = np.array([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12)
S0, S1, S2, S4, S10, S16 = 0.0, 1.19, 1.4, 1.2, 1.4, 1.98   # ConstParams.f90:28
PI = 3.141592653589793                                     # ConstParams.f90:36
CS2 = 1.0 / 3.0                                            # ConstParams.f90:39
# incoming sets and mirror sources per face (FluidDomain.f90:645,729,813,897,981,1065 / 697,781,865,949,1033,1117)
FACE_IN = [[1, 7, 9, 11, 13], [2, 8, 10, 12, 14], [3, 7, 8, 15, 17], [4, 9, 10, 16, 18], [5, 11, 12, 15, 16], [6, 13, 14, 17, 18]]
FACE_MIRROR = [[2, 8, 10, 12, 14], [1, 7, 9, 11, 13], [4, 9, 10, 16, 18], [3, 7, 8, 15, 17], [6, 13, 14, 17, 18], [5, 11, 12, 15, 16]]


def f_eq(density, v1, v2, v3):
    """calculate_distribution_funcion, FluidDomain.f90:1827-1834 (broadcasts)."""
    uSqr = 0.0 + v1 * v1
    uSqr = uSqr + v2 * v2
    uSqr = uSqr + v3 * v3
    out = []
    for q in range(19):
        uxyz = v1 * float(EE[q, 0]) + v2 * float(EE[q, 1]) + v3 * float(EE[q, 2])
        out.append(WT[q] * density * (1.0 + 3.0 * uxyz + 4.5 * uxyz * uxyz - 1.5 * uSqr))
    return out


class Flow:
    def __init__(self, nu=0.1, denIn=1.0, uvwIn=(0.0, 0.0, 0.0), shearRateIn=(0.0, 0.0, 0.0), velocityKind=0,
                 volumeForceIn=(0.0, 0.0, 0.0), volumeForceAmp=0.0, volumeForceFreq=0.0, volumeForcePhi=0.0, Uref=1.0,
                 ntolLBM=1, dtolLBM=1e-10, numsubstep=1):
        self.nu, self.denIn, self.uvwIn, self.shearRateIn, self.velocityKind = nu, denIn, tuple(uvwIn), tuple(shearRateIn), velocityKind
        self.volumeForceIn, self.volumeForceAmp, self.volumeForceFreq, self.volumeForcePhi = tuple(volumeForceIn), volumeForceAmp, volumeForceFreq, volumeForcePhi
        self.Uref, self.ntolLBM, self.dtolLBM, self.numsubstep = Uref, ntolLBM, dtolLBM, numsubstep


class Block:
    """type LBMBlock, FluidDomain.f90:17-56."""

    def __init__(self, xDim, yDim, zDim, dh=1.0, xmin=0.0, ymin=0.0, zmin=0.0, BndConds=(301,) * 6, iCollidModel=1,
                 params=(0.0,) * 10, flow=None):
        self.X, self.Y, self.Z, self.dh = xDim, yDim, zDim, dh
        self.mins = (xmin, ymin, zmin)
        self.bc = tuple(BndConds)
        self.model, self.params = iCollidModel, tuple(params)
        self.flow = flow or Flow()
        self.periodic = [0, 0, 0]
        for i in range(3):   # check_periodic_boundary_, FluidDomain.f90:110-125
            if self.bc[2 * i] == 301 or self.bc[2 * i + 1] == 301:
                if self.bc[2 * i] != self.bc[2 * i + 1]:
                    raise ValueError("Periodic boundaries must apper in pairs")
                self.periodic[i] = 1
        dims = (xDim, yDim, zDim)
        self.maxs = [self.mins[i] + dh * (dims[i] - 1) for i in range(3)]   # FluidDomain.f90:94-96 (non-periodic value)
        self.f = np.zeros((19, xDim, yDim, zDim))
        self.den = np.zeros((xDim, yDim, zDim))
        self.uuu = np.zeros((3, xDim, yDim, zDim))
        self.force = np.zeros((3, xDim, yDim, zDim))
        self.tau_all = np.zeros((xDim, yDim, zDim))
        self.hw = [None] * 6
        self.volumeForce = [0.0, 0.0, 0.0]
        self.blktime = 0.0

    # ---- helpers ---------------------------------------------------------------------------------------
    def evaluate_velocity(self, time, zC, yC, xC):
        """FluidDomain.f90:1791-1824."""
        fl = self.flow
        vin, sr = fl.uvwIn, fl.shearRateIn
        if fl.velocityKind == 0:
            return (vin[0] + 0 * sr[0] + yC * sr[1] + zC * sr[2],
                    vin[1] + xC * sr[0] + 0 * sr[1] + zC * sr[2],
                    vin[2] + xC * sr[0] + yC * sr[1] + 0 * sr[2])
        if fl.velocityKind == 2:
            amp, freq, phi = sr
            return (vin[0] + amp * math.cos(2 * PI * freq * time + phi / 180.0 * PI), vin[1] + 0 * (xC + yC + zC), vin[2] + 0 * (xC + yC + zC))
        raise ValueError("velocityKind")

    def coords(self):
        x = self.mins[0] + self.dh * np.arange(self.X, dtype=float)[:, None, None]
        y = self.mins[1] + self.dh * np.arange(self.Y, dtype=float)[None, :, None]
        z = self.mins[2] + self.dh * np.arange(self.Z, dtype=float)[None, None, :]
        return x, y, z

    # ---- set-up ----------------------------------------------------------------------------------------
    def initialise(self, time=0.0):
        """initialise_, FluidDomain.f90:433-545."""
        self.blktime = time
        self.tau = self.flow.nu / (self.dh * CS2) + 0.5      # :452
        self.Omega = 1.0 / self.tau
        self.tau_all[...] = self.tau
        self.Omega2 = 0.0
        if self.model == 2:                                   # :458-464
            lam = self.params[0]
            tmp = (lam * 4.0 - 1.0) * self.Omega + 2.0
            self.Omega2 = 2.0 * (2.0 - self.Omega) / tmp
        if self.model == 3:
            self.calculate_MRT_params()
        x, y, z = self.coords()
        v1, v2, v3 = self.evaluate_velocity(time, z, y, x)
        shape = (self.X, self.Y, self.Z)
        v1, v2, v3 = (np.broadcast_to(np.asarray(v, float), shape) for v in (v1, v2, v3))
        self.uuu[0], self.uuu[1], self.uuu[2] = v1, v2, v3   # :535
        self.den[...] = self.flow.denIn                      # :536
        for q, d in enumerate(f_eq(self.flow.denIn, v1, v2, v3)):
            self.f[q] = d
        self.force[...] = 0.0

    def calculate_MRT_params(self):
        """FluidDomain.f90:466-522; sums ascending in the inner index, starting from zero."""
        M = np.zeros((19, 19))
        for I in range(19):
            e1, e2, e3 = (float(v) for v in EE[I])
            sq = float(EE[I, 0] ** 2 + EE[I, 1] ** 2 + EE[I, 2] ** 2)
            M[0, I] = 1.0
            M[1, I] = 19.0 * sq - 30.0
            M[2, I] = (21.0 * (sq * sq) - 53.0 * sq + 24.0) / 2.0
            M[3, I], M[5, I], M[7, I] = e1, e2, e3
            M[4, I], M[6, I], M[8, I] = (5.0 * sq - 9.0) * e1, (5.0 * sq - 9.0) * e2, (5.0 * sq - 9.0) * e3
            M[9, I] = 3.0 * (e1 * e1) - sq
            M[10, I] = (3.0 * sq - 5.0) * (3.0 * (e1 * e1) - sq)
            M[11, I] = e2 * e2 - e3 * e3
            M[12, I] = (3.0 * sq - 5.0) * (e2 * e2 - e3 * e3)
            M[13, I], M[14, I], M[15, I] = e1 * e2, e2 * e3, e3 * e1
            M[16, I] = (e2 * e2 - e3 * e3) * e1
            M[17, I] = (e3 * e3 - e1 * e1) * e2
            M[18, I] = (e1 * e1 - e2 * e2) * e3

        def matmul(A, B):
            C = np.zeros((19, 19))
            for k in range(19):
                C = C + A[:, k:k + 1] * B[k:k + 1, :]
            return C
        MI = M.T.copy()
        MM = matmul(M, MI)
        for I in range(19):
            MI[:, I] = MI[:, I] / MM[I, I]
        S = np.diag([S0, S1, S2, S0, S4, S0, S4, S0, S4, self.Omega, S10, self.Omega, S10, self.Omega, self.Omega, self.Omega, S16, S16, S16])
        self.M_COLLID = matmul(matmul(MI, S), M)
        self.M_FORCE = np.eye(19) - 0.5 * self.M_COLLID

    # ---- passes ----------------------------------------------------------------------------------------
    def update_volume_force(self):
        """FluidDomain.f90:1174-1180."""
        fl = self.flow
        self.volumeForce[0] = fl.volumeForceIn[0] + fl.volumeForceAmp * math.sin(2.0 * PI * fl.volumeForceFreq * self.blktime + fl.volumeForcePhi / 180.0 * PI)
        self.volumeForce[1] = fl.volumeForceIn[1]
        self.volumeForce[2] = fl.volumeForceIn[2]

    def calculate_macro_quantities(self):
        """FluidDomain.f90:1128-1145."""
        f = self.f
        den = f[0].copy()
        for q in range(1, 19):
            den = den + f[q]
        self.den = den
        for k in range(3):
            m = None
            for q in range(19):
                e = EE[q, k]
                if e == 0:
                    continue
                t = f[q] if e > 0 else -f[q]
                m = t.copy() if m is None else m + t
            self.uuu[k] = (m + 0.5 * self.volumeForce[k] * self.dh) / den

    def ResetVolumeForce(self):
        self.force[...] = 0.0

    def add_volume_force(self):
        for k in range(3):
            self.force[k] = self.force[k] + self.volumeForce[k]

    def collision(self):
        """collision_, FluidDomain.f90:1208-1238 (SRT, TRT, MRT)."""
        f, u, F, den = self.f, self.uuu, self.force, self.den
        dt3 = 3.0 * self.dh
        uSqr = 0.0 + u[0] * u[0]
        uSqr = uSqr + u[1] * u[1]
        uSqr = uSqr + u[2] * u[2]
        fEq, Flb = [None] * 19, [None] * 19
        for q in range(19):
            e1, e2, e3 = (float(v) for v in EE[q])
            uxyz = u[0] * e1 + u[1] * e2 + u[2] * e3
            fEq[q] = WT[q] * den * ((1.0 - 1.5 * uSqr) + uxyz * (3.0 + 4.5 * uxyz)) - f[q]
            Flb[q] = dt3 * WT[q] * ((e1 - u[0] + 3.0 * uxyz * e1) * F[0] + (e2 - u[1] + 3.0 * uxyz * e2) * F[1] + (e3 - u[2] + 3.0 * uxyz * e3) * F[2])
        Om, Om2 = self.Omega, self.Omega2
        if self.model == 1:
            for q in range(19):
                f[q] = f[q] + Om * fEq[q] + (1.0 - 0.5 * Om) * Flb[q]
        elif self.model == 2:
            fEq[0] = Om * fEq[0] + (1.0 - 0.5 * Om) * Flb[0]
            sym, asym = {}, {}
            for p, n in zip(POSITIVEDIRS, NEGATIVEDIRS):
                sym[p] = 0.5 * Om * (fEq[p] + fEq[n]) + (0.5 - 0.25 * Om) * (Flb[p] + Flb[n])
                asym[p] = 0.5 * Om2 * (fEq[p] - fEq[n]) + (0.5 - 0.25 * Om2) * (Flb[p] - Flb[n])
            for p, n in zip(POSITIVEDIRS, NEGATIVEDIRS):
                fEq[p] = sym[p] + asym[p]
                fEq[n] = sym[p] - asym[p]
            for q in range(19):
                f[q] = f[q] + fEq[q]
        elif self.model == 3:
            for i in range(19):
                mc = np.zeros_like(den)
                mf = np.zeros_like(den)
                for k in range(19):
                    mc = mc + self.M_COLLID[i, k] * fEq[k]
                for k in range(19):
                    mf = mf + self.M_FORCE[i, k] * Flb[k]
                f[i] = f[i] + mc + mf
        elif self.model in (11, 14, 15):
            fneq = [-e for e in fEq]
            if self.model == 11:
                omega = self._smag(fneq, den)
            elif self.model == 14:
                omega = self._wale(fneq, den)
            else:
                omega = self._vrem()
            for q in range(19):
                f[q] = f[q] + omega * fEq[q] + (1.0 - 0.5 * omega) * Flb[q]
        else:
            raise ValueError("model")

    # ---- LES closures contained in collision_ (FluidDomain.f90:1265-1281, 1311-1424, 1435-1507) ------------
    @staticmethod
    def _Q(fn):
        Q11 = fn[1] + fn[2] + fn[7] + fn[8] + fn[9] + fn[10] + fn[11] + fn[12] + fn[13] + fn[14]
        Q22 = fn[3] + fn[4] + fn[7] + fn[8] + fn[9] + fn[10] + fn[15] + fn[16] + fn[17] + fn[18]
        Q33 = fn[5] + fn[6] + fn[11] + fn[12] + fn[13] + fn[14] + fn[15] + fn[16] + fn[17] + fn[18]
        Q12 = fn[7] - fn[8] - fn[9] + fn[10]
        Q13 = fn[11] - fn[12] - fn[13] + fn[14]
        Q23 = fn[15] - fn[16] - fn[17] + fn[18]
        return Q11, Q22, Q33, Q12, Q13, Q23

    def _grad(self, k, axis):
        """d uuu(k) / d axis: center_diff inside, onesid_diff on the first/last plane; "invdh" is dh (sic, :1322,1444)."""
        u = np.moveaxis(self.uuu[k], axis, 0)
        g = np.empty_like(u)
        invdh = self.dh
        g[1:-1] = (u[2:] - u[:-2]) * invdh
        g[0] = (-3.0 * u[0] + 4.0 * u[1] - u[2]) * invdh
        g[-1] = (-3.0 * u[-1] + 4.0 * u[-2] - u[-3]) * invdh
        return np.moveaxis(g, 0, axis)

    def _smag(self, fneq, rho):
        CsmagConst = 2.0 * 0.17 * 0.17 * math.sqrt(2.0) * 9.0
        Q11, Q22, Q33, Q12, Q13, Q23 = self._Q(fneq)
        Qq = Q11 * Q11 + Q22 * Q22 + Q33 * Q33 + 2.0 * (Q12 * Q12 + Q13 * Q13 + Q23 * Q23)
        tau_t = np.sqrt(self.tau * self.tau + CsmagConst * np.sqrt(Qq) / rho)
        self.tau_all = 0.5 * (self.tau + tau_t)
        return 2.0 / (self.tau + tau_t)

    def _wale(self, fneq, rho):
        invdh = self.dh
        Q11, Q22, Q33, Q12, Q13, Q23 = self._Q(fneq)
        t = self.tau_all
        S11, S22, S33 = -1.5 * invdh * Q11 / (rho * t), -1.5 * invdh * Q22 / (rho * t), -1.5 * invdh * Q33 / (rho * t)
        S12, S13, S23 = -1.5 * invdh * Q12 / (rho * t), -1.5 * invdh * Q13 / (rho * t), -1.5 * invdh * Q23 / (rho * t)
        S = S11 * S11 + S22 * S22 + S33 * S33 + 2.0 * (S12 * S12 + S13 * S13 + S23 * S23)
        ox = 0.5 * (self._grad(2, 1) - self._grad(1, 2))
        oy = 0.5 * (self._grad(0, 2) - self._grad(2, 0))
        oz = 0.5 * (self._grad(1, 0) - self._grad(0, 1))
        O12, O13, O23 = -0.5 * oz, 0.5 * oy, -0.5 * ox
        O = 2.0 * (O12 * O12 + O23 * O23 + O13 * O13)
        SO11 = -(0.0 + S11 * S11 * O12 * O12 + S11 * S11 * O13 * O13 + 0.0 + S12 * S12 * O12 * O12 + S12 * S12 * O13 * O13 + 0.0 + S13 * S13 * O12 * O12 + S13 * S13 * O13 * O13)
        SO22 = -(S12 * S12 * O12 * O12 + 0.0 + S12 * S12 * O23 * O23 + S22 * S22 * O12 * O12 + 0.0 + S22 * S22 * O23 * O23 + S23 * S23 * O12 * O12 + 0.0 + S23 * S23 * O23 * O23)
        SO33 = -(S13 * S13 * O13 * O13 + S13 * S13 * O23 * O23 + 0.0 + S23 * S23 * O13 * O13 + S23 * S23 * O23 * O23 + 0.0 + S33 * S33 * O13 * O13 + S33 * S33 * O23 * O23 + 0.0)
        SO12 = -(0.0 + 0.0 + S11 * S12 * O13 * O23 + 0.0 + 0.0 + S12 * S22 * O13 * O23 + 0.0 + 0.0 + S13 * S23 * O13 * O23)
        SO13 = (0.0 + S11 * S13 * O12 * O23 + 0.0 + 0.0 + S12 * S23 * O12 * O23 + 0.0 + 0.0 + S13 * S33 * O12 * O23 + 0.0)
        SO23 = -(S12 * S13 * O12 * O13 + 0.0 + 0.0 + S22 * S23 * O12 * O13 + 0.0 + 0.0 + S23 * S33 * O12 * O13 + 0.0 + 0.0)
        SO = SO11 + SO22 + SO33 + 2.0 * (SO12 + SO13 + SO23)
        SdSd = (S * S + O * O) / 6.0 + 2.0 * S * O / 3.0 + 2.0 * SO
        with np.errstate(all="ignore"):
            OP = SdSd ** 1.5 / (S ** 2.5 + SdSd ** 1.25)
        OP = np.where(np.isfinite(OP) & ~(OP < 0.0), OP, 0.0)
        tau__ = (self.flow.nu + 0.5 * 0.5 * OP * self.dh * self.dh) / (self.dh * CS2) + 0.5
        self.tau_all = tau__
        return 1.0 / tau__

    def _vrem(self):
        a = [[0.5 * self._grad(i, j) for j in range(3)] for i in range(3)]
        b = [[a[i][j] * a[i][j] for j in range(3)] for i in range(3)]
        aa = b[0][0] + b[0][1] + b[0][2] + b[1][0] + b[1][1] + b[1][2] + b[2][0] + b[2][1] + b[2][2]
        d12 = a[0][0] * a[1][0] + a[0][1] * a[1][1] + a[0][2] * a[1][2]
        d13 = a[0][0] * a[2][0] + a[0][1] * a[2][1] + a[0][2] * a[2][2]
        d23 = a[1][0] * a[2][0] + a[1][1] * a[2][1] + a[1][2] * a[2][2]
        r1, r2, r3 = b[0][0] + b[0][1] + b[0][2], b[1][0] + b[1][1] + b[1][2], b[2][0] + b[2][1] + b[2][2]
        bb = r1 * r2 - d12 * d12 + r1 * r3 - d13 * d13 + r2 * r3 - d23 * d23
        with np.errstate(all="ignore"):
            OP = np.sqrt(bb / aa)
        OP = np.where(np.isfinite(OP), OP, 0.0)
        tau__ = (self.flow.nu + 2.5 * 0.17 * 0.17 * OP * self.dh * self.dh) / (self.dh * CS2) + 0.5
        self.tau_all = tau__
        return 1.0 / tau__

    def halfwayBCset(self):
        """FluidDomain.f90:567-614 (copies only where the stash exists, which the reference guarantees by call order)."""
        for face in range(6):
            if self.bc[face] in (203, 204) and self.hw[face] is not None:
                self.hw[face] = self.f[self._layer(face, 0)].copy()

    def streaming(self):
        """streaming_, FluidDomain.f90:1514-1625: periodic shift of population q by e_q on every axis."""
        for q in range(19):
            self.f[q] = np.roll(self.f[q], shift=(int(EE[q, 0]), int(EE[q, 1]), int(EE[q, 2])), axis=(0, 1, 2))

    def _layer(self, face, layer):
        """index tuple selecting f[:, <face layer>] (all q)"""
        axis, hi = face // 2, face % 2
        n = (self.X, self.Y, self.Z)[axis]
        idx = [slice(None)] * 4
        idx[1 + axis] = (n - 1 - layer) if hi else layer
        return tuple(idx)

    def _face_coords(self, face, halfway=False):
        axis, hi = face // 2, face % 2
        x = self.mins[0] + self.dh * np.arange(self.X, dtype=float)
        y = self.mins[1] + self.dh * np.arange(self.Y, dtype=float)
        z = self.mins[2] + self.dh * np.arange(self.Z, dtype=float)
        wall = self.maxs[axis] if hi else self.mins[axis]
        if halfway:
            wall = wall + self.dh * 0.5 if hi else wall - self.dh * 0.5
        if axis == 0:
            return wall, y[:, None], z[None, :]
        if axis == 1:
            return x[:, None], wall, z[None, :]
        return x[:, None], y[None, :], wall

    def set_boundary_conditions(self):
        """set_boundary_conditions_, FluidDomain.f90:616-1126, faces in source order on the live array."""
        f = self.f
        for face in range(6):
            code = self.bc[face]
            I, Mi = FACE_IN[face], FACE_MIRROR[face]
            L1, L2, L3 = self._layer(face, 0), self._layer(face, 1), self._layer(face, 2)
            m1, m2 = tuple(L1[1:]), tuple(L2[1:])
            shape2 = f[L1].shape[1:]
            if code in (301, 0, 1):
                continue
            if code in (101, 102, 202, 204):
                xC, yC, zC = self._face_coords(face, halfway=(code == 204))
                v = [np.broadcast_to(np.asarray(c, float), shape2) for c in self.evaluate_velocity(self.blktime, zC, yC, xC)]
            if code == 101:
                for q, d in enumerate(f_eq(self.flow.denIn, v[0], v[1], v[2])):
                    f[(q,) + m1] = d
            elif code == 102:
                fe = f_eq(self.flow.denIn, v[0], v[1], v[2])
                fei = f_eq(self.den[m2], self.uuu[(0,) + m2], self.uuu[(1,) + m2], self.uuu[(2,) + m2])
                for q in I:
                    f[(q,) + m1] = fe[q] + (f[(q,) + m2] - fei[q])
            elif code == 103:
                for q in I:
                    f[(q,) + m1] = f[(q,) + m2]
            elif code == 104:
                m3 = tuple(L3[1:])
                for q in I:
                    f[(q,) + m1] = 2.0 * f[(q,) + m2] - f[(q,) + m3]
            elif code == 201:
                tmp = [f[(int(OPPO[q]),) + m1].copy() for q in I]
                for q, t in zip(I, tmp):
                    f[(q,) + m1] = t
            elif code == 203:
                if self.hw[face] is None:
                    self.hw[face] = np.zeros((19,) + shape2)       # first call allocates and skips, :660-661
                else:
                    for q in I:
                        f[(q,) + m1] = self.hw[face][int(OPPO[q])]
            elif code in (202, 204):
                if code == 204 and self.hw[face] is None:
                    self.hw[face] = np.zeros((19,) + shape2)
                    continue
                src = f[L1].copy() if code == 202 else self.hw[face]
                for q in I:
                    e1, e2, e3 = (float(t) for t in EE[q])
                    uxyz = v[0] * e1 + v[1] * e2 + v[2] * e3
                    f[(q,) + m1] = src[int(OPPO[q])] + 2.0 * WT[q] * self.flow.denIn * uxyz * 3.0
            elif code == 302:
                tmp = [f[(q,) + m1].copy() for q in Mi]
                for q, t in zip(I, tmp):
                    f[(q,) + m1] = t
            else:
                raise ValueError("has no such boundary condition")

    def ComputeFieldStat(self):
        """FluidDomain.f90:1739-1768."""
        inv = 1.0 / self.flow.Uref
        out = []
        n = float(self.X * self.Y * self.Z)
        for k in range(3):
            out.append(math.sqrt(float(np.sum((self.uuu[k] * inv) ** 2)) / n))
        for k in range(3):
            out.append(float(np.max(np.abs(self.uuu[k] * inv))))
        return np.array(out)

    def step(self, bodies=(), rootBC=None):
        """LBMBlockComm.f90:283-303 without sons."""
        self.update_volume_force()
        self.calculate_macro_quantities()
        self.ResetVolumeForce()
        it = calculate_interaction_force(self, bodies, self.bc if rootBC is None else rootBC) if len(bodies) else 0
        self.add_volume_force()
        self.collision()
        self.halfwayBCset()
        self.streaming()
        self.set_boundary_conditions()
        return it


# ---- immersed boundary (Solidbody.f90) ----------------------------------------------------------------------
class Body:
    def __init__(self, n, v_move=0, iBodyModel=1):
        self.n, self.v_move, self.iBodyModel, self.count_Interp = n, v_move, iBodyModel, 0
        self.v_Exyz, self.v_Evel = np.zeros((n, 3)), np.zeros((n, 3))
        self.v_Ea, self.v_Eforce = np.zeros(n), np.zeros((n, 3))
        self.v_Ei = np.zeros((n, 12), dtype=np.int16)
        self.v_Ew = np.zeros((n, 12), dtype=np.float32)


def Phi(x_):
    """Solidbody.f90:822-833."""
    r = abs(x_)
    if r < 1.0:
        return (3.0 - 2.0 * r + math.sqrt(1.0 + 4.0 * r * (1.0 - r))) * 0.125
    if r < 2.0:
        return (5.0 - 2.0 * r - math.sqrt(-7.0 + 4.0 * r * (3.0 - r))) * 0.125
    return 0.0


def trimedindex(i_, n, lo, hi):
    """Solidbody.f90:834-866, 1-based."""
    out = []
    for k in (-1, 0, 1, 2):
        v = i_ + k
        if v < 1:
            if lo == 301: v += n
            elif lo in (302, 201) and v == 0: v = 2
            elif lo == 203 and v == 0: v = 1
            else: raise IndexError("index out of xmin bound")
        elif v > n:
            if hi == 301: v -= n
            elif hi in (302, 201) and v == n + 1: v = n - 1
            elif hi == 203 and v == n + 1: v = n
            else: raise IndexError("index out of xmax bound")
        out.append(v)
    return out


def UpdateElmtInterp(body, blk, rootBC):
    """Solidbody.f90:760-806."""
    dh = blk.dh
    invdh = 1.0 / dh
    dims = (blk.X, blk.Y, blk.Z)
    anchors = []
    for a in range(3):
        i0 = math.floor((body.v_Exyz[0, a] - blk.mins[a]) * invdh)
        x0 = blk.mins[a] + float(i0) * dh
        anchors.append((x0, i0 + 1))
    for e in range(body.n):
        for a in range(3):
            x0, i0 = anchors[a]
            off = (body.v_Exyz[e, a] - x0) * invdh
            idx = math.floor(off)
            det = off - float(idx)
            idx += i0
            body.v_Ei[e, 4 * a:4 * a + 4] = trimedindex(idx, dims[a], rootBC[2 * a], rootBC[2 * a + 1])
            body.v_Ew[e, 4 * a:4 * a + 4] = [np.float32(Phi(float(m) - det)) for m in (-1, 0, 1, 2)]


def PenaltyForce(body, blk, dt):
    """Solidbody.f90:981-1049.  Returns (tolerance, ntolsum)."""
    dh = blk.dh
    invh3 = 0.5 * dt * ((1.0 / dh) * (1.0 / dh) * (1.0 / dh)) / blk.flow.denIn
    u = blk.uuu
    tol = 0.0
    felt = np.zeros((body.n, 3))
    for e in range(body.n):
        ix, jy, kz = (body.v_Ei[e, 0:4].astype(int) - 1, body.v_Ei[e, 4:8].astype(int) - 1, body.v_Ei[e, 8:12].astype(int) - 1)
        rx, ry, rz = (body.v_Ew[e, 0:4].astype(np.float64), body.v_Ew[e, 4:8].astype(np.float64), body.v_Ew[e, 8:12].astype(np.float64))
        vel = [0.0, 0.0, 0.0]
        for x in range(4):
            for y in range(4):
                for z in range(4):
                    for k in range(3):
                        vel[k] = vel[k] + u[k, ix[x], jy[y], kz[z]] * rx[x] * ry[y] * rz[z]
        d = [body.v_Evel[e, k] - vel[k] for k in range(3)]
        ft = [d[k] * body.v_Ea[e] for k in range(3)]
        tol = tol + abs(d[0]) + abs(d[1]) + abs(d[2])
        for k in range(3):
            body.v_Eforce[e, k] = body.v_Eforce[e, k] + ft[k]
            felt[e, k] = ft[k] * invh3
    if not math.isfinite(tol):
        raise FloatingPointError("Nan found in PenaltyForce")
    for e in range(body.n):
        ix, jy, kz = (body.v_Ei[e, 0:4].astype(int) - 1, body.v_Ei[e, 4:8].astype(int) - 1, body.v_Ei[e, 8:12].astype(int) - 1)
        rx, ry, rz = (body.v_Ew[e, 0:4].astype(np.float64), body.v_Ew[e, 4:8].astype(np.float64), body.v_Ew[e, 8:12].astype(np.float64))
        for x in range(4):
            for y in range(4):
                for z in range(4):
                    for k in range(3):
                        u[k, ix[x], jy[y], kz[z]] = u[k, ix[x], jy[y], kz[z]] - felt[e, k] * rx[x] * ry[y] * rz[z]
    return tol, float(body.n)


def FluidVolumeForce(body, blk):
    """Eulerian half of FluidVolumeForce_, Solidbody.f90:968-976."""
    dh = blk.dh
    invh3 = (1.0 / dh) * (1.0 / dh) * (1.0 / dh)
    F = blk.force
    for e in range(body.n):
        ix, jy, kz = (body.v_Ei[e, 0:4].astype(int) - 1, body.v_Ei[e, 4:8].astype(int) - 1, body.v_Ei[e, 8:12].astype(int) - 1)
        rx, ry, rz = (body.v_Ew[e, 0:4].astype(np.float64), body.v_Ew[e, 4:8].astype(np.float64), body.v_Ew[e, 8:12].astype(np.float64))
        fe = [body.v_Eforce[e, k] * invh3 for k in range(3)]
        for x in range(4):
            for y in range(4):
                for z in range(4):
                    for k in range(3):
                        F[k, ix[x], jy[y], kz[z]] = F[k, ix[x], jy[y], kz[z]] + (-fe[k] * rx[x] * ry[y] * rz[z])


def calculate_interaction_force(blk, bodies, rootBC, dt=None):
    """Solidbody.f90:869-918.  Returns iterLBM."""
    dt = blk.dh if dt is None else dt   # LBMBlockComm.f90:328 passes dh for dt
    for b in bodies:
        if b.v_move == 1 or b.iBodyModel == 2 or b.count_Interp == 0:
            UpdateElmtInterp(b, blk, rootBC)
            b.count_Interp = 1
        b.v_Eforce[...] = 0.0
    it = 0
    if len(bodies):
        dmax = 1e10
        while it < blk.flow.ntolLBM and dmax > blk.flow.dtolLBM:
            dmax, dsum = 0.0, 0.0
            for b in bodies:
                tol, ntol = PenaltyForce(b, blk, dt)
                dmax = dmax + tol
                dsum = dsum + ntol
            dmax = dmax / (dsum * blk.flow.Uref)
            it += 1
    for b in bodies:
        FluidVolumeForce(b, blk)
    return it


# ---- grid refinement: CommPair and father<->son transfers (LBMBlockComm.f90) ----------------------------------
def grid_transform(f, coeff, volumeForce, dh):
    """fIn_GridTransform, LBMBlockComm.f90:958-979, on arrays f[q, ...]."""
    den = 0.0 + f[0]
    for q in range(1, 19):
        den = den + f[q]
    u = []
    for k in range(3):
        m = 0.0 * f[0]
        for q in range(19):
            m = m + f[q] * float(EE[q, k])
        u.append((m + 0.5 * volumeForce[k] * dh) / den)
    uSqr = 0.0 + u[0] * u[0]
    uSqr = uSqr + u[1] * u[1]
    uSqr = uSqr + u[2] * u[2]
    out = np.empty_like(f)
    for q in range(19):
        uxyz = u[0] * float(EE[q, 0]) + u[1] * float(EE[q, 1]) + u[2] * float(EE[q, 2])
        fEq = WT[q] * den * ((1.0 - 1.5 * uSqr) + uxyz * (3.0 + 4.5 * uxyz))
        out[q] = fEq + coeff * (f[q] - fEq)
    return out


def interpolate_plane(fF, aS, bS, scheme):
    """interpolate_fIn (LBMBlockComm.f90:808-905) / interpolate_tau (:907-956) on arrays fF[..., aF, bF] -> [..., aS, bS]
    (b is the reference's first, faster index)."""
    lead = fF.shape[:-2]
    fS = np.zeros(lead + (aS, bS))
    bT = bS - 1 if bS % 2 == 0 else bS
    aT = aS - 1 if aS % 2 == 0 else aS
    nb, na = (bT + 1) // 2, (aT + 1) // 2
    C = fF[..., :na, :nb]
    fS[..., 0:aT:2, 0:bT:2] = C
    if scheme == 2:
        # even b (1-based) between coincident nodes, along b, rows with odd a only (:828-841)
        for i in range(nb - 1):           # i = b1-1 for b = 2*i+1 ; target 0-based column 2*i+1
            if i == 0:
                v = 0.375 * C[..., :, 0] + 0.75 * C[..., :, 1] - 0.125 * C[..., :, 2]
            elif 2 * i + 1 == bT - 2:
                v = 0.375 * C[..., :, i + 1] + 0.75 * C[..., :, i] - 0.125 * C[..., :, i - 1]
            else:
                v = -0.0625 * C[..., :, i - 1] + 0.5625 * C[..., :, i] + 0.5625 * C[..., :, i + 1] - 0.0625 * fF[..., :na, i + 2]
            fS[..., 0:aT:2, 2 * i + 1] = v
        for a in range(1, aT, 2):         # 0-based odd rows = 1-based even a (:844-856)
            a1 = a + 1                    # 1-based
            if a1 == 2:
                fS[..., a, :bT] = 0.375 * fS[..., a - 1, :bT] + 0.75 * fS[..., a + 1, :bT] - 0.125 * fS[..., a + 3, :bT]
            elif a1 == aT - 1:
                fS[..., a, :bT] = 0.375 * fS[..., a + 1, :bT] + 0.75 * fS[..., a - 1, :bT] - 0.125 * fS[..., a - 3, :bT]
            else:
                fS[..., a, :bT] = -0.0625 * fS[..., a - 3, :bT] + 0.5625 * fS[..., a - 1, :bT] + 0.5625 * fS[..., a + 1, :bT] - 0.0625 * fS[..., a + 3, :bT]
        if bS % 2 == 0:
            fS[..., :aT, bT] = -0.0625 * fS[..., :aT, bT - 3] + 0.5625 * fS[..., :aT, bT - 1] + 0.5625 * fS[..., :aT, 0] - 0.0625 * fS[..., :aT, 2]
        if aS % 2 == 0:
            fS[..., aT, :bT] = -0.0625 * fS[..., aT - 3, :bT] + 0.5625 * fS[..., aT - 1, :bT] + 0.5625 * fS[..., 0, :bT] - 0.0625 * fS[..., 2, :bT]
            if bS % 2 == 0:
                fS[..., aT, bT] = -0.0625 * fS[..., aT - 3, bT] + 0.5625 * fS[..., aT - 1, bT] + 0.5625 * fS[..., 0, bT] - 0.0625 * fS[..., 2, bT]
    else:
        fS[..., 0:aT:2, 1:bT:2] = (C[..., :, :-1] + C[..., :, 1:]) * 0.5
        fS[..., 1:aT:2, :bT] = (fS[..., 0:aT - 1:2, :bT] + fS[..., 2:aT:2, :bT]) * 0.5
        if bS % 2 == 0:
            fS[..., :aT, bT] = (fS[..., :aT, bT - 1] + fS[..., :aT, 0]) * 0.5
        if aS % 2 == 0:
            fS[..., aT, :bT] = (fS[..., aT - 1, :bT] + fS[..., 0, :bT]) * 0.5
            if bS % 2 == 0:
                fS[..., aT, bT] = (fS[..., aT - 1, bT] + fS[..., 0, bT]) * 0.5
    return fS


class Pair:
    """type CommPair (LBMBlockComm.f90:11-18), build_blocks_comunication (:32-96), and the transfers."""

    def __init__(self, father: Block, son: Block, interpolateScheme=1):
        self.F, self.S, self.scheme = father, son, interpolateScheme
        F, S = father, son
        sdims = [S.X, S.Y, S.Z]
        r = [0 if S.periodic[k] else 1 for k in range(3)]
        smax = [S.maxs[k] + (S.dh if S.periodic[k] else 0.0) for k in range(3)]
        bad = abs(F.dh - S.dh * 2.0) > 1e-8 or any(sdims[k] % 2 != r[k] for k in range(3))
        res1 = sum((S.mins[k] - F.mins[k]) / F.dh for k in range(3))
        res2 = sum((smax[k] - F.mins[k]) / F.dh for k in range(3))
        res = abs(res1 - round(res1)) + abs(res2 - round(res2))
        if bad or res > 1e-8:
            raise ValueError("grid points do not match between fluid blocks")
        self.sds = [(1 if j % 2 == 0 else -1) if S.bc[j] == 0 else 0 for j in range(6)]
        sD = [sdims[k] - (1 if S.periodic[k] else 0) for k in range(3)]
        ratio = math.floor(F.dh / S.dh + 0.5)
        self.s, self.f = [0] * 6, [0] * 6
        for k in range(3):
            self.s[2 * k], self.s[2 * k + 1] = 1, sD[k]
            self.f[2 * k] = math.floor((S.mins[k] - F.mins[k]) / F.dh + 1.5)
            self.f[2 * k + 1] = self.f[2 * k] + (sD[k] - 1) // ratio
        self.si = [self.s[j] + self.sds[j] * ratio for j in range(6)]
        self.fi = [self.f[j] + self.sds[j] for j in range(6)]
        self.dimS = sdims
        self.dimF = [self.f[2 * k + 1] - self.f[2 * k] + 1 for k in range(3)]
        self.buf = {}      # (face, 't1'|'t2') -> (f[19,a,b], tau[a,b])

    def _father_plane_index(self, j):
        """index tuple (x,y,z slices, 0-based) of the father's plane f(j) over the son's footprint"""
        axis = j // 2
        idx = []
        for k in range(3):
            if k == axis:
                idx.append(self.f[j] - 1)
            else:
                idx.append(slice(self.f[2 * k] - 1, self.f[2 * k] - 1 + self.dimF[k]))
        return tuple(idx)

    def extract_interpolate_layer(self, time):
        """LBMBlockComm.f90:340-505."""
        for j in range(6):
            if self.S.bc[j] != 0:
                continue
            idx = self._father_plane_index(j)
            f = self.F.f[(slice(None),) + idx].copy()
            tau = self.F.tau_all[idx].copy()
            self.buf[(j, time)] = (f, tau)
            if time == 2:
                f1, t1 = self.buf[(j, 1)]
                self.buf[(j, 1)] = (0.5 * (f1 + f), 0.5 * (t1 + tau))

    def interpolation_father_to_son(self, n_timeStep):
        """LBMBlockComm.f90:655-806."""
        S = self.S
        for j in range(6):
            if self.sds[j] == 0:
                continue
            axis = j // 2
            inplane = [k for k in range(3) if k != axis]        # [a-axis, b-axis] in x<y<z order = (slower, faster)
            aS, bS = self.dimS[inplane[0]], self.dimS[inplane[1]]
            fF, tauF = self.buf[(j, 1 if n_timeStep == 0 else 2)]
            tmpf = interpolate_plane(fF, aS, bS, self.scheme)
            tmptau = interpolate_plane(tauF, aS, bS, 1)
            idx = [slice(None)] * 3
            idx[axis] = self.s[j] - 1
            idx = tuple(idx)
            coeff = (S.tau_all[idx] / tmptau) / 2.0
            S.f[(slice(None),) + idx] = grid_transform(tmpf, coeff, self.F.volumeForce, self.F.dh)

    def deliver_son_to_father(self):
        """LBMBlockComm.f90:546-653."""
        S, F = self.S, self.F
        for j in range(6):
            if self.sds[j] == 0:
                continue
            axis = j // 2
            sidx, fidx = [None] * 3, [None] * 3
            for k in range(3):
                if k == axis:
                    sidx[k], fidx[k] = self.si[j] - 1, self.fi[j] - 1
                else:
                    lo, hi = self.si[2 * k], self.si[2 * k + 1]
                    n = (hi - lo) // 2 + 1
                    sidx[k] = slice(lo - 1, hi, 2)
                    fidx[k] = slice(self.fi[2 * k] - 1, self.fi[2 * k] - 1 + n)
            sidx, fidx = tuple(sidx), tuple(fidx)
            coeff = (F.tau_all[fidx] / S.tau_all[sidx]) * 2.0
            F.f[(slice(None),) + fidx] = grid_transform(S.f[(slice(None),) + sidx].copy(), coeff, S.volumeForce, S.dh)


class Node:
    def __init__(self, block, bodies=()):
        self.block, self.bodies, self.sons, self.comm = block, list(bodies), [], []

    def add_son(self, node, interpolateScheme=1):
        self.sons.append(node)
        self.comm.append(Pair(self.block, node.block, interpolateScheme))
        return node


def tree_step(node, rootBC=None, iters=None):
    """tree_collision_streaming_IBM_FEM, LBMBlockComm.f90:279-318 (FEM Solver excluded)."""
    b = node.block
    rootBC = b.bc if rootBC is None else rootBC
    b.update_volume_force()
    b.calculate_macro_quantities()
    b.ResetVolumeForce()
    it = calculate_interaction_force(b, node.bodies, rootBC) if node.bodies else 0
    if iters is not None:
        iters.append(it)
    b.add_volume_force()
    for p in node.comm:
        p.extract_interpolate_layer(1)
    b.collision()
    b.halfwayBCset()
    b.streaming()
    b.set_boundary_conditions()
    for p in node.comm:
        p.extract_interpolate_layer(2)
    for son, p in zip(node.sons, node.comm):
        for n in range(2):
            son.block.blktime = son.block.blktime + float(n) * son.block.dh
            tree_step(son, rootBC, iters)
            p.interpolation_father_to_son(n)
        p.deliver_son_to_father()


def set_blktime_all(node, time):
    node.block.blktime = time
    for s in node.sons:
        set_blktime_all(s, time)
