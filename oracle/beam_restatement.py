"""TEST INFRASTRUCTURE -- an independent numpy restatement of the reference's structural solver
(/root/reference/src/SolidSolver.f90) and of the host half of its marker bookkeeping (Solidbody.f90:604-646,
:945-967), written from the Fortran separately from harness/beam_solver.cpp so the two can be cross-checked.
The reference ships no fixtures and no Fortran compiler exists here; its beam solver is run by oracle/ftn/ in the flexible-plate pin
case (tests/reference_cases.py);
the pins are the analytic known-answer tests in tests/test_beam_kat.py.

Only tests/ may import this module.  Style differs from the C++ on purpose: all elements are handled at once as
stacked (nEL,12,12) arrays, the tangent operator is assembled into one dense global matrix, triads are (nEL,3,3)
stacks with columns = axes.  The iteration structure (Newmark-beta, full Newton-Raphson, block-Jacobi CG with lifted
Dirichlet values and its 1e-6 absolute residual stop) follows the reference exactly, because the CG stop criterion
makes the answer depend on the iteration path at the 1e-6/stiffness level.
"""
from __future__ import annotations

import math

import numpy as np

M_PI = 3.141592653589793          # m_pi, SolidSolver.f90:1261
PI_TYPO = 3.141562653589793       # the private pi of AoAtoTTT, :2406


def AoAtoTTT(AoA):
    """:2404-2472.  TTT = Rx * Ry * Rz with snap-to-axis within 1e-5."""
    def cs(a):
        c, s = math.cos(a), math.sin(a)
        if abs(a) < 1e-5:
            c, s = 1.0, 0.0
        if abs(a - 0.5 * PI_TYPO) < 1e-5:
            c, s = 0.0, 1.0
        if abs(a + 0.5 * PI_TYPO) < 1e-5:
            c, s = 0.0, -1.0
        return c, s
    c, s = cs(AoA[0]); Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], float)
    c, s = cs(AoA[1]); Ry = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], float)
    c, s = cs(AoA[2]); Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], float)
    return Rx @ (Ry @ (Rz @ np.eye(3)))


def rotvec_between(T1, T2):
    """Segment_get_angle_triad, :909-952, for stacks (...,3,3): rotation vector of T2 * T1^T."""
    R = T2 @ np.swapaxes(T1, -1, -2)
    d = 0.5 * np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    tr = np.clip((R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1.0) / 2.0, -1.0, 1.0)
    sint = np.sqrt((d * d).sum(-1))
    theta = np.arccos(tr)
    small = (sint < 1e-10) | (theta < 1e-10)
    fac = np.where(small, 1.0, theta / np.where(small, 1.0, sint))
    return d * fac[..., None]


def finite_rot(t):
    """Segment_FiniteRot (Rodrigues), :1066-1124, for a stack (...,3)."""
    t = np.asarray(t, float)
    tt = np.sqrt((t * t).sum(-1))
    small = tt < 1e-10
    safe = np.where(small, 1.0, tt)
    c1 = np.where(small, 1.0, np.sin(tt) / safe)
    c2 = np.where(small, 0.5, (1.0 - np.cos(tt)) / (safe * safe))
    t1, t2, t3 = t[..., 0], t[..., 1], t[..., 2]
    z = np.zeros_like(t1)
    S = np.stack([np.stack([z, -t3, t2], -1), np.stack([t3, z, -t1], -1), np.stack([-t2, t1, z], -1)], -2)
    S2 = np.stack([np.stack([-t3 * t3 - t2 * t2, t2 * t1, t3 * t1], -1), np.stack([t1 * t2, -t3 * t3 - t1 * t1, t3 * t2], -1),
                   np.stack([t1 * t3, t2 * t3, -t2 * t2 - t1 * t1], -1)], -2)
    return np.eye(3) + S * c1[..., None, None] + S2 * c2[..., None, None]


def axis_dir_triad(ax, d):
    """Segment_BuildAxisDirTriad, :624-697, one element: columns ex (axis), ey (span direction, Gram-Schmidt), ez."""
    ex = np.array(ax, float)
    n = math.sqrt(ex @ ex)
    if n > 1e-14:
        ex = ex / n
    dr = np.array(d, float)
    n = math.sqrt(dr @ dr)
    if n > 1e-14:
        dr = dr / n
    ey = dr - (dr @ ex) * ex
    n = math.sqrt(ey @ ey)
    if n <= 1e-10:
        if abs(ex[2]) > 0.995:
            ey = np.array([0.0, 1.0, 0.0]); ez = np.array([-ex[2], 0.0, 0.0])
        else:
            dd = math.sqrt(ex[0] ** 2 + ex[1] ** 2)
            ey = np.array([-ex[1] / dd, ex[0] / dd, 0.0]); ez = np.array([-ex[0] * ex[2] / dd, -ex[1] * ex[2] / dd, dd])
    else:
        ey = ey / n
        ez = np.cross(ex, ey)
        n = math.sqrt(ez @ ez)
        if n > 1e-14:
            ez = ez / n
    return np.stack([ex, ey, ez], 1)


def local_stiffness(E, G, A, Jt, Iy, Iz, L):
    """Segment_FormStiffMatrix, :365-478, stacked over elements (arrays of length nEL) -> (nEL,12,12)."""
    n = len(L)
    k = np.zeros((n, 12, 12))
    ks = 5.0 / 6.0
    phy = 12.0 * E * Iz / (ks * G * A * L * L)
    phz = 12.0 * E * Iy / (ks * G * A * L * L)

    def put(i, j, v):
        k[:, i - 1, j - 1] = v
        k[:, j - 1, i - 1] = v
    ax = A * E / L
    put(1, 1, ax); put(7, 7, ax); put(1, 7, -ax)
    tor = G * Jt / L
    put(4, 4, tor); put(10, 10, tor); put(4, 10, -tor)
    for (I, ph, tr, rot, sg) in ((Iz, phy, (2, 8), (6, 12), 1.0), (Iy, phz, (3, 9), (5, 11), -1.0)):
        k1 = 12.0 * E * I / (L ** 3 * (1.0 + ph)); k2 = 6.0 * E * I / (L ** 2 * (1.0 + ph))
        k3 = (4.0 + ph) * E * I / (L * (1.0 + ph)); k4 = (2.0 - ph) * E * I / (L * (1.0 + ph))
        a, b = tr
        c, d = rot
        put(a, a, k1); put(b, b, k1); put(a, b, -k1)
        put(c, c, k3); put(d, d, k3); put(c, d, k4)
        put(a, c, sg * k2); put(a, d, sg * k2); put(c, b, -sg * k2); put(b, d, -sg * k2)
    return k


def local_geometric(s, E, G, A, Jt, Iy, Iz, L):
    """Segment_FormGeomMatrix, :480-600."""
    n = len(L)
    g = np.zeros((n, 12, 12))
    ks = 5.0 / 6.0
    phy = 12.0 * E * Iz / (ks * G * A * L * L)
    phz = 12.0 * E * Iy / (ks * G * A * L * L)

    def put(i, j, v):
        g[:, i - 1, j - 1] = v
        g[:, j - 1, i - 1] = v
    for (ph, tr, rot, sg) in ((phy, (2, 8), (6, 12), 1.0), (phz, (3, 9), (5, 11), -1.0)):
        den = (1.0 + ph) ** 2
        g1 = s / L * (6.0 / 5.0 + 2.0 * ph + ph * ph) / den
        g2 = s / L * (L / 10.0) / den
        g3 = s / L * (2.0 * L * L / 15.0 + ph * L * L / 6.0 + ph * ph * L * L / 12.0) / den
        g4 = s / L * (-L * L / 30.0 - ph * L * L / 6.0 - ph * ph * L * L / 12.0) / den
        a, b = tr
        c, d = rot
        put(a, a, g1); put(b, b, g1); put(a, b, -g1)
        put(c, c, g3); put(d, d, g3); put(c, d, g4)
        put(a, c, sg * g2); put(a, d, sg * g2); put(c, b, -sg * g2); put(b, d, -sg * g2)
    gt = s * Jt / (A * L)
    put(4, 4, gt); put(10, 10, gt); put(4, 10, -gt)
    return g


def rotate_blocks(K, T):
    """Segment_RKR, :709-748, with R = triad_ee^T (RotateMatrix :699): every 3x3 block B -> T B T^T."""
    n = K.shape[0]
    B = K.reshape(n, 4, 3, 4, 3).transpose(0, 1, 3, 2, 4)
    Tt = np.swapaxes(T, -1, -2)
    B = T[:, None, None] @ B @ Tt[:, None, None]
    return B.transpose(0, 1, 3, 2, 4).reshape(n, 12, 12)


class Beam:
    """type BeamSolver for a chain mesh (element n joins nodes n, n+1), material taken from `prop` (isKB = 2 in the
    reference's terms) or derived as Beam_calculate_angle_material does for isKB = 0."""

    def __init__(self, xyz, Lspan, Rspan, dirc, constraint, Nspan, *, iBodyModel=2, isMotionGiven=(1,) * 6, prop=None,
                 isKB=2, EmR=0.0, tcR=0.0, psR=0.3, denR=1.0, Lref=1.0, Uref=1.0, denIn=1.0, Freq=0.0, XYZo=(0, 0, 0), initXYZVel=(0, 0, 0),
                 XYZAmpl=(0, 0, 0), XYZPhi_deg=(0, 0, 0), AoAo_deg=(0, 0, 0), AoAAmpl_deg=(0, 0, 0), AoAPhi_deg=(0, 0, 0),
                 dampK=0.0, dampM=0.0, GeoGamma=1.0, NewmarkGamma=0.5, NewmarkBeta=0.25, dtolFEM=1e-10, ntolFEM=20, g=(0, 0, 0),
                 IBPenaltyAlpha=1.0):
        xyz = np.asarray(xyz, float)
        self.nND = len(xyz); self.nEL = self.nND - 1; self.gEQ = 6 * self.nND
        self.n0 = np.arange(self.nEL); self.n1 = self.n0 + 1
        self.l2g = np.concatenate([6 * self.n0[:, None] + np.arange(6), 6 * self.n1[:, None] + np.arange(6)], 1)   # :73-78
        Lspan = np.broadcast_to(np.asarray(Lspan, float), (self.nND,)); Rspan = np.broadcast_to(np.asarray(Rspan, float), (self.nND,))
        dirc = np.broadcast_to(np.asarray(dirc, float), (self.nND, 3))
        self.x00 = np.zeros((self.nEL, 12)); self.x00[:, 0:3] = xyz[self.n0]; self.x00[:, 6:9] = xyz[self.n1]        # :79-82
        self.Lspan = 0.5 * (Lspan[self.n0] + Lspan[self.n1])                                                       # :90
        self.spanlen = 0.5 * (Rspan[self.n0] + Rspan[self.n1]) + self.Lspan                                        # :91
        d = 0.5 * (dirc[self.n0] + dirc[self.n1])
        self.dirc00 = d / np.sqrt((d * d).sum(1))[:, None]                                                         # :92-95
        self.Nspan = np.broadcast_to(np.asarray(Nspan, int), (self.nEL,)).copy()
        con = np.asarray(constraint, int).copy()
        bc = np.concatenate([con[self.n0], con[self.n1]], 1)                                                       # :86-87
        img = np.asarray(isMotionGiven, int)
        for e in range(self.nEL):                                                                                  # Beam_adjustBC :1386
            if bc[e, 0] == 1:
                bc[e, 0:6] = img
            if bc[e, 6] == 1:
                bc[e, 6:12] = img
        self.fixed = np.zeros(self.gEQ, bool)
        for e in range(self.nEL):
            self.fixed[self.l2g[e][bc[e] > 0]] = True
        self.iBodyModel = iBodyModel
        self.Freq = Freq
        rad = lambda a: np.asarray(a, float) / 180.0 * M_PI                                                        # :1546-1549
        self.XYZo = np.asarray(XYZo, float); self.initXYZVel = np.asarray(initXYZVel, float); self.XYZAmpl = np.asarray(XYZAmpl, float)
        self.XYZPhi = rad(XYZPhi_deg); self.AoAo = rad(AoAo_deg); self.AoAAmpl = rad(AoAAmpl_deg); self.AoAPhi = rad(AoAPhi_deg)
        self.dampK, self.dampM, self.GeoGamma, self.NG, self.NB = dampK, dampM, GeoGamma, NewmarkGamma, NewmarkBeta
        self.dtolFEM, self.ntolFEM = dtolFEM, ntolFEM
        self.g = np.asarray(g, float)
        self.beta_pen = -IBPenaltyAlpha * 2.0 * denIn                                                              # Solidbody.f90:613
        if isKB == 0:                                                                                              # :1579-1595
            ln = self.spanlen
            E = np.full(self.nEL, EmR * denIn * Uref ** 2)
            th = tcR * Lref
            A = ln * th
            ratio = th / ln
            self.prop = np.stack([E, E / (2.0 * (1.0 + psR)), A, denR * ln * Lref * denIn / A, np.zeros(self.nEL),
                                  ln * th ** 3 / 3.0 * (1.0 - 0.63 * ratio + 0.052 * ratio ** 5), ln * th ** 3 / 12.0, th * ln ** 3 / 12.0], 1)
        else:
            self.prop = np.broadcast_to(np.asarray(prop, float), (self.nEL, 8)).copy()
        self.cg_iterations = 0
        self.Initialise(0.0)

    # -- kinematics of the prescribed motion -------------------------------------------------------------------
    def _prescribed(self, t):
        ph = 2.0 * M_PI * self.Freq * t
        XYZ = self.XYZo + self.XYZAmpl * np.cos(ph + self.XYZPhi) + self.initXYZVel * t
        AoA = self.AoAo + self.AoAAmpl * np.cos(ph + self.AoAPhi)
        return XYZ, AoA

    def _rigid_velocity(self, t, AoA, TTT):
        ph = 2.0 * M_PI * self.Freq * t
        UVW = -2.0 * M_PI * self.Freq * self.XYZAmpl * np.sin(ph + self.XYZPhi) + self.initXYZVel
        W1 = -2.0 * M_PI * self.Freq * self.AoAAmpl * np.sin(ph + self.AoAPhi)
        W2 = np.array([W1[0] * math.cos(AoA[1]) + W1[2],
                       W1[0] * math.sin(AoA[1]) * math.sin(AoA[2]) + W1[1] * math.cos(AoA[2]),
                       W1[0] * math.sin(AoA[1]) * math.cos(AoA[2]) - W1[1] * math.sin(AoA[2])])
        W3 = TTT @ W2
        rel = self.pos[:, 0:3] - self.XYZ
        self.vel[:, 0:3] = np.cross(W3, rel) + UVW                                                                 # :1478-1494
        self.vel[:, 3:6] = W3

    def _map(self, TTT, XYZ, AoAd):
        out = np.zeros((self.nEL, 12))
        out[:, 0:3] = self.x00[:, 0:3] @ TTT.T + XYZ
        out[:, 6:9] = self.x00[:, 6:9] @ TTT.T + XYZ
        out[:, 3:6] = AoAd; out[:, 9:12] = AoAd
        return out

    def _scatter_pos(self):
        self.pos[self.n0] = self.x1[:, 0:6]
        self.pos[self.n1] = self.x1[:, 6:12]

    def _axis(self, x):
        d = x[:, 6:9] - x[:, 0:3]
        ln = np.sqrt((d * d).sum(1))
        return d, ln

    def Initialise(self, time):
        """Beam_Initialise, :1401-1446."""
        self.XYZ, self.AoA = self._prescribed(time)
        self.TTT0 = AoAtoTTT(self.AoA)
        AoAd = rotvec_between(self.TTT0, self.TTT0.copy())
        self.x0 = self._map(self.TTT0, self.XYZ, AoAd)
        self.dirc0 = self.dirc00 @ self.TTT0.T
        self.x1 = self.x0.copy(); self.xnxt = self.x0.copy()
        self.d0, self.len0 = self._axis(self.x0)
        self.d1, self.len1 = self.d0.copy(), self.len0.copy()
        self.pos = np.zeros((self.nND, 6)); self._scatter_pos()
        self.dsp = np.zeros((self.nND, 6)); self.vel = np.zeros((self.nND, 6)); self.acc = np.zeros((self.nND, 6))
        self.lodFlow = np.zeros(self.gEQ); self.lodRepl = np.zeros(self.gEQ)
        if self.iBodyModel == 1:
            self._rigid_velocity(time, self.AoA, self.TTT0)
        self.T1 = np.stack([axis_dir_triad(self.d0[e] / self.len0[e], self.dirc0[e]) for e in range(self.nEL)])   # InitTriad_D :602
        self.T2 = self.T1.copy(); self.Te = self.T1.copy()
        A, rho, Iy, Iz = self.prop[:, 2], self.prop[:, 3], self.prop[:, 6], self.prop[:, 7]
        roal = rho * A * self.len0 / 2.0                                                                           # FormMassMatrix :318
        diag = np.stack([roal, roal, roal, roal * (Iy + Iz) / A, roal * Iy / A, roal * Iz / A], 1)
        Mloc = np.zeros((self.nEL, 12, 12))
        idx = np.arange(6)
        Mloc[:, idx, idx] = diag
        Mloc[:, idx + 6, idx + 6] = diag
        self.Mel = rotate_blocks(Mloc, self.Te)
        self.geoFRM = np.zeros(self.nEL)
        gv = np.zeros(12); gv[0:3] = self.g; gv[6:9] = self.g
        self.lodGrav = np.zeros(self.gEQ)
        np.add.at(self.lodGrav, self.l2g, self._mass_times(np.broadcast_to(gv, (self.nEL, 12))))
        self.mss = np.zeros((self.nND, 3))
        np.add.at(self.mss, self.n0, np.stack([self.Mel[:, j, j] for j in range(3)], 1))
        np.add.at(self.mss, self.n1, np.stack([self.Mel[:, j + 6, j + 6] for j in range(3)], 1))
        self.iterNR, self.dnorm = 0, 0.0

    def _mass_times(self, q):
        """Segment_MassMultiply, :195-234: diagonal translational blocks, full rotational 3x3 blocks."""
        out = np.zeros_like(q)
        for o in (0, 6):
            out[:, o:o + 3] = np.stack([self.Mel[:, o + j, o + j] for j in range(3)], 1) * q[:, o:o + 3]
            out[:, o + 3:o + 6] = np.einsum("eij,ej->ei", self.Mel[:, o + 3:o + 6, o + 3:o + 6], q[:, o + 3:o + 6])
        return out

    # -- one structural sub-step -------------------------------------------------------------------------------
    def structure(self, time, isubstep, deltat, subdeltat):
        """Beam_structure, :1820-1877."""
        t = time - deltat + float(isubstep) * subdeltat
        self.XYZ, self.AoA = self._prescribed(t)
        TTTn = AoAtoTTT(self.AoA)
        AoAd = rotvec_between(self.TTT0, TTTn)
        self.xnxt = self._map(TTTn, self.XYZ, AoAd)
        if self.iBodyModel == 1:
            self.x1 = self.xnxt.copy()
            dirc1 = self.dirc00 @ TTTn.T
            self._scatter_pos()
            self.d1, self.len1 = self._axis(self.x1)
            self.T1 = np.stack([axis_dir_triad(self.d1[e] / self.len1[e], dirc1[e]) for e in range(self.nEL)])
            self.T2 = self.T1.copy(); self.Te = self.T1.copy()
            self._rigid_velocity(t, self.AoA, TTTn)
            return
        dt, be, ga = subdeltat, self.NB, self.NG                                                                   # UpdateNewmarkCoeffs :1879
        c = np.zeros(8)
        c[2] = 1.0 / (be * dt); c[1] = ga * c[2]; c[0] = c[2] / dt; c[3] = 0.5 / be - 1.0; c[4] = ga / be - 1.0
        c[5] = dt * (0.5 * ga / be - 1.0); c[6] = dt * (1.0 - ga); c[7] = dt * ga
        self._newton(c)

    def _end_rotations(self):
        """the local deformation vector ub of BodyStress_D / StrainEnergy_D, :769-808."""
        du = self.d1 - self.d0
        dl = ((self.d0 + self.d1) * du).sum(1) / (self.len0 + self.len1)
        ub = np.zeros((self.nEL, 12))
        for T, o in ((self.T1, 3), (self.T2, 9)):
            th = rotvec_between(self.Te, T)
            ub[:, o:o + 3] = np.einsum("eji,ej->ei", self.Te, th)                                                  # Segment_global_to_local :954
        ub[:, 6] = dl
        return ub

    def _newton(self, c):
        """Beam_Solver, :1894-1914."""
        vBC = np.zeros(self.gEQ)
        vBC[self.l2g] = self.xnxt - self.x1                                                                        # :1921-1925 (later elements overwrite)
        lodExte = self.lodFlow + self.lodGrav + self.lodRepl
        dspO, velO, accO = self.dsp.copy(), self.vel.copy(), self.acc.copy()
        E, G, A, Jt, Iy, Iz = (self.prop[:, i] for i in (0, 1, 2, 5, 6, 7))
        it = 0
        for it in range(1, self.ntolFEM + 1):
            Kloc = local_stiffness(E, G, A, Jt, Iy, Iz, self.len0)
            ub = self._end_rotations()
            self.geoFRM = ub[:, 6] * E * A / self.len0                                                             # :813-816
            fb = np.einsum("eij,ej->ei", Kloc, ub)
            fg = np.einsum("eij,eaj->eai", self.Te, fb.reshape(self.nEL, 4, 3)).reshape(self.nEL, 12)              # :829-836
            lodInte = np.zeros(self.gEQ)
            np.add.at(lodInte, self.l2g, fg)
            Kg = rotate_blocks(local_geometric(self.geoFRM, E, G, A, Jt, Iy, Iz, self.len0), self.Te)
            Ke = rotate_blocks(Kloc, self.Te)
            coef = Ke + self.GeoGamma * Kg + c[0] * self.Mel + c[1] * self.dampM * self.Mel                        # :151-152
            if self.dampK > 0.0:
                coef = coef + c[1] * self.dampK * Ke
            lodEffe = lodExte - lodInte
            dd = (dspO - self.dsp); v = self.vel; a = self.acc                                                     # UpdateLoad :159-193
            qM = c[0] * dd + c[2] * v + c[3] * a
            qC = c[1] * dd + c[4] * v + c[5] * a
            ele = lambda f: np.concatenate([f[self.n0], f[self.n1]], 1)
            np.add.at(lodEffe, self.l2g, self._mass_times(ele(qM + self.dampM * qC)))
            if self.dampK > 0.0:
                np.add.at(lodEffe, self.l2g, self.dampK * np.einsum("eij,ej->ei", Ke, ele(qC)))
            dspn = self._cg(coef, lodEffe, vBC if it == 1 else np.zeros(self.gEQ))
            self._update(dspn)
            self.dnorm = float(np.max(dspn * dspn))                                                                # :2336 (beta = 1)
            if self.dnorm <= self.dtolFEM:
                break
        else:
            it = self.ntolFEM + 1
        self.iterNR = it
        self.acc = c[0] * (self.dsp - dspO) - c[2] * velO - c[3] * accO                                            # :2356-2357
        self.vel = velO + c[6] * accO + c[7] * self.acc

    def _cg(self, coef, b, lift):
        """Beam_CG_Solve, :1949-2020, with the node-wise 6x6 block-Jacobi preconditioner (:2174-2232)."""
        n = self.gEQ
        Aglob = np.zeros((n, n))
        np.add.at(Aglob, (self.l2g[:, :, None], self.l2g[:, None, :]), coef)
        fx = self.fixed
        xF = np.where(fx, lift, 0.0)
        x = xF.copy()
        r = b - Aglob @ x
        r[fx] = 0.0
        Binv = np.zeros((self.nND, 6, 6))
        for nd in range(self.nND):
            B = Aglob[6 * nd:6 * nd + 6, 6 * nd:6 * nd + 6].copy()
            f6 = fx[6 * nd:6 * nd + 6]
            B[f6, :] = 0.0; B[:, f6] = 0.0
            B[f6, f6] = 1.0
            Binv[nd] = np.linalg.inv(B)

        def prec(rr):
            z = np.einsum("nij,nj->ni", Binv, rr.reshape(self.nND, 6)).reshape(n)
            z[fx] = 0.0
            return z
        z = prec(r)
        rn = math.sqrt(r @ r)
        if rn <= 1e-6:
            return x
        p = z.copy()
        rs = r @ z
        for _ in range(10000):
            self.cg_iterations += 1
            Ap = Aglob @ p
            Ap[fx] = 0.0
            alpha = rs / (p @ Ap)
            x = x + alpha * p
            x[fx] = xF[fx]
            r = r - alpha * Ap
            r[fx] = 0.0
            rn = math.sqrt(r @ r)
            if rn <= 1e-6:
                break
            z = prec(r)
            rsn = r @ z
            p = z + (rsn / rs) * p
            p[fx] = 0.0
            rs = rsn
        return x

    def _update(self, dspn):
        """Beam_UpdateDspANDTride, :2304-2338."""
        inc = dspn.reshape(self.nND, 6)
        self.dsp = self.dsp + inc
        self.x1[:, 0:6] = self.x0[:, 0:6] + self.dsp[self.n0]
        self.x1[:, 6:12] = self.x0[:, 6:12] + self.dsp[self.n1]
        self._scatter_pos()
        self.d1, self.len1 = self._axis(self.x1)
        self.T1 = finite_rot(inc[self.n0, 3:6]) @ self.T1                                                          # UpdateTriad_D :966
        self.T2 = finite_rot(inc[self.n1, 3:6]) @ self.T2
        e1 = self.d1 / self.len1[:, None]                                                                          # MakeTriad_ee :999
        half = 0.5 * rotvec_between(self.T1, self.T2)
        Ta = finite_rot(half) @ self.T1
        r2 = (Ta[:, :, 1] * e1).sum(1); r3 = (Ta[:, :, 2] * e1).sum(1)
        e2 = Ta[:, :, 1] - r2[:, None] * (Ta[:, :, 0] + e1) / 2.0
        e2 = e2 - (e2 * e1).sum(1)[:, None] * e1
        nn = np.sqrt((e2 * e2).sum(1))
        e2 = np.where((nn > 1e-14)[:, None], e2 / np.where(nn > 1e-14, nn, 1.0)[:, None], e2)
        e3 = np.cross(e1, e2)
        self.Te = np.stack([e1, e2, e3], 2)

    def strain_energy(self):
        """Segment_StrainEnergy_D, :841-907 -> (stretch, bend+torsion) per element."""
        E, G, A, Jt, Iy, Iz = (self.prop[:, i] for i in (0, 1, 2, 5, 6, 7))
        K = local_stiffness(E, G, A, Jt, Iy, Iz, self.len0)
        ub = self._end_rotations()
        tot = 0.5 * np.einsum("ei,eij,ej->e", ub, K, ub)
        st = 0.5 * K[:, 6, 6] * ub[:, 6] ** 2
        return st, tot - st

    # -- marker bookkeeping (host half of Solidbody.f90) -------------------------------------------------------
    def markers(self):
        """PlateUpdatePosVelArea_, Solidbody.f90:604-646 -> v_Exyz (n,3), v_Evel (n,3), v_Ea (n)."""
        X, V, Aa = [], [], []
        for e in range(self.nEL):
            c = 0.5 * (self.pos[e, 0:3] + self.pos[e + 1, 0:3])
            u = 0.5 * (self.vel[e, 0:3] + self.vel[e + 1, 0:3])
            om = 0.5 * (self.vel[e, 3:6] + self.vel[e + 1, 3:6])
            dl = self.spanlen[e] / float(self.Nspan[e])
            d = 0.5 * (self.T1[e][:, 1] + self.T2[e][:, 1])
            nn = math.sqrt(d @ d)
            d = d / nn if nn > 1e-12 else self.Te[e][:, 1]
            for s in range(1, self.Nspan[e] + 1):
                r = d * (dl * (0.5 + float(s - 1)) - self.Lspan[e])
                X.append(c + r); V.append(u + np.cross(om, r)); Aa.append(dl * self.len1[e] * self.beta_pen)
        return np.array(X), np.array(V), np.array(Aa)

    def fluid_loads(self, Exyz, Eforce):
        """lodFlow = 0 (Solidbody.f90:911) + the nodal-load half of FluidVolumeForce_ (:945-967)."""
        self.lodFlow = np.zeros(self.gEQ)
        m = 0
        for e in range(self.nEL):
            xc = 0.5 * (self.x1[e, 0:3] + self.x1[e, 6:9])
            for _ in range(self.Nspan[e]):
                F = Eforce[m]
                Mo = np.cross(Exyz[m] - xc, F)
                self.lodFlow[self.l2g[e][0:3]] += 0.5 * F; self.lodFlow[self.l2g[e][6:9]] += 0.5 * F
                self.lodFlow[self.l2g[e][3:6]] += 0.5 * Mo; self.lodFlow[self.l2g[e][9:12]] += 0.5 * Mo
                m += 1
