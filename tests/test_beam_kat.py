"""Analytic known-answer tests that pin BOTH restatements of the reference's structural solver (SolidSolver.f90):
harness/beam_solver.cpp (the stand-in driver's) and oracle/beam_restatement.py (the independent numpy one).
The reference ships no structural fixtures and cannot be built here, so closed-form beam theory is the anchor:
Timoshenko cantilever deflection in both bending planes, St. Venant twist, axial stretch, the first bending
frequency, the elastica roll-up under an end moment (finite rotations + nodal triads), and rigid kinematics."""
import math

import numpy as np
import pytest

from tests.beam_cases import chain, open_cpp, open_numpy

E, Gm, A, Iy, Iz, Jt, RHO = 1.0e4, 1.0e4 / 2.6, 0.01, 8.333e-6, 2.0e-5, 1.0e-5, 1.0
MAT = (E, Gm, A, RHO, 0.0, Jt, Iy, Iz)


class CppBeam:
    def __init__(self, tmp_path, n, **inflow):
        self.sb = open_cpp(str(tmp_path), chain(n), material=MAT, **inflow)
        self.b = self.sb.VBodies[0]
        self.nND = n

    def load(self, lod): self.b.set_lodFlow(lod)
    def step(self, t, dt): self.b.structure(t, 1, dt, dt)
    pos = property(lambda s: s.b.pos)
    dsp = property(lambda s: s.b.dsp)


class NumpyBeam:
    def __init__(self, tmp_path, n, **inflow):
        self.b = open_numpy(chain(n), material=MAT, **inflow)
        self.nND = n

    def load(self, lod): self.b.lodFlow = np.asarray(lod, float).reshape(-1).copy()
    def step(self, t, dt): self.b.structure(t, 1, dt, dt)
    pos = property(lambda s: s.b.pos)
    dsp = property(lambda s: s.b.dsp)


BACKENDS = [pytest.param(CppBeam, id="cpp"), pytest.param(NumpyBeam, id="numpy")]


def relax(beam, lod, steps=600, dt=0.05):
    for k in range(1, steps + 1):
        beam.load(lod)
        beam.step(k * dt, dt)


@pytest.mark.parametrize("backend", BACKENDS)
def test_cantilever_statics(backend, tmp_path):
    """Tip loads on a clamped beam, critically mass-damped until rest: delta = PL^3/(3EI) + PL/(ks G A) in both
    planes (SolidSolver.f90:409-422 uses Iz with v, Iy with w; ks = 5/6), twist TL/(G Jt), stretch PL/(EA)."""
    n = 11
    beam = backend(tmp_path, n, dampM=20.0, dtolFEM=1e-18, ntolFEM=20)
    P = 1.0e-3
    ks = 5.0 / 6.0
    # one load at a time: the solver is geometrically nonlinear, so combined loads couple (twist rotates the section,
    # tension stiffens bending)
    lod = np.zeros((n, 6)); lod[-1, 2] = P
    relax(beam, lod)
    assert beam.dsp[-1, 2] == pytest.approx(P / (3 * E * Iy) + P / (ks * Gm * A), rel=2e-4)
    lod[...] = 0.0; lod[-1, 1] = -P
    relax(beam, lod)
    assert beam.dsp[-1, 1] == pytest.approx(-(P / (3 * E * Iz) + P / (ks * Gm * A)), rel=2e-4)
    lod[...] = 0.0; lod[-1, 3] = 1.0e-3
    relax(beam, lod)
    assert beam.dsp[-1, 3] == pytest.approx(1.0e-3 / (Gm * Jt), rel=5e-4)
    lod[...] = 0.0; lod[-1, 0] = 0.1
    relax(beam, lod)
    assert beam.dsp[-1, 0] == pytest.approx(0.1 / (E * A), rel=1e-3)


@pytest.mark.parametrize("backend", BACKENDS)
def test_elastica_quarter_circle(backend, tmp_path):
    """A pure end moment M = kappa*E*Iy bends the cantilever into a circular arc of curvature kappa: tip at
    (sin(kL)/k, (1-cos(kL))/k).  kL = pi/2, 20 elements: the finite-rotation triad update (:966-1064) must hold."""
    n = 21
    beam = backend(tmp_path, n, dampM=20.0, dtolFEM=1e-20, ntolFEM=30)
    kap = math.pi / 2
    M = kap * E * Iy
    dt = 0.05
    for k in range(1, 1001):
        lod = np.zeros((n, 6)); lod[-1, 4] = -M * min(1.0, k / 200.0)
        beam.load(lod)
        beam.step(k * dt, dt)
    tip = beam.pos[-1]
    assert tip[0] == pytest.approx(math.sin(kap) / kap, rel=1e-3)
    assert tip[2] == pytest.approx((1 - math.cos(kap)) / kap, rel=1e-3)
    assert abs(tip[1]) < 1e-12
    assert beam.dsp[-1, 4] == pytest.approx(-kap, rel=1e-3)


@pytest.mark.parametrize("backend", BACKENDS)
def test_first_bending_frequency(backend, tmp_path):
    """Undamped free vibration after a short tip pulse (Newmark gamma = 1/2, beta = 1/4): the dominant frequency of
    the tip is the first cantilever mode, omega_1 = 1.8751^2 sqrt(E Iy / (rho A L^4)), to the accuracy of a
    20-element lumped-mass model."""
    n = 21
    beam = backend(tmp_path, n, dampM=0.0, dtolFEM=1e-22, ntolFEM=20)
    dt, steps = 0.004, 1500
    z = np.zeros(steps)
    for k in range(1, steps + 1):
        lod = np.zeros((n, 6))
        if k <= 40:
            lod[-1, 2] = 1.0e-4 * math.sin(math.pi * k / 40.0)
        beam.load(lod)
        beam.step(k * dt, dt)
        z[k - 1] = beam.pos[-1, 2]
    w1 = 1.8751 ** 2 * math.sqrt(E * Iy / (RHO * A))
    t = dt * np.arange(1, steps + 1)
    sel = t > 40 * dt
    zz = z[sel] - z[sel].mean()
    ws = w1 * np.linspace(0.85, 1.15, 601)
    power = [abs(np.sum(zz * np.exp(-1j * w * t[sel]))) for w in ws]
    assert ws[int(np.argmax(power))] == pytest.approx(w1, rel=0.03)


@pytest.mark.parametrize("backend", BACKENDS)
def test_rigid_prescribed_motion(backend, tmp_path):
    """iBodyModel = 1: nodes follow x = TTT(AoA(t)) x00 + XYZ(t), v = UVW + WWW3 x r (:1826-1857) with
    TTT = Rx Ry Rz (:2404-2472); checked for a pitch about y plus heave in z at t = 0.37."""
    n = 6
    grp = dict(iBodyModel=1, freq=0.8, XYZAmpl=(0.0, 0.0, 0.1), XYZPhi=(0.0, 0.0, 30.0), AoAo=(0.0, 5.0, 0.0), AoAAmpl=(0.0, 20.0, 0.0),
               AoAPhi=(0.0, 90.0, 0.0), firstXYZ=(0.3, 0.2, 0.1), initXYZVel=(0.01, 0.0, 0.0))
    beam = backend(tmp_path, n, group=grp)
    t = 0.37
    beam.b.structure(t, 1, 0.0, 0.0)
    w = 2 * math.pi * 0.8
    XYZ = np.array([0.3 + 0.01 * t, 0.2, 0.1 + 0.1 * math.cos(w * t + math.radians(30.0))])
    th = math.radians(5.0) + math.radians(20.0) * math.cos(w * t + math.radians(90.0))
    Ry = np.array([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
    x00 = chain(n)
    want = x00 @ Ry.T + XYZ
    assert np.allclose(beam.pos[:, 0:3], want, rtol=0, atol=1e-14)
    UVW = np.array([0.01, 0.0, -w * 0.1 * math.sin(w * t + math.radians(30.0))])
    om = np.array([0.0, -w * math.radians(20.0) * math.sin(w * t + math.radians(90.0)), 0.0])
    vel = beam.b.vel
    assert np.allclose(vel[:, 3:6], om[None, :], rtol=0, atol=1e-14)
    assert np.allclose(vel[:, 0:3], UVW[None, :] + np.cross(om, want - XYZ), rtol=0, atol=1e-14)
