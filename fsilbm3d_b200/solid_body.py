"""Marker-level mirror of `type VirtualBody` (Solidbody.f90:25-68) and a rigid, prescribed-motion
plate that produces markers the way PlateUpdatePosVelArea_ (Solidbody.f90:604-646) does.

This is host-side code on both sides of the boundary: in the reference and in a drop-in build the
Fortran driver owns the beam state and calls PlateUpdatePosVelArea_ itself; the library only
receives v_Exyz/v_Evel/v_Ea and returns v_Eforce.  The plate here exists so tests and bench.py can
feed the oracle and the CUDA path identical markers without the structural FEM (out of scope,
SURVEY 8f1).  Only translation is prescribed (XYZ(t) of SolidSolver.f90:1831,1849); the plate does
not rotate.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


class VirtualBody:
    def __init__(self, nelmts: int, v_move: int = 0, iBodyModel: int = 1):
        self.v_nelmts = nelmts
        self.v_move = v_move
        self.iBodyModel = iBodyModel
        self.count_Interp = 0
        self.v_Exyz = np.zeros((nelmts, 3))      # v_Exyz(3,n)
        self.v_Evel = np.zeros((nelmts, 3))
        self.v_Ea = np.zeros(nelmts)
        self.v_Eforce = np.zeros((nelmts, 3))


class RigidPlate:
    """Flat plate: nEL beam elements along `chord_dir` from `origin`, each of length len1, with Nspan
    markers across `span_dir` (width spanlen, starting Lspan to the 'left' of the beam axis)."""

    def __init__(self, origin: Sequence[float], nEL: int, len1: float, Nspan: int, spanlen: float, Lspan: float,
                 chord_dir=(1.0, 0.0, 0.0), span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha: float = 1.0, denIn: float = 1.0,
                 XYZAmpl=(0.0, 0.0, 0.0), XYZPhi=(0.0, 0.0, 0.0), Freq: float = 0.0, initXYZVel=(0.0, 0.0, 0.0)):
        self.nEL, self.Nspan, self.len1, self.spanlen, self.Lspan = nEL, Nspan, len1, spanlen, Lspan
        cd = np.asarray(chord_dir, float); cd /= np.linalg.norm(cd)
        sd = np.asarray(span_dir, float); sd /= np.linalg.norm(sd)
        self.dirc = sd
        self.node_ref = np.asarray(origin, float)[None, :] + np.arange(nEL + 1)[:, None] * len1 * cd[None, :]
        self.XYZAmpl = np.asarray(XYZAmpl, float); self.XYZPhi = np.asarray(XYZPhi, float)
        self.Freq = Freq; self.initXYZVel = np.asarray(initXYZVel, float)
        self.alpha, self.denIn = IBPenaltyAlpha, denIn
        moving = (np.abs(self.initXYZVel).sum() > 1e-5) or (np.abs(self.XYZAmpl).sum() > 1e-5)   # Solidbody.f90:292-297
        self.body = VirtualBody(nEL * Nspan, v_move=1 if moving else 0, iBodyModel=1)
        self.pos = self.node_ref.copy()
        self.vel = np.zeros((nEL + 1, 6))
        self.structure(0.0, 1, 0.0, 0.0)   # Beam_Initialise at time 0 (SolidSolver.f90:1415,1432)
        self.PlateUpdatePosVelArea()

    def structure(self, time: float, isubstep: int, deltat: float, subdeltat: float):
        """Rigid branch of Beam_structure (SolidSolver.f90:1826-1857), translation only."""
        m_pi = 3.141592653589793
        t = time - deltat + float(isubstep) * subdeltat
        XYZ = np.array([self.XYZAmpl[k] * math.cos(2.0 * m_pi * self.Freq * t + self.XYZPhi[k]) + self.initXYZVel[k] * t for k in range(3)])
        UVW = np.array([-2.0 * m_pi * self.Freq * self.XYZAmpl[k] * math.sin(2.0 * m_pi * self.Freq * t + self.XYZPhi[k]) + self.initXYZVel[k] for k in range(3)])
        self.pos = self.node_ref + XYZ[None, :]
        self.vel[:, 0:3] = UVW[None, :]
        self.vel[:, 3:6] = 0.0

    def PlateUpdatePosVelArea(self):
        """Solidbody.f90:604-646."""
        b = self.body
        IBPenaltyBeta = -self.alpha * 2.0 * self.denIn          # :613
        dl = self.spanlen / float(self.Nspan)                    # :622
        area = dl * self.len1 * IBPenaltyBeta                    # :624
        cnt = 0
        for i in range(self.nEL):
            tmpxyz = 0.5 * (self.pos[i] + self.pos[i + 1])
            tmpvel = 0.5 * (self.vel[i, 0:3] + self.vel[i + 1, 0:3])
            omega = 0.5 * (self.vel[i, 3:6] + self.vel[i + 1, 3:6])
            for s in range(1, self.Nspan + 1):
                ls = dl * (0.5 + float(s - 1)) - self.Lspan      # :636
                rspan = self.dirc * ls
                wspin = np.array([omega[1] * rspan[2] - omega[2] * rspan[1],
                                  omega[2] * rspan[0] - omega[0] * rspan[2],
                                  omega[0] * rspan[1] - omega[1] * rspan[0]])
                b.v_Exyz[cnt] = tmpxyz + rspan
                b.v_Evel[cnt] = tmpvel + wspin
                b.v_Ea[cnt] = area
                cnt += 1

    def UpdatePosVelArea(self):
        """UpdatePosVelArea_, Solidbody.f90:729-738."""
        if self.body.v_move == 1 or self.body.iBodyModel == 2:
            self.PlateUpdatePosVelArea()
