"""Generates tests/golden/*.npz from the independent numpy restatement (oracle/numpy_restatement.py).

    python tests/golden/make_golden.py

The reference itself cannot produce vectors here (Fortran, no compiler in the image: SURVEY F4), so the
golden vectors come from the second restatement written separately from the C oracle.  The C oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py) are both held to them.
Each file stores the case definition (so the tests rebuild the identical case), the seeded initial
populations and the state after N steps.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import numpy_restatement as NR   # noqa: E402
from tests.common import perturbed_state     # noqa: E402

CASES = {
    # name: dims, bc, model, params, dh, mins, flow, steps, plate
    "srt_periodic_force": dict(dims=(8, 6, 10), bc=(301,) * 6, model=1, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                               flow=dict(nu=0.1, volumeForceIn=(1e-6, 2e-7, -3e-7)), steps=12),
    "trt_halfway_channel": dict(dims=(6, 8, 10), bc=(301, 301, 203, 203, 301, 301), model=2, params=(3 / 16,) + (0.0,) * 9, dh=0.5,
                                mins=(-1.0, 0.5, 2.0), flow=dict(nu=0.05, volumeForceIn=(1e-6, 0.0, 0.0)), steps=12),
    "mrt_osc_force": dict(dims=(6, 6, 8), bc=(301,) * 6, model=3, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                          flow=dict(nu=0.08, volumeForceIn=(1e-6, 0.0, 0.0), volumeForceAmp=5e-7, volumeForceFreq=0.01, volumeForcePhi=30.0), steps=10),
    "srt_all_faces_mixed": dict(dims=(7, 8, 9), bc=(102, 104, 202, 204, 203, 201), model=1, params=(0.0,) * 10, dh=0.5, mins=(-1.0, 0.5, 2.0),
                                flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), shearRateIn=(0.0, 5e-4, 2e-4)), steps=12),
    "trt_inlet_outlet_symmetric": dict(dims=(8, 6, 7), bc=(101, 103, 302, 302, 301, 301), model=2, params=(0.25,) + (0.0,) * 9, dh=1.0,
                                       mins=(0.0, 0.0, 0.0), flow=dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0)), steps=12),
    "srt_oscillatory_inflow": dict(dims=(8, 5, 6), bc=(101, 104, 301, 301, 301, 301), model=1, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                                   flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), velocityKind=2, shearRateIn=(0.01, 0.02, 45.0)), steps=10),
    "les_smag_halfway_channel": dict(dims=(8, 8, 10), bc=(301, 301, 203, 203, 301, 301), model=11, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                                     flow=dict(nu=0.002, uvwIn=(0.05, 0.0, 0.0), volumeForceIn=(1e-6, 0.0, 0.0)), steps=10, wave_amp=2e-2),
    "les_wale_inflow": dict(dims=(9, 7, 8), bc=(101, 104, 201, 302, 301, 301), model=14, params=(0.0,) * 10, dh=0.5, mins=(0.0, 0.0, 0.0),
                            flow=dict(nu=0.002, uvwIn=(0.05, 0.0, 0.0)), steps=10, wave_amp=2e-2),
    "les_vrem_periodic": dict(dims=(7, 8, 9), bc=(301,) * 6, model=15, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                              flow=dict(nu=0.002, uvwIn=(0.05, 0.01, 0.0), volumeForceIn=(1e-6, 0.0, 0.0)), steps=10, wave_amp=2e-2),
    "srt_plate_shear": dict(dims=(16, 14, 12), bc=(101, 104, 202, 202, 301, 301), model=1, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                            flow=dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, 2e-4, 0.0), Uref=0.05, ntolLBM=3, dtolLBM=1e-30), steps=8,
                            plate=dict(origin=(5.3, 6.2, 3.4), nEL=4, len1=1.0, Nspan=5, spanlen=5.0, Lspan=0.0, chord_dir=(1.0, 0.3, 0.0))),
    "srt_plate_periodic_wrap": dict(dims=(12, 10, 10), bc=(301,) * 6, model=1, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0),
                                    flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, ntolLBM=4, dtolLBM=1e-30), steps=8,
                                    plate=dict(origin=(9.7, 4.2, 7.6), nEL=4, len1=1.0, Nspan=4, spanlen=4.0, Lspan=0.0, chord_dir=(1.0, 0.2, 0.0))),
}


def plate_markers(p, denIn, alpha=1.0):
    """Markers of a rigid flat plate exactly as PlateUpdatePosVelArea_ lays them out (Solidbody.f90:604-646)."""
    cd = np.asarray(p["chord_dir"], float); cd /= np.linalg.norm(cd)
    sd = np.array([0.0, 0.0, 1.0])
    nodes = np.asarray(p["origin"], float)[None, :] + np.arange(p["nEL"] + 1)[:, None] * p["len1"] * cd[None, :]
    beta = -alpha * 2.0 * denIn
    dl = p["spanlen"] / float(p["Nspan"])
    xyz, ea = [], []
    for i in range(p["nEL"]):
        c = 0.5 * (nodes[i] + nodes[i + 1])
        for s in range(1, p["Nspan"] + 1):
            ls = dl * (0.5 + float(s - 1)) - p["Lspan"]
            xyz.append(c + sd * ls)
            ea.append(dl * p["len1"] * beta)
    return np.array(xyz), np.zeros((len(xyz), 3)), np.array(ea)


def run_case(name, c):
    fl = NR.Flow(**c["flow"])
    X, Y, Z = c["dims"]
    b = NR.Block(X, Y, Z, dh=c["dh"], xmin=c["mins"][0], ymin=c["mins"][1], zmin=c["mins"][2], BndConds=c["bc"],
                 iCollidModel=c["model"], params=c["params"], flow=fl)
    b.initialise(0.0)
    f0 = perturbed_state(c["dims"], fl, wave_amp=c.get("wave_amp", 1e-3))
    b.f[...] = f0
    b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()   # main.f90:62-64
    bodies = []
    out = {}
    if "plate" in c:
        xyz, vel, ea = plate_markers(c["plate"], fl.denIn)
        body = NR.Body(len(ea))
        body.v_Exyz[...] = xyz; body.v_Evel[...] = vel; body.v_Ea[...] = ea
        bodies = [body]
        out.update(Exyz=xyz, Evel=vel, Ea=ea)
    its = []
    for n in range(1, c["steps"] + 1):
        b.blktime = c["dh"] * n
        its.append(b.step(bodies))
    b.calculate_macro_quantities()   # main.f90:107
    out.update(f0=f0, fIn=b.f, den=b.den, uuu=b.uuu, iters=np.array(its), case=json.dumps(c))
    if c["model"] >= 11:
        out.update(tau_all=b.tau_all)
    if bodies:
        out.update(Eforce=bodies[0].v_Eforce, Ei=bodies[0].v_Ei, Ew=bodies[0].v_Ew)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {X}x{Y}x{Z}, {c['steps']} steps, max|u| {np.abs(b.uuu).max():.3e}, mean rho {b.den.mean():.12f}")


REFINE_CASES = {
    "refine_linear": dict(scheme=1, fbc=(102, 104, 301, 301, 203, 203), fdims=(14, 10, 10), sbc=(0,) * 6, sdims=(13, 9, 9), smins=(4.0, 3.0, 3.0),
                          models=(1, 1), steps=6),
    "refine_cubic_periodic_son": dict(scheme=2, fbc=(101, 104, 301, 301, 301, 301), fdims=(14, 10, 10), sbc=(0, 0, 301, 301, 0, 0), sdims=(13, 20, 11),
                                      smins=(4.0, 0.0, 2.0), models=(2, 1), steps=6),
}
REFINE_FLOW = dict(nu=0.02, uvwIn=(0.03, 0.005, 0.0), Uref=0.03, volumeForceIn=(1e-6, 0.0, 0.0))
REFINE_PARAMS = (0.25,) + (0.0,) * 9


def run_refine_case(name, c):
    fl = NR.Flow(**REFINE_FLOW)
    Fb = NR.Block(*c["fdims"], dh=1.0, BndConds=c["fbc"], iCollidModel=c["models"][0], params=REFINE_PARAMS, flow=fl)
    Sb = NR.Block(*c["sdims"], dh=0.5, xmin=c["smins"][0], ymin=c["smins"][1], zmin=c["smins"][2], BndConds=c["sbc"],
                  iCollidModel=c["models"][1], params=REFINE_PARAMS, flow=fl)
    Fb.initialise(0.0); Sb.initialise(0.0)
    f0F, f0S = perturbed_state(c["fdims"], fl, seed=1), perturbed_state(c["sdims"], fl, seed=2)
    Fb.f[...] = f0F; Sb.f[...] = f0S
    root = NR.Node(Fb); root.add_son(NR.Node(Sb), c["scheme"])
    for b in (Fb, Sb):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for n in range(1, c["steps"] + 1):
        NR.set_blktime_all(root, float(n))
        NR.tree_step(root)
    p = root.comm[0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), f0F=f0F, f0S=f0S, fF=Fb.f, fS=Sb.f,
                        pair=np.array(p.sds + p.s + p.f + p.si + p.fi + p.dimS + p.dimF), case=json.dumps(c))
    print(f"{name}: father {c['fdims']} son {c['sdims']} scheme {c['scheme']}, {c['steps']} father steps")


if __name__ == "__main__":
    for name, c in CASES.items():
        run_case(name, c)
    for name, c in REFINE_CASES.items():
        run_refine_case(name, c)
