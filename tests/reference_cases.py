"""Cases that pin the CPU oracle (and through it the CUDA path) against the REFERENCE ITSELF.

Each case is one description that yields
  (i)   the reference's own input files (inFlow.dat, plate.dat, an injected ./DatContinue/continue state) -- what
        tests/golden/make_reference_golden.py feeds to the reference's unmodified main program, executed from its Fortran
        sources by the interpreter in oracle/ftn/ (this image has no Fortran compiler); the populations, marker forces and
        beam state that run leaves are committed as tests/golden/ref_<case>.npz;
  (ii)  the same run on the C oracle (run_oracle below; tests/test_reference_golden.py, CPU);
  (iii) the same input directory for the C++ stand-in driver over the CUDA library (tests/test_gpu_reference_golden.py).

The injected state goes through the reference's own restart path (check_is_continue, FluidDomain.f90:128-237, isConCmpt=2:
populations taken from the file, time and step start at zero, main.f90:70-75).
"""
import json
import os
import struct

import numpy as np

from tests.common import SEED, perturbed_state

DIRS = ("DatFlow", "DatContinue", "DatInfo", "DatBody", "DatBodySpan", "DatTemp", "DatOthe")
P0 = (0.0,) * 10


def _fluid(dims, bc, model=1, params=P0, steps=8, uvwIn=(0.0, 0.0, 0.0), Uref=0.05, Lref=4.0, Re=20.0, wave=1e-3, **flow):
    return dict(kind="fluid", dims=dims, bc=bc, model=model, params=params, steps=steps, uvwIn=uvwIn, Uref=Uref, Lref=Lref, Re=Re, wave=wave,
                flow=flow)


CASES = {
    # collision models on a periodic box with a body force (FluidDomain.f90:1208-1263) from a perturbed state
    "srt_periodic_force": _fluid((10, 8, 12), (301,) * 6, model=1, steps=10, volumeForceIn=(1e-6, 2e-7, -3e-7)),
    "trt_periodic_force": _fluid((10, 8, 12), (301,) * 6, model=2, params=(3.0 / 16.0,) + (0.0,) * 9, steps=10, volumeForceIn=(1e-6, 2e-7, -3e-7)),
    "mrt_osc_force": _fluid((8, 10, 8), (301,) * 6, model=3, steps=8, volumeForceIn=(1e-6, 0.0, 0.0), volumeForceAmp=5e-7, volumeForceFreq=0.7,
                            volumeForcePhi=30.0),
    # every boundary rule (set_boundary_conditions_, FluidDomain.f90:616-1190), face-order precedence on shared edges
    "srt_all_faces_mixed": _fluid((10, 9, 8), (102, 104, 202, 204, 203, 201), steps=10, uvwIn=(0.03, 0.0, 0.0), Uref=0.03,
                                  shearRateIn=(0.0, 5e-4, 2e-4)),
    "trt_inlet_outlet_symmetric": _fluid((10, 8, 8), (101, 103, 302, 302, 301, 301), model=2, params=(0.25,) + (0.0,) * 9, steps=10,
                                         uvwIn=(0.04, 0.0, 0.0), Uref=0.04),
    "srt_oscillatory_inflow": _fluid((10, 8, 8), (101, 104, 301, 301, 301, 301), steps=10, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, velocityKind=2,
                                     shearRateIn=(0.01, 0.6, 45.0)),
    # LES closures (FluidDomain.f90:1264-1507)
    "les_smag_halfway_channel": _fluid((8, 9, 8), (301, 301, 203, 203, 301, 301), model=11, steps=8, volumeForceIn=(2e-6, 0.0, 0.0), wave=2e-2),
    "les_wale_inflow": _fluid((9, 8, 8), (101, 104, 301, 301, 301, 301), model=14, steps=8, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, wave=2e-2),
    "les_vrem_periodic": _fluid((8, 8, 9), (301,) * 6, model=15, steps=8, volumeForceIn=(1e-6, 0.0, 0.0), wave=2e-2),
}

# father 12x10x10 with a 2:1 refined son (LBMBlockComm.f90:279-505), linear and cubic interpolation in time/space
for _name, _scheme in (("refine_linear", 1), ("refine_cubic", 2)):
    CASES[_name] = dict(kind="refine", dims=(14, 10, 10), bc=(101, 104, 301, 301, 301, 301), sdims=(11, 9, 9), smins=(4.0, 3.0, 3.0), scheme=_scheme,
                        model=1, params=P0, steps=6, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, Lref=4.0, Re=20.0, wave=1e-3,
                        flow=dict(volumeForceIn=(1e-6, 0.0, 0.0)))

# a son that is periodic in y (even number of nodes, closures of interpolate_fIn, LBMBlockComm.f90:857-871) under a TRT father, cubic;
# Smagorinsky father and son (tau_all fields interpolated and rescaled, interpolate_tau :907-956)
CASES["refine_cubic_periodic_son_trt"] = dict(
    kind="refine", dims=(14, 10, 10), bc=(101, 104, 301, 301, 301, 301), sdims=(13, 20, 11), smins=(4.0, 0.0, 2.0), sbc=(0, 0, 301, 301, 0, 0),
    scheme=2, model=2, params=(3.0 / 16.0,) + (0.0,) * 9, smodel=1, sparams=P0, steps=6, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, Lref=4.0, Re=20.0,
    wave=1e-3, flow=dict(volumeForceIn=(1e-6, 0.0, 0.0)))
CASES["refine_les_smag"] = dict(CASES["refine_linear"], model=11, wave=2e-2, steps=5)
# MRT father and MRT son: the son's relaxation matrices are built from ITS tau (dh / 2), FluidDomain.f90:466-522, and fIn_GridTransform
# rescales the non-equilibrium part between the two (LBMBlockComm.f90:958-979).  Oracle-against-reference only (gpu=False): the CUDA side of
# several MRT blocks with different matrices is held against the oracle in tests/test_gpu_parity.py
CASES["refine_mrt_father_and_son"] = dict(CASES["refine_linear"], model=3, steps=5, gpu=False)
# a son periodic in y AND z: only its two x faces are coupled, and they carry the closures of interpolate_fIn in both face directions plus
# the corner closure (LBMBlockComm.f90:857-871 cubic, :892-905 linear); one case per scheme
for _name, _scheme in (("refine_linear_periodic_son_yz", 1), ("refine_cubic_periodic_son_yz", 2)):
    CASES[_name] = dict(kind="refine", dims=(14, 10, 10), bc=(101, 104, 301, 301, 301, 301), sdims=(13, 20, 20), smins=(4.0, 0.0, 0.0),
                        sbc=(0, 0, 301, 301, 301, 301), scheme=_scheme, model=1, params=P0, steps=5, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, Lref=4.0,
                        Re=20.0, wave=1e-3, flow=dict(volumeForceIn=(1e-6, 0.0, 0.0)))

# a rigid plate in prescribed heave (iBodyModel 1; Solidbody.f90:760-1049 IBM, SolidSolver.f90:1826-1857 motion)
CASES["rigid_plate_heave"] = dict(
    kind="body", dims=(20, 14, 14), bc=(101, 104, 202, 202, 301, 301), model=1, params=P0, steps=8, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=40.0, wave=1e-3, flow=dict(shearRateIn=(0.0, 2e-4, 0.0)), ntolLBM=3, dtolLBM=1e-30, numsubstep=1,
    plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, denR=1.0, psR=0.3, EmR=1.0, tcR=0.05, freq=0.02, XYZAmpl=(0.0, 1.0, 0.0), XYZPhi=(0.0, 20.0, 0.0),
               AoAo=(0.0, 0.0, 10.0), firstXYZ=(6.3, 6.6, 5.2)), isKB=0)
# a flexible plate, leading edge held, passively deforming (iBodyModel 2: the beam FEM, SolidSolver.f90:1860-2020)
CASES["flexible_plate"] = dict(
    kind="body", dims=(20, 14, 14), bc=(101, 104, 301, 301, 301, 301), model=1, params=P0, steps=6, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=40.0, wave=1e-3, flow={}, ntolLBM=3, dtolLBM=1e-30, numsubstep=2,
    plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=1.0, psR=0.3, KB=0.02, KS=500.0, AoAo=(0.0, 0.0, 12.0), firstXYZ=(6.3, 6.6, 5.2)), isKB=1)


# two rigid plates whose stencil boxes overlap (Gauss-Seidel order over the bodies, Solidbody.f90:898-903) and a finite tolerance, so
# that the loop control (:895-906) ends the penalty iteration early on some steps
CASES["two_plates_gauss_seidel"] = dict(
    kind="body", dims=(22, 16, 14), bc=(101, 104, 202, 202, 301, 301), model=1, params=P0, steps=8, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=40.0, wave=1e-3, flow=dict(shearRateIn=(0.0, 2e-4, 0.0)), ntolLBM=6, dtolLBM=0.5, numsubstep=1,
    plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    groups=[dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, AoAo=(0.0, 0.0, 8.0), firstXYZ=(6.3, 6.6, 5.2)),
            dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, freq=0.03, XYZAmpl=(0.0, 0.8, 0.0), AoAo=(0.0, 0.0, -6.0),
                 firstXYZ=(8.7, 8.1, 5.4))], isKB=0)
# TRT with a heaving AND pitching rigid plate between symmetric side faces
CASES["trt_plate_pitching"] = dict(
    kind="body", dims=(20, 14, 14), bc=(101, 104, 302, 302, 301, 301), model=2, params=(3.0 / 16.0,) + (0.0,) * 9, steps=8, uvwIn=(0.05, 0.0, 0.0),
    Uref=0.05, Lref=4.0, Re=40.0, wave=1e-3, flow={}, ntolLBM=3, dtolLBM=1e-30, numsubstep=1, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, freq=0.02, XYZAmpl=(0.0, 0.6, 0.0), AoAAmpl=(0.0, 0.0, 10.0),
               AoAPhi=(0.0, 0.0, 90.0), firstXYZ=(6.3, 6.6, 5.2)), isKB=0)
# a flexible plate driven in heave and pitch at its leading edge (BASELINE configs[3] in miniature), four structural sub-steps
CASES["flexible_plate_heaving"] = dict(
    kind="body", dims=(20, 14, 14), bc=(101, 104, 301, 301, 301, 301), model=1, params=P0, steps=6, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, Lref=4.0,
    Re=40.0, wave=1e-3, flow={}, ntolLBM=3, dtolLBM=1e-30, numsubstep=4, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=2.0, psR=0.3, KB=0.05, KS=800.0, freq=0.02, XYZAmpl=(0.0, 0.8, 0.0),
               AoAAmpl=(0.0, 0.0, 10.0), AoAPhi=(0.0, 0.0, 90.0), firstXYZ=(6.3, 6.6, 5.2)), isKB=1)
# the same plate over 30 steps (120 structural sub-steps, a softer and longer beam of 8 elements): the closed fluid-structure loop amplifies a
# one-ulp difference within a few steps (that is how the Young's-modulus slip was found), so a long run is the sharp test of the whole
# chain -- IBM order, nodal loads, Newton / CG iteration path.  Oracle-against-reference only (gpu=False).
CASES["flexible_plate_heaving_30_steps"] = dict(
    CASES["flexible_plate_heaving"], steps=30, plate=dict(nEL=8, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=2.0, psR=0.3, KB=0.02, KS=600.0, freq=0.02, XYZAmpl=(0.0, 0.8, 0.0),
               AoAAmpl=(0.0, 0.0, 10.0), AoAPhi=(0.0, 0.0, 90.0), firstXYZ=(6.3, 6.6, 5.2)), gpu=False)
# configs[4] in small: TWO flexible plates in tandem, the second in the wake of the first and close enough for their stencil boxes to meet
# (Gauss-Seidel order of the penalty sweeps, Solidbody.f90:898-903; Solver over all bodies per sub-step, :386-397), one heaving, one passive,
# with a tolerance that ends some iterations early.  Oracle-against-reference only (gpu=False).
CASES["two_flexible_plates_tandem"] = dict(
    kind="body", dims=(26, 16, 14), bc=(101, 104, 301, 301, 301, 301), model=1, params=P0, steps=12, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, Lref=4.0,
    Re=40.0, wave=1e-3, flow={}, ntolLBM=6, dtolLBM=0.3, numsubstep=2, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    groups=[dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=2.0, psR=0.3, KB=0.05, KS=800.0, freq=0.02, XYZAmpl=(0.0, 0.8, 0.0),
                 firstXYZ=(6.3, 7.6, 5.2)),
            dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=1.5, psR=0.3, KB=0.03, KS=600.0, AoAo=(0.0, 0.0, 6.0), firstXYZ=(11.4, 8.2, 5.4))],
    isKB=1, gpu=False)
# structural variants of the flexible plate: explicit modulus / thickness (isKB = 0) with a hinged leading edge (rotation about z free);
# Rayleigh damping, a dissipative Newmark pair, reduced geometric stiffness and a three-dimensional incidence
CASES["flexible_plate_hinged_iskb0"] = dict(
    CASES["flexible_plate"], steps=4, isKB=0,
    group=dict(iBodyModel=2, isMotionGiven=(1, 1, 1, 1, 1, 0), denR=1.0, psR=0.3, EmR=2.0e3, tcR=0.02, AoAo=(0.0, 0.0, 12.0), firstXYZ=(6.3, 6.6, 5.2)))
CASES["flexible_plate_damped_3d"] = dict(
    CASES["flexible_plate"], steps=4, solid=dict(dampK=0.01, dampM=0.02, NewmarkGamma=0.6, NewmarkBeta=0.3025, GeoGamma=0.5, IBPenaltyAlpha=0.8),
    group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=1.0, psR=0.3, KB=0.02, KS=500.0, freq=0.03, XYZAmpl=(0.2, 0.6, 0.0), XYZPhi=(10.0, 20.0, 0.0),
               AoAo=(5.0, -7.0, 12.0), AoAAmpl=(0.0, 0.0, 8.0), AoAPhi=(0.0, 0.0, 45.0), firstXYZ=(6.3, 6.6, 5.2)))
# stencil folding at the faces (trimedindex, Solidbody.f90:834-866): one plate two cells above a moving wall (mirror 0 -> 2) and across the
# periodic z face (wrap), one below a half-way wall (0 -> 1)
CASES["plates_near_walls"] = dict(
    kind="body", dims=(20, 12, 12), bc=(101, 104, 202, 203, 301, 301), model=1, params=P0, steps=6, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=40.0, wave=1e-3, flow=dict(shearRateIn=(0.0, 2e-4, 0.0)), ntolLBM=3, dtolLBM=1e-30, numsubstep=1, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    groups=[dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, firstXYZ=(6.3, 1.4, 10.1)),
            dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, AoAo=(0.0, 0.0, -4.0), firstXYZ=(9.2, 9.8, 5.3))], isKB=0)
# MRT with every kind of face and a Smagorinsky block with a plate: collision models other than SRT next to boundaries / bodies
CASES["mrt_all_faces_mixed"] = _fluid((9, 10, 8), (101, 103, 204, 202, 201, 203), model=3, steps=8, uvwIn=(0.03, 0.0, 0.0), Uref=0.03,
                                      shearRateIn=(0.0, 4e-4, 1e-4))
# configs[2] in small, started the way a first run starts: isConCmpt = 0, i.e. from initialise_ itself (FluidDomain.f90:433-545: f_eq of the
# sheared inflow profile, evaluate_shear_velocity :1803-1810) instead of an injected state -- a rigid plate in shear inflow between moving
# walls.  Oracle-against-reference only (gpu=False): the CUDA initialise_ is held against the oracle's in tests/test_gpu_parity.py.
CASES["plate_in_shear_from_initialise"] = dict(
    kind="body", dims=(22, 14, 12), bc=(101, 104, 202, 202, 301, 301), model=1, params=P0, steps=8, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=40.0, wave=0.0, flow=dict(shearRateIn=(0.0, 3e-4, 0.0)), ntolLBM=5, dtolLBM=1e-30, numsubstep=1, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, AoAo=(0.0, 0.0, 8.0), firstXYZ=(7.3, 6.6, 4.2)), isKB=0,
    from_initialise=True, gpu=False)
CASES["les_smag_plate"] = dict(
    kind="body", dims=(20, 14, 14), bc=(101, 104, 301, 301, 301, 301), model=11, params=P0, steps=6, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0,
    Re=400.0, wave=2e-2, flow={}, ntolLBM=3, dtolLBM=1e-30, numsubstep=1, plate=dict(nEL=4, chord=4.0, span=4.0, Nspan=4),
    group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, AoAo=(0.0, 0.0, 15.0), firstXYZ=(6.3, 6.6, 5.2)), isKB=0)


# a rigid plate inside a 2:1 refined son block: the plate is carried by the son (FindCarrierFluidBlock, FluidDomain.f90:1974-2017), the IBM
# runs at the son's spacing twice per root step, the stencil folding still uses the ROOT block's boundary codes (Solidbody.f90:337)
CASES["plate_in_son"] = dict(
    kind="refine_body", dims=(16, 12, 12), bc=(101, 104, 301, 301, 301, 301), sdims=(17, 13, 13), smins=(4.0, 3.0, 3.0), scheme=1, carrier_son=True,
    model=1, params=P0, steps=5, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, Lref=4.0, Re=40.0, wave=1e-3, flow={}, ntolLBM=3, dtolLBM=1e-30, numsubstep=1,
    plate=dict(nEL=6, chord=3.0, span=3.0, Nspan=6),
    group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, AoAo=(0.0, 0.0, 10.0), firstXYZ=(6.3, 5.6, 4.7)), isKB=0)


# the reference's output files (main.f90:124-143): flow fields at the first and last step, flux and probe lines at the last step
CASES["outputs_inlet_outlet"] = dict(_fluid((12, 9, 10), (101, 104, 202, 202, 301, 301), steps=6, uvwIn=(0.04, 0.0, 0.0), Uref=0.04,
                                            shearRateIn=(0.0, 3e-4, 0.0), volumeForceIn=(5e-7, 0.0, 0.0)),
                                     outputs=True, probes=[(3.5, 4.25, 6.0), (8.0, 2.0, 1.5)])
CASES["outputs_two_blocks"] = dict(CASES["refine_linear"], steps=4, outputs=True, probes=[(6.25, 4.5, 5.0)])
# outputtype 3: running means of u and of the Reynolds stresses (calculate_turbulent_statistic_, FluidDomain.f90:1147-1172, with its
# real(4) 1/n) and the MeanFlow file next to the flow file
CASES["outputs_mean_flow"] = dict(_fluid((10, 9, 8), (301, 301, 203, 203, 301, 301), steps=6, volumeForceIn=(2e-6, 0.0, 0.0), wave=2e-2),
                                  outputs=True, outputtype=3, probes=[(4.5, 3.25, 2.0)])


# WALE and Vreman with a plate: the velocity they difference is the IBM-corrected one (calculate_interaction_force has rewritten uuu near
# the body before collision_, LBMBlockComm.f90:287-293)
CASES["les_vrem_plate"] = dict(CASES["les_smag_plate"], model=15)
CASES["les_wale_plate"] = dict(CASES["les_smag_plate"], model=14)


# a HEAVING rigid plate and a FLEXIBLE plate carried by a refined son: IBM_FEM runs inside the son's two sub-cycles with the son's time
# and spacing (LBMBlockComm.f90:307-317,320-338: dt_solid = dh_son / numsubstep)
CASES["heaving_plate_in_son"] = dict(CASES["plate_in_son"], steps=4,
                                     group=dict(iBodyModel=1, isMotionGiven=(1,) * 6, EmR=1.0, tcR=0.05, freq=0.04, XYZAmpl=(0.0, 0.4, 0.0),
                                                AoAo=(0.0, 0.0, 10.0), firstXYZ=(6.3, 5.6, 4.7)))
CASES["flexible_plate_in_son"] = dict(CASES["plate_in_son"], steps=4, numsubstep=2, isKB=1,
                                      group=dict(iBodyModel=2, isMotionGiven=(1,) * 6, denR=1.0, psR=0.3, KB=0.02, KS=500.0, AoAo=(0.0, 0.0, 10.0),
                                                 firstXYZ=(6.3, 5.6, 4.7)))


# three levels and two sons: root (dh 1) with a son (dh 0.5) that carries a grandson (dh 0.25, four sub-cycles per root step), and a second
# son elsewhere in the root -- build_block_tree / array_to_tree (LBMBlockComm.f90:98-211) and the nested sub-cycling (:307-317)
CASES["three_levels_two_sons"] = dict(
    kind="refine", dims=(18, 12, 12), bc=(101, 104, 301, 301, 301, 301), scheme=2, model=1, params=P0, steps=3, uvwIn=(0.04, 0.0, 0.0), Uref=0.04,
    Lref=4.0, Re=20.0, wave=1e-3, flow=dict(volumeForceIn=(1e-6, 0.0, 0.0)),
    sons=[dict(dims=(13, 11, 11), mins=(3.0, 3.0, 3.0), dh=0.5), dict(dims=(9, 9, 9), mins=(4.0, 4.0, 4.0), dh=0.25),
          dict(dims=(9, 9, 9), mins=(11.0, 5.0, 4.0), dh=0.5)])


def son_list(case):
    """Every block below the root: dicts with dims, mins, dh (and optionally bc, model, params)."""
    if "sons" in case:
        return case["sons"]
    if "sdims" in case:
        return [dict(dims=case["sdims"], mins=case["smins"], dh=0.5, bc=case.get("sbc", (0,) * 6), model=case.get("smodel", case["model"]),
                     params=case.get("sparams", case["params"]))]
    return []


def has_son(case):
    return bool(son_list(case))


def has_body(case):
    return "plate" in case


def case_groups(case):
    return case["groups"] if "groups" in case else [case["group"]]


def golden_path(name):
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_{name}.npz")


def initial_states(case):
    """Seeded perturbed populations per block, [19][X][Y][Z] (the state injected through ./DatContinue/continue)."""
    class _F:
        pass
    fl = _F(); fl.uvwIn = case["uvwIn"]; fl.denIn = 1.0
    out = [perturbed_state(case["dims"], fl, wave_amp=case["wave"], seed=SEED)]
    for k, sn in enumerate(son_list(case)):
        out.append(perturbed_state(sn["dims"], fl, wave_amp=case["wave"], seed=SEED + 1 + k))
    return out


def restart_states(case):
    """initial_states as check_is_continue (FluidDomain.f90:128-237) hands them to the blocks: every node takes the populations of the
    FIRST saved block, in the reference's dh-"sorted" order, that contains it, by trilinear interpolation (weights 0/1 at coincident
    nodes).  That order is what the reference's exchange loop produces (:153-165 compares the UNSORTED dh of slots i and j): with more
    than two levels it is not finest-first -- e.g. a grandson is re-gridded from its father's saved state.  The re-gridding is the
    product's host-side restatement (fsilbm3d_b200.flow_io.regrid_from_continue), which the goldens therefore pin as well."""
    from fsilbm3d_b200 import flow_io
    st = initial_states(case)
    geo = [dict(dims=case["dims"], mins=(0.0, 0.0, 0.0), dh=1.0)] + list(son_list(case))
    saved = [dict(xmin=g["mins"][0], ymin=g["mins"][1], zmin=g["mins"][2], dh=g["dh"], xDim=g["dims"][0], yDim=g["dims"][1], zDim=g["dims"][2], fIn=f)
             for g, f in zip(geo, st)]

    class _Blk:
        def __init__(self, g, f):
            self.xDim, self.yDim, self.zDim = g["dims"]
            self.xLocal, self.xOffset = self.xDim, 0
            self.xmin, self.ymin, self.zmin = g["mins"]
            self.dh, self._f = g["dh"], f

        def download_fIn(self):
            return self._f
    out = []
    for g, f in zip(geo, st):
        r = flow_io.regrid_from_continue(_Blk(g, f), saved)
        out.append(np.ascontiguousarray(f if r is None else r))
    return out


def block_list(case):
    blocks = [dict(ID=1, iCollidModel=case["model"], dims=case["dims"], dh=1.0, xyzmin=(0.0, 0.0, 0.0), BndConds=case["bc"], params=case["params"],
                   outputtype=case.get("outputtype", 1))]
    for k, sn in enumerate(son_list(case)):
        blocks.append(dict(ID=2 + k, iCollidModel=sn.get("model", case["model"]), offsetOutput=1, dims=sn["dims"], dh=sn["dh"], xyzmin=sn["mins"],
                           BndConds=sn.get("bc", (0,) * 6), params=sn.get("params", case["params"])))
    return blocks


def write_continue(path, blocks, states):
    """write_continue_blocks' layout (FluidDomain.f90:268-285)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(blocks), 0)); f.write(struct.pack("<d", 0.0))
        for b, s in zip(blocks, states):
            f.write(struct.pack("<4d", *b["xyzmin"], b["dh"])); f.write(struct.pack("<3i", *b["dims"]))
            f.write(np.ascontiguousarray(s, dtype="<f8").tobytes())


def write_inputs(case, wd, continue_at_end=False):
    """The reference's input files for this case in directory wd.  continue_at_end: ask for ./DatContinue/continue... at the last
    step (main.f90:118-120), for drivers whose end state is read from that file."""
    from fsilbm3d_b200 import solid_solver as S
    from tests.beam_cases import chain
    for d in DIRS:
        os.makedirs(os.path.join(wd, d), exist_ok=True)
    Tref = case["Lref"] / case["Uref"]
    total = (case["steps"] - 0.5) / Tref                    # main.f90:93: do while(time/Tref < timeSimTotal), dt = 1
    blocks = block_list(case)
    groups = []
    if has_body(case):
        p = case["plate"]
        S.write_plate_dat(os.path.join(wd, "plate.dat"), chain(p["nEL"] + 1, p["chord"]), 0.5 * p["span"], 0.5 * p["span"], (0.0, 0.0, 1.0),
                          Nspan=p["Nspan"])
        groups = [dict(g, fishNum=1, mesh="plate.dat") for g in case_groups(case)]
    extra = dict(timeContiDelta=case["steps"] / Tref) if continue_at_end else {}
    if case.get("outputs"):
        extra.update(timeFlowDelta=case["steps"] / Tref, timeInfoDelta=case["steps"] / Tref, fluidProbes=case["probes"], inWhichBlock=1)
    text = S.inflow_text(npsize=1, isConCmpt=0 if case.get("from_initialise") else 2, numsubstep=case.get("numsubstep", 1), timeSimTotal=total, Re=case["Re"], uvwIn=case["uvwIn"],
                         LrefType=1, Lref=case["Lref"], TrefType=0, UrefType=9, Uref=case["Uref"], ntolLBM=case.get("ntolLBM", 3),
                         dtolLBM=case.get("dtolLBM", 1e-8), interpolateScheme=case.get("scheme", 1), blocks=blocks, groups=groups,
                         isKB=case.get("isKB", 0), dtolFEM=1e-12, ntolFEM=20, **extra, **case.get("solid", {}), **case["flow"])
    with open(os.path.join(wd, "inFlow.dat"), "w") as f:
        f.write(text)
    if not case.get("from_initialise"):
        write_continue(os.path.join(wd, "DatContinue", "continue"), blocks, initial_states(case))


def run_oracle(O, case, sb=None):
    """The case on the C oracle in main.f90's order.  Bodies: `sb` is the C++ structural side opened on the same input directory
    (harness/libfsilbm_solid.so; host code on both sides of the boundary).  Returns (blocks, oracle body or None, iteration counts)."""
    nu = case["Uref"] * case["Lref"] / case["Re"]                           # Solidbody.f90:282
    fl = O.Flow(nu=nu, uvwIn=case["uvwIn"], Uref=case["Uref"], ntolLBM=case.get("ntolLBM", 3), dtolLBM=case.get("dtolLBM", 1e-8), **case["flow"])
    states = None if case.get("from_initialise") else restart_states(case)
    X, Y, Z = case["dims"]
    Fb = O.LBMBlock(X, Y, Z, dh=1.0, BndConds=case["bc"], iCollidModel=case["model"], params=case["params"], flow=fl)
    blocks = [Fb]
    nodes = [O.TreeNode(Fb)]
    root = nodes[0]
    geo = [dict(dims=case["dims"], mins=(0.0, 0.0, 0.0), dh=1.0)] + list(son_list(case))
    for sn in son_list(case):
        sx, sy, sz = sn["dims"]
        Sb = O.LBMBlock(sx, sy, sz, dh=sn["dh"], xmin=sn["mins"][0], ymin=sn["mins"][1], zmin=sn["mins"][2], BndConds=sn.get("bc", (0,) * 6),
                        iCollidModel=sn.get("model", case["model"]), params=sn.get("params", case["params"]), flow=fl)
        blocks.append(Sb)
        nodes.append(O.TreeNode(Sb))

    def upper(g, k):       # xmax of read_fuild_blocks (FluidDomain.f90:94-105): one spacing further on a periodic axis
        bc = g.get("bc", (0,) * 6)
        return g["mins"][k] + (g["dims"][k] - 1) * g["dh"] + (g["dh"] if bc[2 * k] == 301 and bc[2 * k + 1] == 301 else 0.0)

    def inside(a, b):      # block a lies in block b (CompareBlocks, LBMBlockComm.f90)
        return all(geo[b]["mins"][k] <= geo[a]["mins"][k] and upper(geo[a], k) <= upper(geo[b], k) for k in range(3))
    geo[0] = dict(geo[0], bc=case["bc"])
    for i in range(1, len(geo)):   # father = the finest coarser block that contains it (array_to_tree, LBMBlockComm.f90:135-193)
        cands = [j for j in range(len(geo)) if j != i and geo[j]["dh"] > geo[i]["dh"] and inside(i, j)]
        fa = min(cands, key=lambda j: geo[j]["dh"])
        nodes[fa].add_son(nodes[i], case["scheme"])
    for k, b in enumerate(blocks):
        b.initialise(0.0)
        if states is not None:
            b.fIn[...] = states[k]
    for b in blocks:
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    ovs, its = [], []
    if has_body(case):
        for body in sb.VBodies:
            ovs.append(O.VirtualBody(body.v_nelmts, v_move=body.v_move, iBodyModel=body.iBodyModel))
        carrier = root.sons[0] if case.get("carrier_son") else root      # FindCarrierFluidBlock: the finest block that holds the bodies
        carrier.bodies = ovs
    nsub = case.get("numsubstep", 1)

    def before_ibm(node):          # UpdatePosVelArea of FSInteraction_force (Solidbody.f90:597-600)
        if node.bodies:
            for body, ov in zip(sb.VBodies, ovs):
                body.UpdatePosVelArea()
                ov.v_Exyz[...] = body.v_Exyz; ov.v_Evel[...] = body.v_Evel; ov.v_Ea[...] = body.v_Ea

    def after_ibm(node):           # nodal loads, then IBM_FEM's sub-steps with the carrier's blktime and dh (LBMBlockComm.f90:325,332-335)
        if node.bodies:
            dh = node.block.dh
            for body, ov in zip(sb.VBodies, ovs):
                body.v_Eforce[...] = ov.v_Eforce
                body.FluidLoads()
            for isub in range(1, nsub + 1):      # Solver advances every carried body per sub-step (Solidbody.f90:386-397)
                for body in sb.VBodies:
                    body.structure(node.block.blktime, isub, dh, dh / nsub)

    for n in range(1, case["steps"] + 1):
        O.set_blktime_all(root, float(n))
        O.tree_collision_streaming_IBM_FEM(root, iters=its, before_ibm=before_ibm, after_ibm=after_ibm)
    for b in blocks:
        b.calculate_macro_quantities()
    return blocks, (ovs[0] if len(ovs) == 1 else (ovs or None)), its


def output_files(wd):
    """{relative path: bytes} of the flow / flux / probe files a run left in its work directory."""
    out = {}
    for sub, pat in (("DatFlow", "Flow"), ("DatFlow", "MeanFlow"), ("DatInfo", "FluidFlux"), ("DatInfo", "FluidProbes")):
        d = os.path.join(wd, sub)
        for fn in sorted(os.listdir(d)) if os.path.isdir(d) else ():
            if fn.startswith(pat):
                with open(os.path.join(d, fn), "rb") as f:
                    out[f"{sub}/{fn}"] = f.read()
    return out


def load(name):
    g = np.load(golden_path(name))
    return json.loads(str(g["case"])), g
