"""The ISO_C_BINDING shim fortran/fsilbm_gpu.f90 EXECUTED against libfsilbm_b200.so.

This image has no Fortran compiler, so the shim cannot be compiled here; instead the small Fortran interpreter of oracle/ftn/
(test infrastructure) runs a Fortran driver program (tests/fortran/drive_shim.f90) that uses the shim the way INTEGRATION.md edits
the reference's main loop.  bind(C) interface calls are forwarded to the real shared library through ctypes with the argument
association the standard prescribes (oracle/ftn/cbind.py: VALUE dummies by value, everything else by reference, bind(C) derived
types as C structs, type(c_ptr) as void*), so argument order, kinds and by-value/by-reference conventions of every interface the
driver touches are exercised for real.

CPU: the shim parses completely with the interpreter's front end, and without a GPU the run ends the reference's way
(`write(*,*) message; stop`).  GPU: a periodic body-force channel and a shear flow past a plate, populations and marker forces
bit-identical to the CPU oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "fortran", "fsilbm_gpu.f90")
DRIVER = os.path.join(ROOT, "tests", "fortran", "drive_shim.f90")


def write_case(wd, dims, steps, model, bc, nu, Uref, uvwIn, shear, vforce, f0, markers=None, ntol=3, dtol=1e-30):
    X, Y, Z = dims
    nm = 0 if markers is None else len(markers[2])
    with open(os.path.join(wd, "case.txt"), "w") as f:
        f.write(f"{X} {Y} {Z} {steps} {model}\n" + " ".join(str(b) for b in bc) + "\n" + f"{nu!r} 1.0 {Uref!r}\n")
        for v in (uvwIn, shear, vforce):
            f.write(" ".join(repr(float(x)) for x in v) + "\n")
        f.write(f"{nm} {ntol} {dtol!r}\n")
    np.ascontiguousarray(f0).tofile(os.path.join(wd, "f0.bin"))      # C [19][X][Y][Z] = Fortran fIn(z,y,x,0:18)
    if markers is not None:
        with open(os.path.join(wd, "markers.bin"), "wb") as f:
            for a in markers:                                        # (n,3) C arrays = Fortran (3,n)
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())


def run_driver(wd):
    import fsilbm3d_b200 as F
    from oracle.ftn.interp import Interp
    I = Interp(cwd=wd)
    I.bind_c_library(F.library_path())
    I.load([SHIM, DRIVER])
    stop = I.run_program()
    return stop, I.io.stdout_lines


def test_shim_parses_with_the_fortran_front_end():
    from oracle.ftn.parse import parse_source
    units = parse_source(open(SHIM).read(), SHIM)
    mods = {u.name: u for u in units}
    import fsilbm3d_b200 as F
    assert sorted(p.cname for p in mods["fsilbm_c"].interfaces) == sorted(F.declared_symbols())
    assert len(mods["fsilbm_gpu"].procs) >= 50


def test_driver_over_the_shim_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    wd = str(tmp_path)
    write_case(wd, (8, 6, 6), 2, 1, (301,) * 6, 0.1, 0.05, (0, 0, 0), (0, 0, 0), (1e-6, 0, 0), np.zeros((19, 8, 6, 6)))
    stop, lines = run_driver(wd)
    # gpu_init -> fsilbm_init fails -> fsilbm_check: write(*,*) fsilbm_last_error() ; stop  (the reference's error convention)
    assert stop is not None and stop.startswith("STOP")
    assert any("no CUDA device" in l or "no CPU path" in l for l in lines), lines


@pytest.mark.gpu
@pytest.mark.parametrize("with_plate", [False, True], ids=["periodic_channel", "plate_in_shear_flow"])
def test_driver_over_the_shim_matches_the_oracle(oracle, tmp_path, with_plate):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fsilbm3d_b200 as F
    from tests.common import perturbed_state
    wd = str(tmp_path)
    if with_plate:
        dims, bc, steps = (24, 20, 16), (101, 104, 202, 202, 301, 301), 8
        flowkw = dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, 2e-4, 0.0), Uref=0.05, ntolLBM=3, dtolLBM=1e-30)
        plate = F.RigidPlate(origin=(8.3, 8.2, 4.4), nEL=6, len1=1.0, Nspan=8, spanlen=8.0, Lspan=0.0, chord_dir=(1.0, 0.3, 0.0), denIn=1.0)
        markers = (plate.body.v_Exyz, plate.body.v_Evel, plate.body.v_Ea)
    else:
        dims, bc, steps = (16, 12, 10), (301,) * 6, 12
        flowkw = dict(nu=0.1, volumeForceIn=(1e-6, 2e-7, -3e-7), Uref=0.05, ntolLBM=3, dtolLBM=1e-30)
        markers = None
    of = oracle.Flow(**flowkw)
    f0 = perturbed_state(dims, of)
    write_case(wd, dims, steps, 1, bc, flowkw["nu"], flowkw["Uref"], flowkw.get("uvwIn", (0, 0, 0)), flowkw.get("shearRateIn", (0, 0, 0)),
               flowkw.get("volumeForceIn", (0, 0, 0)), f0, markers)
    stop, lines = run_driver(wd)
    assert stop is None, (stop, lines[-5:])
    ob = oracle.LBMBlock(*dims, BndConds=bc, flow=of)
    ob.initialise(0.0)
    ob.fIn[...] = f0
    ob.update_volume_force(); ob.set_boundary_conditions(); ob.calculate_macro_quantities()
    ovs = []
    if with_plate:
        ov = oracle.VirtualBody(len(markers[2]), v_move=0, iBodyModel=1)
        ov.v_Exyz[...] = markers[0]; ov.v_Evel[...] = markers[1]; ov.v_Ea[...] = markers[2]
        ovs = [ov]
    its = 0
    for n in range(1, steps + 1):
        ob.set_blktime(float(n))
        its += ob.step(ovs)
    f1 = np.fromfile(os.path.join(wd, "f1.bin")).reshape(ob.fIn.shape)
    assert np.array_equal(f1, ob.fIn), f"max |df| {np.abs(f1 - ob.fIn).max():.3e}"
    ob.calculate_macro_quantities()                       # main.f90:107 before computeFieldStat_blocks (:150)
    st = ob.ComputeFieldStat()
    assert lines[-1] == f" FIELDSTAT L2 u {st[0]:18.12f}"
    assert lines[-2] == f" tau {3.0 * flowkw['nu'] + 0.5:10.6f} IBM iterations {its:6d}"
    if with_plate:
        force = np.fromfile(os.path.join(wd, "force.bin")).reshape(-1, 3)
        assert np.array_equal(force, np.array(ovs[0].v_Eforce))
