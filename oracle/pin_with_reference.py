#!/usr/bin/env python
"""Cross-checks the committed reference goldens against a GFORTRAN BUILD of the reference.  Test infrastructure only.

tests/golden/ref_*.npz hold what the reference's own main program leaves on the 38 cases of tests/reference_cases.py; they were
produced in an image without a Fortran compiler by executing the reference's unmodified sources with the interpreter in
oracle/ftn/ (DESIGN.md section 5).  Where gfortran exists this script closes the remaining gap (interpreter vs compiler):

    python oracle/pin_with_reference.py [--ref /root/reference] [--keep]

  1. `make -C oracle _ref` builds the unmodified reference (its own flags, sources where they lie) into oracle/_ref/FSILBM3D;
  2. every case's input directory (inFlow.dat, plate.dat, injected ./DatContinue/continue; the files the interpreter run read) is
     given to that binary with OMP_NUM_THREADS=1; it leaves ./DatContinue/continue<time> = every block's fp64 populations
     (FluidDomain.f90:268-285) at the last step;
  3. the populations are compared with the committed golden file: bit-identical is expected for an x86-64 build without -march
     (no fused multiply-adds); the script prints the number of differing values and the largest relative difference and exits 0
     only if every case is within 1e-12;
  4. the two-block case of tests/golden/inFlow_two_blocks.dat (100 root steps) is run as well and compared with the C oracle.
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SAMPLE = os.path.join(ROOT, "tests", "golden", "inFlow_two_blocks.dat")
DIRS = ("DatFlow", "DatContinue", "DatInfo", "DatBody", "DatBodySpan", "DatTemp", "DatOthe")


def read_continue(path):
    """write_continue_blocks, FluidDomain.f90:268-285 (stream access): nblocks, step, time; per block xmin ymin zmin dh, xDim yDim
    zDim, fIn(z,y,x,0:18) -- i.e. C order [19][X][Y][Z]."""
    with open(path, "rb") as f:
        nblocks, step = struct.unpack("<ii", f.read(8))
        (time,) = struct.unpack("<d", f.read(8))
        blocks = []
        for _ in range(nblocks):
            geo = struct.unpack("<4d", f.read(32))
            X, Y, Z = struct.unpack("<3i", f.read(12))
            fIn = np.frombuffer(f.read(8 * 19 * X * Y * Z), dtype="<f8").reshape(19, X, Y, Z)
            blocks.append((geo, (X, Y, Z), fIn))
    return step, time, blocks


def oracle_two_blocks(nsteps):
    from oracle import oracle as O
    fl = O.Flow(nu=0.04 * 8.0 / 100.0, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, volumeForceIn=(1e-6, 0.0, 0.0), ntolLBM=3, dtolLBM=1e-8)
    Fb = O.LBMBlock(24, 16, 16, dh=1.0, BndConds=(101, 104, 301, 301, 301, 301), flow=fl)
    Sb = O.LBMBlock(17, 13, 13, dh=0.5, xmin=6.0, ymin=4.0, zmin=4.0, BndConds=(0,) * 6, flow=fl)
    Fb.initialise(0.0); Sb.initialise(0.0)
    root = O.TreeNode(Fb); root.add_son(O.TreeNode(Sb), 1)
    for b in (Fb, Sb):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for n in range(1, nsteps + 1):
        O.set_blktime_all(root, float(n))
        O.tree_collision_streaming_IBM_FEM(root)
    return [Fb, Sb]


def run_exe(exe, wd):
    for d in DIRS:
        os.makedirs(os.path.join(wd, d), exist_ok=True)
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
        raise SystemExit(1)
    files = sorted(f for f in glob.glob(os.path.join(wd, "DatContinue", "continue*")) if not f.endswith("continue"))
    return r, read_continue(files[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--keep", action="store_true", help="keep the scratch directories")
    args = ap.parse_args()
    if shutil.which("gfortran") is None:
        print("pin_with_reference: no gfortran on PATH; the reference cannot be compiled here.  The committed goldens come from the "
              "interpreter run of its sources (oracle/ftn, DESIGN.md section 5); this cross-check needs a machine with gfortran.")
        return 2
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref", f"REF={args.ref}"], check=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "FSILBM3D")
    from tests import reference_cases as RC
    ok = True
    for name, case in RC.CASES.items():
        g = np.load(RC.golden_path(name))
        wd = tempfile.mkdtemp(prefix=f"fsilbm_pin_{name}_")
        RC.write_inputs(case, wd, continue_at_end=True)
        r, (step, _, blocks) = run_exe(exe, wd)
        assert step == case["steps"], (step, case["steps"])
        for k, (_, dims, fIn) in enumerate(blocks):
            ref = g[f"fIn{k}"]
            ndiff = int((fIn != ref).sum())
            rel = float(np.abs(fIn - ref).max() / np.abs(ref).max())
            print(f"{name} block {k} {dims}: {ndiff} of {fIn.size} populations differ from the committed golden, max rel diff {rel:.3e}")
            ok &= rel <= 1e-12
        if not args.keep:
            shutil.rmtree(wd, ignore_errors=True)
    wd = tempfile.mkdtemp(prefix="fsilbm_pin_two_blocks_")
    shutil.copy(SAMPLE, os.path.join(wd, "inFlow.dat"))
    r, (step, _, blocks) = run_exe(exe, wd)
    for (geo, dims, fIn), ob in zip(blocks, oracle_two_blocks(step)):
        ndiff = int((fIn != ob.fIn).sum())
        rel = float(np.abs(fIn - ob.fIn).max() / np.abs(ob.fIn).max())
        print(f"two_blocks {dims}: {ndiff} of {fIn.size} populations differ from the C oracle, max rel diff {rel:.3e}")
        ok &= rel <= 1e-12
    print("FIELDSTAT lines of the last run:")
    print("\n".join(l for l in r.stdout.splitlines() if "FIELDSTAT" in l)[-800:])
    print("goldens CONFIRMED by the gfortran build" if ok else "gfortran build and committed goldens DIFFER")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
