#!/usr/bin/env python
"""Pins the CPU oracle against the REAL reference -- for an environment that has gfortran.  Test infrastructure only.

This image has no Fortran compiler, so the script cannot run here and the oracle stays "parity unpinned" (DESIGN.md "Oracle").
Where gfortran exists:

    python oracle/pin_with_reference.py [--ref /root/reference] [--keep]

  1. `make -C oracle _ref` builds the unmodified reference (its own flags, sources where they lie) into oracle/_ref/FSILBM3D;
  2. the two-block case of tests/golden/inFlow_two_blocks.dat (root 24x16x16 with inlet/outlet, a 2:1 refined son, 100 root
     steps; the case the stand-in driver is tested on, tests/test_gpu_harness.py) is run by the reference in a scratch
     directory; it leaves ./DatContinue/continue0000050000 = every block's full fp64 populations (FluidDomain.f90:268-285);
  3. the oracle runs the same case (oracle.TreeNode, main.f90's order);
  4. the populations are compared block by block: bit-exact is expected for a gfortran -O3 x86-64 build without -march
     (no fused multiply-adds -- the assumption behind the oracle's -ffp-contract=off); the script prints the number of
     differing values and the largest relative difference of fIn, den and uuu, and exits 0 only within 1e-12.
"""
from __future__ import annotations

import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--keep", action="store_true", help="keep the scratch directory")
    args = ap.parse_args()
    if shutil.which("gfortran") is None:
        print("pin_with_reference: no gfortran on PATH; the reference cannot be built (parity stays unpinned)")
        return 2
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref", f"REF={args.ref}"], check=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "FSILBM3D")
    wd = tempfile.mkdtemp(prefix="fsilbm_pin_")
    shutil.copy(SAMPLE, os.path.join(wd, "inFlow.dat"))
    for d in ("DatFlow", "DatContinue", "DatInfo", "DatBody", "DatBodySpan", "DatTemp", "DatOthe"):
        os.makedirs(os.path.join(wd, d), exist_ok=True)
    r = subprocess.run([exe], cwd=wd, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="2"))
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
        return 1
    step, time, blocks = read_continue(os.path.join(wd, "DatContinue", "continue0000050000"))
    assert step == 100, step
    ok = True
    for (geo, dims, fIn), ob in zip(blocks, oracle_two_blocks(step)):
        assert dims == (ob.xDim, ob.yDim, ob.zDim), (dims, ob.xDim)
        ndiff = int((fIn != ob.fIn).sum())
        rel = float(np.abs(fIn - ob.fIn).max() / np.abs(ob.fIn).max())
        den_r = fIn.sum(axis=0)
        ob.calculate_macro_quantities()
        e_den = float(np.abs(den_r - ob.den).max() / np.abs(ob.den).max())
        print(f"block {dims}: {ndiff} of {fIn.size} populations differ, max rel diff fIn {rel:.3e}, den {e_den:.3e}")
        ok &= rel <= 1e-12 and e_den <= 1e-12
    print("FIELDSTAT lines of the reference run:")
    print("\n".join(l for l in r.stdout.splitlines() if "FIELDSTAT" in l or "field" in l.lower())[-800:])
    if not args.keep:
        shutil.rmtree(wd, ignore_errors=True)
    print("oracle PINNED against the reference on this case" if ok else "oracle and reference DIFFER")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
