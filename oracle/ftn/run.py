#!/usr/bin/env python
"""Runs the reference's main program (main.f90) from its own, unmodified Fortran sources in a work directory that holds the
reference's input files (inFlow.dat, mesh files, optionally ./DatContinue/continue).  TEST INFRASTRUCTURE.

    python -m oracle.ftn.run WORKDIR [--ref /root/reference] [--echo]

The sources are read where they lie (never copied).  The run leaves the files the reference writes (DatFlow/, DatContinue/,
DatInfo/, ...); run_main() also returns the interpreter so callers can read the final state from the module variables."""
from __future__ import annotations

import argparse
import os
import sys
import time

DIRS = ("DatFlow", "DatContinue", "DatInfo", "DatBody", "DatBodySpan", "DatTemp", "DatOthe")


def run_main(workdir, ref="/root/reference", echo=False):
    from .interp import load_reference
    for d in DIRS:
        os.makedirs(os.path.join(workdir, d), exist_ok=True)
    I = load_reference(ref, cwd=workdir, echo=echo)
    stop = I.run_program()
    if stop is not None and stop.strip() not in ("STOP", "STOP 0"):
        raise RuntimeError(f"the reference program stopped: {stop}\n" + "\n".join(I.io.stdout_lines[-15:]))
    return I


def block_states(I):
    """[(fIn [19][X][Y][Z], den [X][Y][Z], uuu [3][X][Y][Z])] of every LBMblks(i), copied out of the interpreter's memory."""
    import numpy as np
    out = []
    for b in I.modules["fluiddomain"].vars["lbmblks"].d:
        out.append((np.ascontiguousarray(b.f["fin"].d.transpose(3, 2, 1, 0)), np.ascontiguousarray(b.f["den"].d.transpose(2, 1, 0)),
                    np.ascontiguousarray(b.f["uuu"].d.transpose(3, 2, 1, 0))))
    return out


def body_states(I):
    """Per VBodies(i): marker positions / velocities / forces as (n,3) arrays and the beam's nodal pos / vel / acc."""
    import numpy as np
    vb = I.modules["solidbody"].vars.get("vbodies")
    out = []
    if vb is None:
        return out
    for b in vb.d:
        rbm = b.f["rbm"]
        out.append(dict(v_Exyz=np.ascontiguousarray(b.f["v_exyz"].d.T), v_Evel=np.ascontiguousarray(b.f["v_evel"].d.T),
                        v_Eforce=np.ascontiguousarray(b.f["v_eforce"].d.T), v_Ea=np.array(b.f["v_ea"].d),
                        pos=np.ascontiguousarray(rbm.f["pos"].d.T), vel=np.ascontiguousarray(rbm.f["vel"].d.T), acc=np.ascontiguousarray(rbm.f["acc"].d.T)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workdir")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--echo", action="store_true", help="print what the program writes to standard output")
    a = ap.parse_args()
    t0 = time.time()
    I = run_main(a.workdir, a.ref, a.echo)
    for l in I.io.stdout_lines:
        if "FIELDSTAT" in l:
            print(l)
    print(f"reference main program finished in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    sys.exit(main())
