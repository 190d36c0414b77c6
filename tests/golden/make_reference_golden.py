#!/usr/bin/env python
"""Generates tests/golden/ref_<case>.npz: the state the REFERENCE ITSELF leaves on the cases of tests/reference_cases.py.

The reference is Fortran and this image has no Fortran compiler, so its unmodified sources (/root/reference/src/*.f90, read
where they lie) are executed by the interpreter in oracle/ftn/ -- main.f90 from start to end, input files in, module state out.
Run here (the GPU box has no /root/reference):

    python tests/golden/make_reference_golden.py [case ...]

Each file holds the case description (JSON), the final populations / density / velocity of every block, the FIELDSTAT lines the
program printed and, for body cases, the markers, marker forces and nodal beam state."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ftn.run import block_states, body_states, run_main   # noqa: E402
from tests import reference_cases as RC                          # noqa: E402


def make(name):
    case = RC.CASES[name]
    wd = tempfile.mkdtemp(prefix=f"fsilbm_ref_{name}_")
    RC.write_inputs(case, wd)
    t0 = time.time()
    I = run_main(wd)
    out = dict(case=json.dumps(case), fieldstat=np.array([l for l in I.io.stdout_lines if "FIELDSTAT" in l]))
    steps = [l for l in I.io.stdout_lines if l.startswith(" Steps:")]
    assert len(steps) == case["steps"], (len(steps), case["steps"])
    for k, (f, den, uuu) in enumerate(block_states(I)):
        out[f"fIn{k}"], out[f"den{k}"], out[f"uuu{k}"] = f, den, uuu
    for k, b in enumerate(body_states(I)):
        for key, v in b.items():
            out[f"body{k}_{key}"] = v
    # how often the reference entered PenaltyForce_ (Solidbody.f90:981): sum over the steps of iterLBM x number of bodies
    out["penalty_calls"] = np.array(I.modules["solidbody"].procs["penaltyforce_"].ncalls)
    if case.get("outputs"):      # the files the reference wrote, byte for byte
        for rel, data in RC.output_files(wd).items():
            out["file:" + rel] = np.frombuffer(data, dtype=np.uint8)
    fl = I.modules["flowcondition"].vars["flow"].f
    out["derived"] = json.dumps(dict(nu=float(fl["nu"]), Uref=float(fl["uref"]), Lref=float(fl["lref"]), Tref=float(fl["tref"])))
    np.savez_compressed(RC.golden_path(name), **out)
    print(f"{name}: {case['steps']} steps in {time.time() - t0:.1f} s -> {os.path.relpath(RC.golden_path(name), ROOT)}", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(RC.CASES)):
        make(n)
