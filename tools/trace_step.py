#!/usr/bin/env python
"""Time-line of a few steps of a bench workload on one GPU: every launch of the step is followed by a timing event on its
stream (library option "trace"), the dump gives per mark the device completion time and the host issue time.  Diagnostic
only (the events perturb the run slightly); bench.py is the measurement.

    python tools/trace_step.py --workload plate512 --steps 40 --trace-steps 6 --out gpurun_out/trace_plate512.csv
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="plate512")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--trace-steps", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/trace.csv")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    import fsilbm3d_b200 as F
    F.init_process_group(0, 1, 0, None)
    lib, check = F.lib(), F._lib.check
    for kv in args.opt:
        k, v = kv.split("=")
        check(lib.fsilbm_set_option(k.encode(), int(v)))
    wl = bench.WORKLOADS[args.workload]
    Xl, Y, Z = wl["dims"]
    flowkw, dh = bench.workload_flow(None, wl)
    flow = F.FlowCondType(**flowkw)
    blk = F.LBMBlock(Xl, Y, Z, dh=dh, BndConds=wl["bc"], iCollidModel=wl["model"], flow=flow)
    blk.initialise(0.0); blk.update_volume_force(); blk.set_boundary_conditions()
    sb = bench.build_flex(wl, 1, 0) if wl["plate"] == "flex" else None
    plates = sb.plates if sb else ([bench.build_plate(F, dh, flow.denIn)] if wl["plate"] else [])
    flex = sb is not None
    for n in range(args.steps):
        F.tree_collision_streaming_IBM_FEM(blk, plates, time=(n + 1) * dh, solver=flex)
    if sb:
        sb.flush()
    blk.sync()
    check(lib.fsilbm_set_option(b"trace", 1))
    for n in range(args.trace_steps):
        F.tree_collision_streaming_IBM_FEM(blk, plates, time=(args.steps + n + 1) * dh, solver=flex)
    if sb:
        sb.flush()
    blk.sync()
    check(lib.fsilbm_set_option(b"trace", 0))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    check(lib.fsilbm_trace_dump(args.out.encode()))
    rows = [ln.strip().split(",") for ln in open(args.out)][1:]
    prev = {}
    print(f"{'mark':24s} {'stream':>6s} {'device_done_us':>15s} {'since_prev_on_stream':>21s} {'host_issued_us':>15s}")
    for name, st, dev, host in rows:
        d = float(dev)
        print(f"{name:24s} {st:>6s} {d:15.1f} {d - prev.get(st, d):21.1f} {float(host):15.1f}")
        prev[st] = d
    blk.close()


if __name__ == "__main__":
    main()
