#!/usr/bin/env python
"""Where a step with immersed bodies spends its time: host phases (perf_counter) and device phases (stream syncs
between the phases in a second pass).  Diagnostic only; bench.py is the measurement.

    python tools/step_breakdown.py --workload school8x1 --steps 60
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="school8x1")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    args = ap.parse_args()
    import torch
    import fsilbm3d_b200 as F
    F.init_process_group(0, 1, 0, None)
    wl = bench.WORKLOADS[args.workload]
    Xl, Y, Z = wl["dims"]
    flowkw, dh = bench.workload_flow(None, wl)
    flow = F.FlowCondType(**flowkw)
    blk = F.LBMBlock(Xl, Y, Z, dh=dh, BndConds=wl["bc"], iCollidModel=wl["model"], flow=flow)
    blk.initialise(0.0); blk.update_volume_force(); blk.set_boundary_conditions()
    sb = bench.build_flex(wl, 1, 0) if wl["plate"] == "flex" else None
    plates = sb.plates if sb else [bench.build_plate(F, dh, flow.denIn)]
    nsub = flow.numsubstep

    def run(sync_between: bool, steps: int, first: int):
        T = dict(update_pos=0.0, ibm_call=0.0, collide_launch=0.0, host_advance=0.0, total=0.0)
        t00 = time.perf_counter()
        for n in range(steps):
            blk.set_blktime((first + n) * dh); blk.update_volume_force()
            t0 = time.perf_counter()
            for p in plates: p.UpdatePosVelArea()
            t1 = time.perf_counter()
            blk.calculate_interaction_force([p.body for p in plates], blk.BndConds)
            t2 = time.perf_counter()
            blk.collide_stream()
            if sync_between: blk.sync()
            t3 = time.perf_counter()
            if sb:
                sb.advance([p.body.index for p in plates], blk.blktime, nsub, blk.dh)
            else:
                for p in plates: p.structure(blk.blktime, 1, blk.dh, blk.dh)
            t4 = time.perf_counter()
            T["update_pos"] += t1 - t0; T["ibm_call"] += t2 - t1; T["collide_launch"] += t3 - t2; T["host_advance"] += t4 - t3
        blk.sync()
        T["total"] = time.perf_counter() - t00
        return {k: 1e3 * v / steps for k, v in T.items()}

    run(False, args.warmup, 1)
    a = run(False, args.steps, 1 + args.warmup)
    b = run(True, args.steps, 1 + args.warmup + args.steps)
    print(f"{args.workload}: {len(plates)} bodies, ms per step")
    print("  async   :", {k: round(v, 3) for k, v in a.items()})
    print("  synced  :", {k: round(v, 3) for k, v in b.items()}, "(collide_launch includes the kernel)")
    blk.close()


if __name__ == "__main__":
    main()
