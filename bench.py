#!/usr/bin/env python
"""bench.py -- MLUPS of the fused fp64 D3Q19 collide-stream(+IBM) step on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU algorithm (oracle port)

Default workload (BASELINE.json configs[1]): periodic body-force-driven channel, no body, 256^3 cells per GPU, SRT,
tau = 0.8, volumeForceIn = (1e-6,0,0).  N > 1 weak-scales along x (x-slabs of 256 planes, global grid 256N x 256 x 256)
with the one-plane halo of the outgoing populations.  One "step" = one pass of LBMBlockComm.f90:283-303 over the block
(update_volume_force + the fused macro/force/collide/stream/boundary kernel; with bodies also calculate_interaction_force
and the host structural sub-steps).  Prints ONE JSON line (rank 0).

--workload selects the other configurations of BASELINE.json (WORKLOADS below): plate512 = configs[2] (rigid plate in shear
inflow), heave1024 = configs[3] (heaving flexible plate, 128 x-planes per GPU), school2048 = configs[4] (one flexible plate
per GPU slab), school2048r = configs[4] with every plate in its own refined son block, school8x1 = a diagnostic with the
eight plates in one block.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout must carry exactly one JSON line
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "MLUPS (fp64 D3Q19 collide-stream+IBM)"
UNIT = "MLUPS"
BYTES_PER_LU = 304.0  # 2 x 19 x 8, BASELINE.md section 2
WORKLOADS = {
    # name: (X per GPU, Y, Z, BndConds, model, plate?)
    "channel256": dict(dims=(256, 256, 256), bc=(301,) * 6, model=1, plate=False,
                       desc="configs[1]: periodic body-force channel 256^3 per GPU, SRT, no body"),
    "plate512": dict(dims=(512, 256, 256), bc=(101, 104, 202, 202, 301, 301), model=1, plate=True,
                     desc="configs[2]: rigid plate (8192 markers) in shear inflow 512x256x256, SRT, 5 IBM iterations"),
    # configs[3]: 1024x512x512 at 8 GPUs = 128 x-planes per GPU.  One heaving FLEXIBLE plate (64 beam elements x 128 span markers),
    # numsubstep = 4 structural sub-steps per fluid step on the host (C++ restatement of SolidSolver.f90), St = 0.3
    "heave1024": dict(dims=(128, 512, 512), bc=(101, 104, 301, 301, 301, 301), model=1, plate="flex", layout="one_centre",
                      group=dict(iBodyModel=2, denR=1.0, psR=0.3, KB=0.05, KS=800.0, freq=0.015, XYZAmpl=(0.0, 0.25, 0.0)), numsubstep=4,
                      desc="configs[3]: heaving flexible plate (8192 markers, numsubstep 4, 5 IBM iterations) in uniform inflow, 128x512x512 per GPU"),
    # configs[4]: 2048x512x512 at 8 GPUs = 256 x-planes per GPU, one passively flapping flexible plate per GPU slab, two staggered rows
    "school2048": dict(dims=(256, 512, 512), bc=(101, 104, 301, 301, 301, 301), model=1, plate="flex", layout="one_per_slab",
                       group=dict(iBodyModel=2, denR=1.0, psR=0.3, KB=0.05, KS=800.0, AoAo=(0.0, 0.0, 8.0)), numsubstep=2,
                       desc="configs[4]: one flexible plate (8192 markers each) per GPU slab, two staggered rows, 256x512x512 per GPU, single root block"),
    # configs[4] with "multiple LBMBlockComm blocks": the same school, every plate inside its own 2:1 refined son block (321x129x385 cells
    # of dh/2, two sub-cycles per root step, plate meshed at the son's resolution: 128 elements x 256 span markers).  A son lives whole
    # on the rank whose root slab holds it; its IBM, structural solve and father<->son transfers are local to that rank.
    "school2048r": dict(dims=(256, 512, 512), bc=(101, 104, 301, 301, 301, 301), model=1, plate="flex", layout="one_per_slab", refine=True,
                        group=dict(iBodyModel=2, denR=1.0, psR=0.3, KB=0.05, KS=800.0, AoAo=(0.0, 0.0, 8.0)), numsubstep=2,
                        desc="configs[4], multi-block: one flexible plate (32768 markers) per GPU slab, each in its own 2:1 refined son block "
                             "(321x129x385, two sub-cycles), root 256x512x512 per GPU"),
    # diagnostic: the eight plates of school2048 inside ONE 256x512x512 block (the IBM work every rank of the replicated
    # form carries at 8 GPUs, without the collectives)
    "school8x1": dict(dims=(256, 512, 512), bc=(101, 104, 301, 301, 301, 301), model=1, plate="flex", layout="lattice", lattice=(2, 4),
                      group=dict(iBodyModel=2, denR=1.0, psR=0.3, KB=0.05, KS=800.0, AoAo=(0.0, 0.0, 8.0)), numsubstep=2,
                      desc="diagnostic: eight flexible plates (8192 markers each) in one 256x512x512 block"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.rows, self._stop_evt = device, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def build_plate(F, dh, denIn):
    # configs[2]: chord 1 (64 cells), span 2, nEL=64, Nspan=128 -> 8192 markers, centred
    return F.RigidPlate(origin=(3.0, 2.0 - 0.013, 1.0 + 0.003), nEL=64, len1=dh, Nspan=128, spanlen=2.0, Lspan=0.0,
                        chord_dir=(1.0, 0.0, 0.0), span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha=1.0, denIn=denIn)


def build_flex(wl, world, rank=0):
    """The flexible plates of configs[3]/[4] through the C++ structural side (harness/libfsilbm_solid.so): writes the
    reference's two input files into a scratch directory and opens them.  Every rank holds every body (the beams are
    advanced redundantly with the all-reduced forces)."""
    import tempfile
    import numpy as np
    from fsilbm3d_b200 import solid_solver as S
    dh = 1.0 / 64.0
    Xl, Y, Z = wl["dims"]
    wd = tempfile.mkdtemp(prefix=f"fsilbm_bench_r{rank}_")
    nel = 128 if wl.get("refine") else 64                            # marker spacing = the carrier block's dh (son blocks: dh/2)
    xyz = np.zeros((nel + 1, 3)); xyz[:, 0] = np.linspace(0.0, 1.0, nel + 1)   # chord 1 = 64 root cells
    S.write_plate_dat(os.path.join(wd, "plate.dat"), xyz, 1.0, 1.0, (0.0, 0.0, 1.0), Nspan=2 * nel)   # span 2: one marker per cell
    groups = []
    if wl["layout"] == "one_centre":
        first = (0.5 * Xl * world * dh - 0.5 + 0.003, 0.5 * Y * dh - 0.013, 0.5 * Z * dh + 0.003)
        groups.append(dict(wl["group"], fishNum=1, numXYZ=(1, 1, 1), mesh="plate.dat", firstXYZ=first))
    elif wl["layout"] == "lattice":   # nx x ny plates inside ONE slab (diagnostic workload: every body on every GPU)
        nx, ny = wl["lattice"]
        for i in range(nx):
            for j in range(ny):
                first = ((i + 0.5) * Xl * world * dh / nx - 0.5 + 0.003, (j + 0.5) * Y * dh / ny - 0.013, 0.5 * Z * dh + 0.003)
                groups.append(dict(wl["group"], fishNum=1, numXYZ=(1, 1, 1), mesh="plate.dat", firstXYZ=first))
    else:
        for r in range(world):   # one group per body so that the rows can be staggered
            first = ((r + 0.5) * Xl * dh - 0.5 + 0.003, (0.375 if r % 2 == 0 else 0.625) * Y * dh - 0.013, 0.5 * Z * dh + 0.003)
            groups.append(dict(wl["group"], fishNum=1, numXYZ=(1, 1, 1), mesh="plate.dat", firstXYZ=first))
    txt = S.inflow_text(numsubstep=wl["numsubstep"], Re=100.0, uvwIn=(0.05, 0.0, 0.0), LrefType=0, UrefType=0, TrefType=0, ntolLBM=5, dtolLBM=1e-30,
                        isKB=1, dtolFEM=1e-16, ntolFEM=20, blocks=[dict(dims=(Xl * world, Y, Z), dh=dh, BndConds=wl["bc"])], groups=groups)
    with open(os.path.join(wd, "inFlow.dat"), "w") as f:
        f.write(txt)
    return S.SolidBodies("inFlow.dat", wl["bc"], cwd=wd)   # (its start-up messages go to stderr: see main())


def workload_flow(F_or_O_flow, wl):
    if wl["plate"] == "flex":
        return dict(nu=5e-4, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, ntolLBM=5, dtolLBM=1e-30, numsubstep=wl["numsubstep"]), 1.0 / 64.0
    if wl["plate"]:
        dh = 1.0 / 64.0
        gamma = 0.02 / ((wl["dims"][1] - 1) * dh)
        return dict(nu=5e-4, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, gamma, 0.0), Uref=0.05, ntolLBM=5, dtolLBM=1e-30), dh
    return dict(nu=0.1, volumeForceIn=(1e-6, 0.0, 0.0)), 1.0


# ------------------------------------------------------------------------------------------------------
def cpu_time_oracle(wl, dims, steps, warmup):
    """Times the CPU restatement of the reference's OpenMP path (oracle/; kind 'port') on `dims`."""
    from oracle import oracle as O
    import fsilbm3d_b200 as F
    flowkw, dh = workload_flow(None, wl)
    flowkw.pop("numsubstep", None)
    fl = O.Flow(**flowkw)
    X, Y, Z = dims
    b = O.LBMBlock(X, Y, Z, dh=dh, BndConds=wl["bc"], iCollidModel=wl["model"], flow=fl)
    b.initialise(0.0)
    b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    bodies, plate = [], None
    if wl["plate"] == "flex":
        # fluid + IBM on the initial markers of the first plate, re-stencilled every step as for a moving body; the
        # structural solve is host code on both arms and is left out of the CPU sample
        flowkw.pop("numsubstep", None)
        sb = build_flex(dict(wl, dims=dims), 1)
        for body in sb.VBodies[:1]:
            ov = O.VirtualBody(body.v_nelmts, v_move=1, iBodyModel=2)
            ov.v_Exyz[...] = body.v_Exyz; ov.v_Evel[...] = body.v_Evel; ov.v_Ea[...] = body.v_Ea
            bodies.append(ov)
    elif wl["plate"]:
        plate = build_plate(F, dh, fl.denIn)
        ov = O.VirtualBody(plate.body.v_nelmts, v_move=0, iBodyModel=1)
        ov.v_Exyz[...] = plate.body.v_Exyz; ov.v_Evel[...] = plate.body.v_Evel; ov.v_Ea[...] = plate.body.v_Ea
        bodies = [ov]

    def one(n):
        b.set_blktime(n * dh)
        b.step(bodies)
        b.calculate_macro_quantities()   # main.f90:107

    for n in range(warmup):
        one(n + 1)
    t0 = time.perf_counter()
    for n in range(steps):
        one(warmup + n + 1)
    dt = time.perf_counter() - t0
    return X * Y * Z * steps / dt / 1e6, dt / steps, O.lib().orc_omp_max_threads()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the Fortran cannot be built here: no Fortran
    compiler in the image; the oracle port stands in, kind 'port'), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    X, Y, Z = wl["dims"]
    # bounded sample: calibrate on a thin slab, then take the largest x extent that keeps the run in ~2 minutes
    mlups_cal, t_cal, threads = cpu_time_oracle(wl, (16, Y, Z), 2, 1)
    per_plane = t_cal / 16.0
    budget = 120.0
    if wl["plate"] == "flex":
        candidates = [c for c in (X, X // 2) if c >= 128]   # the plate (64 cells of chord) sits at the slab centre
    elif wl["plate"]:
        candidates = [X, 320]   # the plate occupies planes 192..256; 320 planes keep it well inside
    else:
        candidates = [X >> k for k in range(0, 8) if (X >> k) >= 16]
    xs = candidates[-1]
    for c in candidates:
        if per_plane * c * (args.steps + args.warmup) <= budget:
            xs = c
            break
    mlups, t_step, threads = cpu_time_oracle(wl, (xs, Y, Z), args.steps, args.warmup)
    sample = f"{xs}x{Y}x{Z} of the {X}x{Y}x{Z} grid, {args.steps} steps after {args.warmup} warm-up"
    line = {
        "metric": METRIC, "value": mlups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": args.workload, "desc": wl["desc"], "grid_per_gpu": [X, Y, Z]},
        "cpu_baseline": {"value": mlups, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference OpenMP path (Fortran not buildable here)"},
        "e2e": {"value": mlups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's real stdout; everything else written to fd 1 by native libraries (NCCL's
    version banner, the structural library's start-up messages) was diverted to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import fsilbm3d_b200 as F

    if world > 1:
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)

        def bcast(b):
            t = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                t = torch.tensor(list(b), dtype=torch.uint8)
            dist.broadcast(t, src=0)
            return bytes(t.tolist())
        F.init_process_group(rank, world, local, bcast)
    else:
        F.init_process_group(0, 1, local, None)

    def barrier():
        if world > 1:
            dist.barrier()

    wl = WORKLOADS[args.workload]
    Xl, Y, Z = wl["dims"]
    XG = Xl * world
    flowkw, dh = workload_flow(None, wl)
    flow = F.FlowCondType(**flowkw)
    blk = F.LBMBlock(XG, Y, Z, dh=dh, BndConds=wl["bc"], iCollidModel=wl["model"], flow=flow, xOffset=rank * Xl, xLocal=Xl, device=local)
    blk.initialise(0.0)
    blk.update_volume_force(); blk.set_boundary_conditions()
    plates, sb = [], None
    bodies_total = 0
    if wl["plate"] == "flex":
        sb = build_flex(wl, world, rank)
        plates = sb.plates
        bodies_total = len(plates)
        if world > 1 and not wl.get("refine"):
            # Per-rank body lists: a rank holds (feeds to the library and advances structurally) only the plates that can
            # reach its slab -- chord box widened by 16 cells; the plates are anchored at their leading edge.  The IBM call is
            # then collective (loop control all-reduced), a plate across a slab interface is held by both neighbours.
            lo, hi = (rank * Xl - 16) * dh, ((rank + 1) * Xl + 16) * dh
            plates = [p for p in plates if p.body.v_Exyz[:, 0].max() >= lo and p.body.v_Exyz[:, 0].min() <= hi]
            F._lib.check(F.lib().fsilbm_set_option(b"ibm_force_exchange", 0))
            blk.ibm_collective = True
    elif wl["plate"]:
        plates = [build_plate(F, dh, flow.denIn)]
        bodies_total = 1
    transport = blk.halo_transport
    stream = torch.cuda.ExternalStream(blk.cuda_stream, device=local)
    lib = F.lib()

    flex = sb is not None
    root, son_cells, sons = None, 0, []
    if wl.get("refine"):
        # One son block per plate of this rank's slab: 160 x 64 x 192 root cells around the plate at half the spacing, all six faces
        # fed by the father (BndConds 0); the plate is carried by the son (FluidDomain.f90:1974-2017).  Built by this rank alone.
        root = F.blockTreeNode(blk)
        mine = [p for p in plates if rank * Xl * dh <= p.body.v_Exyz[:, 0].min() < (rank + 1) * Xl * dh]
        for p in mine:
            le = p.body.v_Exyz.min(axis=0)                      # leading edge, lower span end
            smin = (np.floor((le[0] - 0.75) / dh) * dh, np.floor((p.body.v_Exyz[:, 1].mean() - 0.5) / dh) * dh, np.floor((le[2] - 0.5) / dh) * dh)
            sdims = (2 * 160 + 1, 2 * 64 + 1, 2 * 192 + 1)
            son = F.LBMBlock(*sdims, dh=0.5 * dh, xmin=float(smin[0]), ymin=float(smin[1]), zmin=float(smin[2]), BndConds=(0,) * 6,
                             iCollidModel=wl["model"], flow=flow, device=local)
            son.initialise(0.0)
            son.update_volume_force(); son.set_boundary_conditions()
            root.add_son(F.blockTreeNode(son, [p]), 1)
            sons.append(son)
            son_cells += sdims[0] * sdims[1] * sdims[2]
        plates = mine

    def step(n):
        if root is not None:
            F.set_blktime_all(root, n * dh)                     # main.f90:97
            F.tree_collision_streaming_IBM_FEM(root, solver=True)
        else:
            F.tree_collision_streaming_IBM_FEM(blk, plates, time=n * dh, solver=flex)

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    for n in range(args.warmup):
        step(n + 1)
    blk.sync(); torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.fsilbm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for n in range(args.steps):
        step(args.warmup + n + 1)
    if sb is not None:
        sb.flush()      # the structural work of the last step belongs to the timed region
    e1.record(stream)
    blk.sync(); torch.cuda.synchronize()
    launches = lib.fsilbm_launch_count() - l0
    ms = e0.elapsed_time(e1)
    structural = None
    if flex:
        b0 = plates[0].body if plates else sb.VBodies[0]
        structural = {"host_ms_per_step_all_bodies": 1e3 * (sb.host_seconds + sum(p.host_seconds for p in plates)) / (args.steps + args.warmup),
                      "cg_iterations_per_step_body0": float(b0.FishInfo[3]) / (args.steps + args.warmup), "bodies": bodies_total,
                      "bodies_held_by_rank0": len(plates),
                      "note": "C++ restatement of SolidSolver.f90 on the host (one thread per body, as the reference's OpenMP loop), overlapped with the collide-stream launch; "
                              "each rank advances the plates that reach its slab; in the reference this is the Fortran driver's own work"}
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells_total = float(XG) * Y * Z
    if root is not None:   # lattice updates of ALL blocks: a son makes two updates of its cells per root step
        extra = torch.tensor([2.0 * son_cells], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(extra, op=dist.ReduceOp.SUM)
        cells_total += float(extra.item())
    value = cells_total * args.steps / (ms * 1e-3) / 1e6

    # ---- dominant kernel alone (roofline): fused collide-stream launches back to back --------------------
    kern_ms = None
    if not wl["plate"] and world == 1:
        kern_ms = ms / args.steps    # a step of this workload IS one collide_push launch (periodic: no face kernels)
    else:
        blk.sync()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, args.steps // 4)
        old = F.lib().fsilbm_launch_count()
        k0.record(stream)
        for _ in range(reps):
            blk.collide_stream()
        k1.record(stream)
        blk.sync()
        kern_ms = k0.elapsed_time(k1) / reps
    peak, peak_src = measured_peaks()
    cells_local = float(Xl) * Y * Z
    achieved = BYTES_PER_LU * cells_local / (kern_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "collide_push_kernel", "algorithmic_bytes_per_launch": BYTES_PER_LU * cells_local, "peak_source": peak_src,
                "kernel_ms": kern_ms}

    # ---- end to end through the host API with HOST buffers ------------------------------------------------
    # What the reference driver does around this path: populations come from pinned host memory once
    # (check_is_continue / initialise), then every step goes through the public API with host arguments (the
    # time scalar; with a body 7n marker doubles down and 3n force doubles up inside
    # fsilbm_ibm_interaction_force), and every `flow_every` steps den+uuu are read back to pinned host memory
    # (what write_flow_ needs; main.f90:130 timeFlowDelta cadence).  All copies are inside the timed region.
    flow_every = args.flow_every
    f_host = torch.empty((19, Xl, Y, Z), dtype=torch.float64, pin_memory=True)
    den_host = torch.empty((Xl, Y, Z), dtype=torch.float64, pin_memory=True)
    uuu_host = torch.empty((3, Xl, Y, Z), dtype=torch.float64, pin_memory=True)
    blk.download_fIn(f_host.numpy())
    blk.sync(); torch.cuda.synchronize(); barrier()
    check = F._lib.check
    t0 = time.perf_counter()
    blk.upload_fIn(f_host.numpy())
    n_out = 0
    for n in range(args.steps):
        step(args.warmup + args.steps + n + 1)
        if (n + 1) % flow_every == 0 or n + 1 == args.steps:
            # asynchronous read-back (the reference forks its writer): the copy overlaps the following steps; the previous
            # one is waited for before the host arrays are reused
            blk.download_macro_async(den_host.numpy(), uuu_host.numpy())
            n_out += 1
    if sb is not None:
        sb.flush()
    blk.download_wait()
    blk.sync(); torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = cells_total * args.steps / e2e_s / 1e6
    markers = float(sum(p.body.v_nelmts for p in plates))   # held by this rank
    if world > 1:
        t = torch.tensor([markers], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        markers = float(t.item())
    h2d = f_host.numel() * 8 * world + 7 * 8 * markers * args.steps
    d2h = (den_host.numel() + uuu_host.numel()) * 8 * world * n_out + 3 * 8 * markers * args.steps
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
           "segment": f"fIn uploaded from pinned host once, {args.steps} steps through the LBMBlock API with host arguments, den+uuu read back "
                      f"to pinned host every {flow_every} steps ({n_out} read-backs, asynchronous: each overlaps the following steps and is waited for "
                      f"before the next one and at the end); wall clock, bytes averaged per step"}

    # ---- CPU baseline on this box's cores (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sx = Xl if (not wl["plate"] or wl["plate"] == "flex") else 320
            steps_cpu = 6
            mlups, t_step, threads = cpu_time_oracle(wl, (sx, Y, Z), steps_cpu, 2)
            cpu = {"value": mlups, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{sx}x{Y}x{Z}, {steps_cpu} steps after 2 warm-up ({t_step * 1e3:.0f} ms/step); "
                             "C restatement of the reference OpenMP path (Fortran not buildable here)"}
        except Exception as ex:   # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "grid_per_gpu": [Xl, Y, Z], "grid_global": [XG, Y, Z],
                       "decomposition": f"x-slabs x{world}" if world > 1 else "single block",
                       "l2": "working set 5.1 GB per GPU (two population buffers) >> 126 MB L2; no explicit flush needed",
                       "kernel_variant": args.variant, "halo_transport": transport,
                       "structural_solver": structural,
                       "blocks": None if root is None else {"root_cells_per_gpu": Xl * Y * Z, "son_cells_on_rank0": son_cells, "son_updates_per_root_step": 2,
                                                            "note": "value counts the lattice updates of all blocks (root + 2 x son); roofline and e2e "
                                                                    "transfers are the root block's"},
                       "ibm": ("ordered per-cell gather (bit-identical to the serial reference)" if args.ibm_ordered else "fp64 atomics") if wl["plate"] else None},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    if root is not None:
        for pair in root.comm:
            pair.close()
        for son in sons:
            son.close()
    blk.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="channel256", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flow-every", type=int, default=200, help="e2e leg: read den,uuu back to the host every this many steps")
    ap.add_argument("--halo", type=int, default=1, choices=[0, 1], help="multi-GPU halo transport: 1 peer stores over NVLink, 0 NCCL send/recv")
    ap.add_argument("--ibm-single-launch", type=int, default=1, choices=[0, 1], help="IBM penalty iteration: 1 one cooperative kernel, 0 one kernel per phase")
    ap.add_argument("--ibm-ordered", type=int, default=1, choices=[0, 1],
                    help="IBM spreading: 1 ordered per-cell gather (bit-identical to the serial reference), 0 fp64 atomics")
    ap.add_argument("--opt", action="append", default=[], help="library option key=int (fsilbm_set_option), repeatable")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    import fsilbm3d_b200 as F
    if args.variant:
        F._lib.check(F.lib().fsilbm_set_option(b"variant", args.variant))
    F._lib.check(F.lib().fsilbm_set_option(b"halo", args.halo))
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_ordered", args.ibm_ordered))
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_single_launch", args.ibm_single_launch))
    for kv in args.opt:
        k, v = kv.split("=")
        F._lib.check(F.lib().fsilbm_set_option(k.encode(), int(v)))
    run_gpu(args)


if __name__ == "__main__":
    main()
