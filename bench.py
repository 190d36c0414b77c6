#!/usr/bin/env python
"""bench.py -- MLUPS of the fused fp64 D3Q19 collide-stream(+IBM) step on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU algorithm (oracle port)

Default workload: the metric is collide-stream + IBM, so the default is a configuration WITH a body.  --gpus 1 ->
plate512 = BASELINE.json configs[2] (rigid 8192-marker plate in shear inflow, 512x256x256: the largest single-GPU
configuration).  --gpus N > 1 -> heave1024 = configs[3] (heaving flexible plate, 128x512x512 per GPU, x-slabs; 1024x512x512 at
8 GPUs).  Both hold 33.55 M cells per GPU, so the driver's weak-scaling ratio compares like with like.  channel256
(configs[1], no body) is --workload channel256.  One "step" = one pass of LBMBlockComm.f90:279-338 over the block
(update_volume_force, calculate_interaction_force, the fused macro/force/collide/stream/boundary update, and for flexible
bodies the host structural sub-steps).  Prints ONE JSON line (rank 0).

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


def host_cpu():
    """Cores this process may use, physical cores among them and the CPU model (BASELINE.md section 3 asks for all three)."""
    try:
        usable = sorted(os.sched_getaffinity(0))
    except Exception:
        usable = list(range(os.cpu_count() or 1))
    model, phys, cur = "unknown", set(), {}
    try:
        for ln in open("/proc/cpuinfo"):
            if ":" not in ln:
                if cur:
                    if int(cur.get("processor", -1)) in usable:
                        phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                    cur = {}
                continue
            k, v = [t.strip() for t in ln.split(":", 1)]
            cur[k] = v
            if k == "model name":
                model = v
        if cur and int(cur.get("processor", -1)) in usable:
            phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
    except Exception:
        pass
    return {"logical": len(usable), "physical": len(phys) or len(usable), "model": model}


def default_workload(gpus: int) -> str:
    return "plate512" if gpus <= 1 else "heave1024"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.rows, self._stop_evt = device, [], threading.Event()

    def _run_nvml(self) -> bool:
        """Polls NVML directly (a few hundred samples per second, so that a 20-step timed region of 30 ms still holds several);
        False if NVML is not usable here, then nvidia-smi is polled instead."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self._stop_evt.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                except Exception:
                    r = 0
                self.rows.append([str(self.device), str(sm), str(mx), "", ""] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            self._stop_evt.wait(0.003)
        try:
            pynvml.nvmlShutdown()
        except Exception:
            pass
        return True

    def run(self):
        if self._run_nvml():
            return
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
def cpu_time_oracle(wl, dims, steps, warmup, with_bodies=True):
    """Times the CPU restatement of the reference's OpenMP path (oracle/; kind 'port') on `dims`.  with_bodies=False: the fluid
    part only (calibration on a slab too thin to hold the plate)."""
    from oracle import oracle as O
    import fsilbm3d_b200 as F
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm asks for every core this process may use itself
    O.lib().orc_omp_set_threads(host_cpu()["logical"])
    flowkw, dh = workload_flow(None, wl)
    flowkw.pop("numsubstep", None)
    fl = O.Flow(**flowkw)
    X, Y, Z = dims
    b = O.LBMBlock(X, Y, Z, dh=dh, BndConds=wl["bc"], iCollidModel=wl["model"], flow=fl)
    b.initialise(0.0)
    b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    bodies, plate = [], None
    if not with_bodies:
        pass
    elif wl["plate"] == "flex":
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


def bench_config(args, wl, world):
    """The `config` object both arms print (identical for the same command line)."""
    Xl, Y, Z = wl["dims"]
    return {"workload": args.workload, "desc": wl["desc"], "grid_per_gpu": [Xl, Y, Z], "grid_global": [Xl * world, Y, Z],
            "decomposition": f"x-slabs x{world}" if world > 1 else "single block",
            "l2": "working set >= 5.1 GB per GPU (two population buffers) >> 126 MB L2; no explicit flush needed"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on this box's host cores (the Fortran cannot be built here: no Fortran
    compiler in the image; the C restatement of its OpenMP path stands in, kind 'port'), with every core the process may use
    -- set here, because torchrun hands its workers OMP_NUM_THREADS=1.  MLUPS is a rate: each step is a bounded sample of the
    workload (one GPU's share of the global grid, around the body), and the value is what this host sustains on it whatever
    N is -- the driver's ratio at N GPUs is then N GPUs against this one host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    wl = WORKLOADS[args.workload]
    X, Y, Z = wl["dims"]
    cpu = host_cpu()
    # bounded sample: calibrate on a thin slab, then take the largest x extent that keeps the run in ~2 minutes
    mlups_cal, t_cal, threads = cpu_time_oracle(wl, (16, Y, Z), 2, 1, with_bodies=False)
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
    sample = (f"{xs}x{Y}x{Z} cells around the body out of the {X * world}x{Y}x{Z} grid, {args.steps} steps after {args.warmup} warm-up; "
              f"{threads} OpenMP threads on {cpu['physical']} physical / {cpu['logical']} logical cores of {cpu['model']}")
    line = {
        "metric": METRIC, "value": mlups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": bench_config(args, wl, world),
        "cpu_baseline": {"value": mlups, "unit": UNIT, "cores": threads, "physical_cores": cpu["physical"], "cpu_model": cpu["model"],
                         "kind": "port", "sample": sample,
                         "note": "C restatement of the reference OpenMP path (Fortran not buildable here); one host whatever N is"},
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
def parity_check(F, dist, rank, world, local, steps=10):
    """A small case of the same path run in THIS process group just before the timed region, checked against the CPU oracle on
    rank 0 (the oracle is the checker here, never the thing measured): inlet/outlet block with moving walls, a rigid plate whose
    stencil box straddles the interface of the middle ranks, cut into `world` x-slabs, `steps` steps.  Makes multi-rank parity
    visible in the driver's own run."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.common import perturbed_state, rel_err
    X, Y, Z = 10 * world + 1 if world > 1 else 21, 20, 24
    bc = (102, 104, 202, 202, 301, 301)
    kw = dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0), shearRateIn=(0.0, 3e-4, 0.0), Uref=0.04, ntolLBM=3, dtolLBM=1e-30)
    off, cnt = F.slab_range(X, rank, world)
    flow = F.FlowCondType(**kw)
    gb = F.LBMBlock(X, Y, Z, BndConds=bc, flow=flow, xOffset=off, xLocal=cnt, device=local)
    gb.initialise(0.0)
    f0 = perturbed_state((X, Y, Z), flow)
    gb.upload_fIn(np.ascontiguousarray(f0[:, off:off + cnt]))
    gb.update_volume_force(); gb.set_boundary_conditions()
    plate = F.RigidPlate(origin=(X / 2.0 - 4.2, 8.3, 5.2), nEL=8, len1=1.0, Nspan=8, spanlen=8.0, Lspan=0.0, chord_dir=(1.0, 0.2, 0.0), denIn=1.0)
    ob = ov = None
    if rank == 0:
        from oracle import oracle as O
        ob = O.LBMBlock(X, Y, Z, BndConds=bc, flow=O.Flow(**kw))
        ob.initialise(0.0)
        ob.fIn[...] = f0
        ob.update_volume_force(); ob.set_boundary_conditions(); ob.calculate_macro_quantities()
        ov = O.VirtualBody(plate.body.v_nelmts, v_move=0, iBodyModel=1)
        ov.v_Exyz[...] = plate.body.v_Exyz; ov.v_Evel[...] = plate.body.v_Evel; ov.v_Ea[...] = plate.body.v_Ea
    iters_equal, eF = True, 0.0
    for n in range(1, steps + 1):
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [plate], time=float(n), solver=False)
        if rank == 0:
            ob.set_blktime(float(n))
            it_o = ob.step([ov])
            iters_equal = iters_equal and it_o == it_g
            eF = max(eF, rel_err(plate.body.v_Eforce, ov.v_Eforce))
    den, uuu = gb.download_macro()
    floc = gb.download_fIn()
    parts = [(off, cnt, den, uuu, floc)]
    if world > 1:
        parts = [None] * world
        dist.gather_object((off, cnt, den, uuu, floc), parts if rank == 0 else None, dst=0)
    out = None
    if rank == 0:
        ob.calculate_macro_quantities()
        DEN = np.concatenate([p[2] for p in parts], axis=0)
        UUU = np.concatenate([p[3] for p in parts], axis=1)
        FF = np.concatenate([p[4] for p in parts], axis=1)
        out = {"case": f"{X}x{Y}x{Z}, BndConds {list(bc)}, rigid plate (64 markers) across the middle interface, {world} x-slab(s), {steps} steps, vs CPU oracle",
               "fIn_bit_exact": bool(np.array_equal(FF, ob.fIn)), "den_rel_err": rel_err(DEN, ob.den), "u_rel_err": rel_err(UUU, ob.uuu),
               "marker_force_rel_err": eF, "ibm_iterations_equal": bool(iters_equal)}
        out["ok"] = bool(out["fIn_bit_exact"] and out["den_rel_err"] <= 1e-12 and out["u_rel_err"] <= 1e-12 and eF <= 1e-10 and iters_equal)
    gb.close()
    if world > 1:
        dist.barrier()
    return out


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

    parity = None
    if not args.no_parity_check:
        try:
            parity = parity_check(F, dist, rank, world, local)
        except Exception as ex:   # reported, never hidden
            parity = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}
    early0 = int(F.lib().fsilbm_ibm_early_count())

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
    lu_local = cells_total / world          # lattice updates per step and GPU (son blocks count twice per root step)
    step_ms = ms / args.steps
    achieved = BYTES_PER_LU * lu_local / (step_ms * 1e-3) / 1e9
    alone = BYTES_PER_LU * cells_local / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            rec = json.load(open(prof)).get(args.workload, {})
            traffic, traffic_src = rec.get("dram_bytes_per_step"), rec.get("source")
        except Exception:
            traffic = None
    # frac is the WHOLE STEP (every launch of it: collide-stream in its x-ranges, face kernels, the IBM iteration, halo unpack,
    # and whatever host time the device had to wait for) against the 304 B per lattice update the collide-stream needs
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "scope": "whole step: algorithmic 304 B x lattice updates per GPU / ms_per_step",
                "algorithmic_bytes_per_step": BYTES_PER_LU * lu_local, "peak_source": peak_src,
                "collide_alone": {"kernel": "collide_push_kernel (IBM-free instantiation, launched back to back after the run)",
                                  "kernel_ms": kern_ms, "achieved": alone, "frac": alone / peak,
                                  "algorithmic_bytes_per_launch": BYTES_PER_LU * cells_local}}

    # ---- end to end through the host API with HOST buffers ------------------------------------------------
    # What the reference driver does around this path: populations come from pinned host memory once
    # (check_is_continue / initialise), then every step goes through the public API with host arguments (the
    # time scalar; with a body 7n marker doubles down and 3n force doubles up inside
    # fsilbm_ibm_interaction_force), and every `flow_every` steps the output staging array of write_flow_ (OUTtmp:
    # p,u,v,w as real(4), FluidDomain.f90:1640-1699; main.f90:130 timeFlowDelta cadence) is read back to pinned host
    # memory.  All copies are inside the timed region.
    flow_every = args.flow_every
    f_host = torch.empty((19, Xl, Y, Z), dtype=torch.float64, pin_memory=True)
    out_host = torch.empty((4, Xl, Y, Z), dtype=torch.float32, pin_memory=True)
    blk.download_fIn(f_host.numpy())
    # warm-up of the read-back path: its device staging buffer (0.5 GB for 33.5 M cells) is allocated on first use, and a
    # cudaMalloc of that size inside the timed region costs anything between 10 and 250 ms depending on the box
    blk.write_flow_window_async(out_host.numpy(), 0, 1)
    blk.download_wait()
    blk.sync(); torch.cuda.synchronize(); barrier()
    check = F._lib.check
    t0 = time.perf_counter()
    blk.upload_fIn(f_host.numpy())
    upload_s = time.perf_counter() - t0
    n_out = 0
    laps, lap_every = [], max(1, args.steps // 4)
    for n in range(args.steps):
        if n % lap_every == 0:
            laps.append(time.perf_counter() - t0)      # host clock when each quarter of the steps was ISSUED (diagnostic)
        step(args.warmup + args.steps + n + 1)
        if (n + 1) % flow_every == 0 or n + 1 == args.steps:
            # asynchronous read-back (the reference forks its writer): the copy overlaps the following steps; the previous
            # one is waited for before the host array is reused
            blk.write_flow_window_async(out_host.numpy(), 0, 1)
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
    d2h = out_host.numel() * 4 * world * n_out + 3 * 8 * markers * args.steps
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
           "seconds": e2e_s, "upload_seconds": upload_s, "quarter_issue_seconds": laps, "upload_gbs": f_host.numel() * 8 / upload_s / 1e9,
           "segment": f"fIn uploaded from pinned host once, {args.steps} steps through the LBMBlock API with host arguments, write_flow_'s staging array "
                      f"(p,u,v,w as real(4)) read back to pinned host every {flow_every} steps ({n_out} read-backs, asynchronous: each overlaps the "
                      f"following steps and is waited for before the next one and at the end); wall clock, bytes averaged per step"}

    # ---- CPU baseline on this box's cores (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            hc = host_cpu()
            sx = Xl if (not wl["plate"] or wl["plate"] == "flex") else 320
            steps_cpu = 12
            mlups, t_step, threads = cpu_time_oracle(wl, (sx, Y, Z), steps_cpu, 2)
            cpu = {"value": mlups, "unit": UNIT, "cores": threads, "physical_cores": hc["physical"], "cpu_model": hc["model"], "kind": "port",
                   "sample": f"{sx}x{Y}x{Z} cells around the body, {steps_cpu} steps after 2 warm-up ({t_step * 1e3:.0f} ms/step); "
                             "C restatement of the reference OpenMP path (Fortran not buildable here)"}
        except Exception as ex:   # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}

    if args.trace_out:
        blk.sync(); barrier()
        check(lib.fsilbm_set_option(b"trace", 1))
        for n in range(5):
            step(args.warmup + 2 * args.steps + n + 1)
        if sb is not None:
            sb.flush()
        blk.sync()
        check(lib.fsilbm_set_option(b"trace", 0))
        check(lib.fsilbm_trace_dump(f"{args.trace_out}.rank{rank}.csv".encode()))
        barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(args, wl, world),
            "details": {"halo_transport": transport, "structural_solver": structural,
                        "blocks": None if root is None else {"root_cells_per_gpu": Xl * Y * Z, "son_cells_on_rank0": son_cells, "son_updates_per_root_step": 2,
                                                             "note": "value counts the lattice updates of all blocks (root + 2 x son); the e2e "
                                                                     "transfers are the root block's"},
                        "ibm": ("ordered per-cell gather (bit-identical to the serial reference)" if args.ibm_ordered else "fp64 atomics") if wl["plate"] else None,
                        "ibm_calls_overlapped_with_update": int(lib.fsilbm_ibm_early_count()) - early0},
            "parity_check": parity,
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
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: plate512 (configs[2]) at --gpus 1, heave1024 (configs[3]) at --gpus N > 1")
    ap.add_argument("--trace-out", default=None, help="diagnostic: after the measurement, five more steps with the launch trace on; "
                                                       "every rank writes <trace-out>.rank<r>.csv")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the small oracle comparison run before the timed region")
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
    if args.workload is None:
        args.workload = default_workload(args.gpus)
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    import fsilbm3d_b200 as F
    F._lib.check(F.lib().fsilbm_set_option(b"halo", args.halo))
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_ordered", args.ibm_ordered))
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_single_launch", args.ibm_single_launch))
    for kv in args.opt:
        k, v = kv.split("=")
        F._lib.check(F.lib().fsilbm_set_option(k.encode(), int(v)))
    run_gpu(args)


if __name__ == "__main__":
    main()
