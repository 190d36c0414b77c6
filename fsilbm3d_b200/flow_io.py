"""Host side of the reference's fluid output / restart procedures in its own byte formats, fed from the device state
through the C ABI (fsilbm_block_write_flow_window, ..._download_fIn, ..._fluid_flux, ..._probe_velocity).

  write_flow_blocks      FluidDomain.f90:349-366 -> write_flow_ :1628-1737      ./DatFlow/Flow<10d>_b<3d>, MeanFlow_b<3d>
  write_continue_blocks  :268-285 -> write_continue_ :1770-1777                 ./DatContinue/continue<10d>
  check_is_continue      :128-237 -> read_continue_ :1779-1789                  ./DatContinue/continue
  write_fluid_flux       :2019-2056                                             ./DatInfo/FluidFlux.dat
  write_fluid_information FlowCondition.f90:195-222                             ./DatInfo/FluidProbes_<4d>.dat
  computeFieldStat_blocks FluidDomain.f90:368-374 -> ComputeFieldStat_ :1739-1768   FIELDSTAT lines on stdout

All files are Fortran stream-unformatted (no record markers), native little-endian.  The reference forks a child to
write the flow file (:1702); here the staging buffer is ordinary host memory, so a caller may hand it to a thread.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np

from ._lib import check, lib


def _name10(value: float) -> str:
    """write(fileName,'(I10)') nint(value*1d5), right-adjusted, blanks -> '0' (FluidDomain.f90:274-278)."""
    n = int(math.floor(value * 1e5 + 0.5)) if value >= 0 else -int(math.floor(-value * 1e5 + 0.5))   # Fortran NINT
    return f"{n:10d}".replace(" ", "0")


def _whole(block, what: str) -> None:
    """The restart helpers hold a block's full x extent: an x-slab of a multi-GPU run has to be gathered by the caller first."""
    if block.xLocal != block.xDim or getattr(block, "xOffset", 0) != 0:
        raise ValueError(f"{what}: block is an x-slab (planes {block.xOffset}..{block.xOffset + block.xLocal - 1} of {block.xDim}); "
                         "gather the slabs (download_fIn per rank, concatenated along x) or write one file per rank with write_flow")


def flow_window(block, offsetOutput: int):
    """The part of the output window [offsetOutput, dim - offsetOutput) that lies in this block's slab: (first global plane, nx, ny,
    nz) -- the same intersection fsilbm_block_write_flow_window takes, so the staging buffer always has the size the library fills."""
    x0 = getattr(block, "xOffset", 0)
    gx0, gx1 = max(offsetOutput, x0), min(block.xDim - offsetOutput, x0 + block.xLocal)
    return gx0, max(gx1 - gx0, 0), block.yDim - 2 * offsetOutput, block.zDim - 2 * offsetOutput


def write_flow(block, time: float, Tref: float, ID: int = 1, offsetOutput: int = 0, outputtype: int = 1, root: str = ".") -> Optional[str]:
    """write_flow_ for one block.  An x-slab writes ITS planes of the window (header: its own nx and xmin), so ranks of a slab run
    need distinct `root` directories or IDs.  Returns the path of the Flow file (None if outputtype < 1 or the window misses the slab)."""
    if outputtype < 1:
        return None
    gx0, nx, ny, nz = flow_window(block, offsetOutput)
    if nx <= 0 or ny <= 0 or nz <= 0:
        return None
    nf = 13 if outputtype >= 2 else 4
    out = np.empty((nf, nx, ny, nz), dtype=np.float32)
    check(lib().fsilbm_block_write_flow_window(block._h, offsetOutput, outputtype, out.ctypes.data))
    os.makedirs(os.path.join(root, "DatFlow"), exist_ok=True)
    head = np.array([nx, ny, nz, ID], dtype=np.int32).tobytes() + \
        np.array([block.xmin + gx0 * block.dh, block.ymin + offsetOutput * block.dh, block.zmin + offsetOutput * block.dh, block.dh],
                 dtype=np.float64).tobytes()
    path = None
    if outputtype != 2:
        path = os.path.join(root, "DatFlow", f"Flow{_name10(time / Tref)}_b{ID:3d}".replace(" ", "0"))
        with open(path, "wb") as fh:
            fh.write(head)
            fh.write(out[0:4].tobytes())
    if outputtype >= 2:
        mpath = os.path.join(root, "DatFlow", f"MeanFlow_b{ID:3d}".replace(" ", "0"))
        with open(mpath, "wb") as fh:
            fh.write(head)
            fh.write(out[0].tobytes())
            fh.write(out[4:13].tobytes())
        path = path or mpath
    return path


def read_flow(path: str):
    """Inverse of write_flow for the 4-field file: (nx,ny,nz,ID), (xmin,ymin,zmin,dh), p, u, v, w as float32 [x][y][z]."""
    raw = open(path, "rb").read()
    dims = np.frombuffer(raw, dtype=np.int32, count=4)
    geo = np.frombuffer(raw, dtype=np.float64, count=4, offset=16)
    nx, ny, nz = (int(v) for v in dims[:3])
    arr = np.frombuffer(raw, dtype=np.float32, count=4 * nx * ny * nz, offset=48).reshape(4, nx, ny, nz)
    return dims, geo, arr


def write_continue_blocks(blocks: Sequence, step: int, time_over_Tref: float, root: str = ".") -> str:
    """write_continue_blocks (FluidDomain.f90:268-285).  `time_over_Tref` is what main.f90:118 passes (time / Tref)."""
    for b in blocks:
        _whole(b, "write_continue_blocks")
    os.makedirs(os.path.join(root, "DatContinue"), exist_ok=True)
    path = os.path.join(root, "DatContinue", "continue" + _name10(time_over_Tref))
    with open(path, "wb") as fh:
        fh.write(np.array([len(blocks), step], dtype=np.int32).tobytes())
        fh.write(np.array([time_over_Tref], dtype=np.float64).tobytes())
        for b in blocks:
            fh.write(np.array([b.xmin, b.ymin, b.zmin, b.dh], dtype=np.float64).tobytes())
            fh.write(np.array([b.xDim, b.yDim, b.zDim], dtype=np.int32).tobytes())
            fh.write(np.ascontiguousarray(b.download_fIn()).tobytes())
    return path


def read_continue_file(path: str):
    """read(idfile) nblocks,step,time then read_continue_ per block (FluidDomain.f90:146-151,1779-1789)."""
    raw = open(path, "rb").read()
    nblocks, step = (int(v) for v in np.frombuffer(raw, dtype=np.int32, count=2))
    time = float(np.frombuffer(raw, dtype=np.float64, count=1, offset=8)[0])
    off = 16
    out = []
    for _ in range(nblocks):
        geo = np.frombuffer(raw, dtype=np.float64, count=4, offset=off); off += 32
        dims = np.frombuffer(raw, dtype=np.int32, count=3, offset=off); off += 12
        X, Y, Z = (int(v) for v in dims)
        f = np.frombuffer(raw, dtype=np.float64, count=19 * X * Y * Z, offset=off).reshape(19, X, Y, Z); off += 19 * X * Y * Z * 8
        out.append(dict(xmin=geo[0], ymin=geo[1], zmin=geo[2], dh=geo[3], xDim=X, yDim=Y, zDim=Z, fIn=f))
    return nblocks, step, time, out


def regrid_from_continue(block, saved) -> Optional[np.ndarray]:
    """The trilinear re-gridding of check_is_continue (FluidDomain.f90:166-224) for one current block: every node takes
    its populations from the finest saved block that contains it; nodes no saved block covers keep their value.
    Returns the new fIn (19,X,Y,Z) or None when nothing is covered."""
    _whole(block, "regrid_from_continue")
    order = list(range(len(saved)))                                            # sortdh, :153-165: as written there the swap
    for i in range(len(saved) - 1):                                            # test compares the UNSORTED dh of slots i and j
        for j in range(i + 1, len(saved)):
            if saved[i]["dh"] > saved[j]["dh"]:
                order[i], order[j] = order[j], order[i]
    f = np.array(block.download_fIn())
    X, Y, Z = block.xDim, block.yDim, block.zDim
    xC = block.xmin + np.arange(X) * block.dh
    yC = block.ymin + np.arange(Y) * block.dh
    zC = block.zmin + np.arange(Z) * block.dh
    todo = np.ones((X, Y, Z), dtype=bool)
    touched = False
    for j in order:
        s = saved[j]
        mx = [s["xmin"] + (s["xDim"] - 1) * s["dh"], s["ymin"] + (s["yDim"] - 1) * s["dh"], s["zmin"] + (s["zDim"] - 1) * s["dh"]]
        inx = (xC >= s["xmin"]) & (xC <= mx[0]); iny = (yC >= s["ymin"]) & (yC <= mx[1]); inz = (zC >= s["zmin"]) & (zC <= mx[2])
        sel = todo & inx[:, None, None] & iny[None, :, None] & inz[None, None, :]
        if not sel.any():
            continue
        touched = True

        def axis(c, cmin, dim):
            co = (c - cmin) / s["dh"]
            i0 = np.floor(co).astype(int)
            i0 = np.where(i0 == dim - 1, i0 - 1, i0)
            i0 = np.clip(i0, 0, max(dim - 2, 0))       # nodes outside the saved block are masked out by `sel` anyway
            return i0, co - i0
        ix, cx = axis(xC, s["xmin"], s["xDim"]); iy, cy = axis(yC, s["ymin"], s["yDim"]); iz, cz = axis(zC, s["zmin"], s["zDim"])
        IX, IY, IZ = ix[:, None, None], iy[None, :, None], iz[None, None, :]
        CX, CY, CZ = cx[:, None, None], cy[None, :, None], cz[None, None, :]
        g = s["fIn"]
        # the eight terms in the order of :197-204
        new = (g[:, IX, IY, IZ] * (1 - CZ) * (1 - CY) * (1 - CX) + g[:, IX + 1, IY, IZ] * (1 - CZ) * (1 - CY) * CX +
               g[:, IX, IY + 1, IZ] * (1 - CZ) * CY * (1 - CX) + g[:, IX + 1, IY + 1, IZ] * (1 - CZ) * CY * CX +
               g[:, IX, IY, IZ + 1] * CZ * (1 - CY) * (1 - CX) + g[:, IX + 1, IY, IZ + 1] * CZ * (1 - CY) * CX +
               g[:, IX, IY + 1, IZ + 1] * CZ * CY * (1 - CX) + g[:, IX + 1, IY + 1, IZ + 1] * CZ * CY * CX)
        f[:, sel] = new[:, sel]
        todo &= ~sel
    return f if touched else None


def check_is_continue(blocks: Sequence, isContinue: int, root: str = "."):
    """check_is_continue (FluidDomain.f90:128-237).  Returns (step, time) from the file, or (0, 0.0) for a new run."""
    path = os.path.join(root, "DatContinue", "continue")
    if isContinue >= 1 and os.path.exists(path):
        _, step, time, saved = read_continue_file(path)
        for b in blocks:
            f = regrid_from_continue(b, saved)
            if f is not None:
                b.upload_fIn(f)
        return step, time
    return 0, 0.0


def write_fluid_flux(block, time: float, Tref: float, denIn: float, Uref: float, root: str = ".") -> np.ndarray:
    """write_fluid_flux (FluidDomain.f90:2019-2056): appends time/Tref and the three normalised fluxes to DatInfo/FluidFlux.dat."""
    raw = (C.c_double * 3)()
    check(lib().fsilbm_block_fluid_flux(block._h, raw))
    ymax = block.ymin + block.dh * (block.yDim - 1) + (block.dh if block.BndConds[2] == 301 and block.BndConds[3] == 301 else 0.0)
    zmax = block.zmin + block.dh * (block.zDim - 1) + (block.dh if block.BndConds[4] == 301 and block.BndConds[5] == 301 else 0.0)
    Yref, Zref = ymax - block.ymin, zmax - block.zmin
    vals = np.array(raw[:]) / (denIn * Uref * Zref * Yref)
    os.makedirs(os.path.join(root, "DatInfo"), exist_ok=True)
    with open(os.path.join(root, "DatInfo", "FluidFlux.dat"), "a") as fh:
        fh.write("".join(_e20_10(v) for v in (time / Tref, *vals)) + "\n")
    return vals


def write_fluid_information(block, time: float, Tref: float, Uref: float, coords, root: str = ".") -> np.ndarray:
    """write_fluid_information (FlowCondition.f90:195-222): one DatInfo/FluidProbes_<4d>.dat line per probe."""
    co = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
    vel = np.empty_like(co)
    check(lib().fsilbm_block_probe_velocity(block._h, len(co), co.ctypes.data, vel.ctypes.data))
    os.makedirs(os.path.join(root, "DatInfo"), exist_ok=True)
    for i, v in enumerate(vel):
        with open(os.path.join(root, "DatInfo", f"FluidProbes_{i + 1:04d}.dat"), "a") as fh:
            fh.write("".join(_e20_10(x) for x in (time / Tref, *(v / Uref))) + "\n")
    return vel


def _e20_10(v: float) -> str:
    """Fortran edit descriptor E20.10: 0.dddddddddde+XX right-adjusted in 20 columns, ten significant digits correctly rounded
    from the binary value (as the Fortran run-time library does)."""
    if v == 0.0:
        return "    0.0000000000E+00"
    mant, e10 = f"{abs(v):.9e}".split("e")
    s = f"{'-' if v < 0 else ''}0.{mant.replace('.', '')}E{int(e10) + 1:+03d}"
    return s.rjust(20)


def fieldstat_lines(block) -> str:
    """The six FIELDSTAT lines of ComputeFieldStat_ (FluidDomain.f90:1762-1767), format (A,F18.12)."""
    st = block.ComputeFieldStat()
    names = ["L2 u", "L2 v", "L2 w", "Linfinity u", "Linfinity v", "Linfinity w"]
    return "\n".join(f" FIELDSTAT {n} {v:18.12f}" for n, v in zip(names, st))
