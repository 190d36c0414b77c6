"""Output / restart / diagnostics of the fluid blocks, computed from the device state and written in the reference's
byte formats (FluidDomain.f90:1628-1737 Flow files, :268-285,1770-1789 continue files, :128-237 restart with
trilinear re-gridding, :1147-1172 turbulence averages, :2019-2056 flux, FlowCondition.f90:195-222 probes)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fsilbm3d_b200 as F
    return F


def _pair(oracle, F, dims=(14, 12, 18), bc=(101, 104, 203, 203, 301, 301), steps=6, **kw):
    from tests.common import make_pair
    flow = dict(nu=0.03, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, volumeForceIn=(1e-6, 0.0, 0.0))
    flow.update(kw)
    ob, gb = make_pair(oracle, F, dims, BndConds=bc, dh=0.5, mins=(-1.0, 0.25, 2.0), **flow)
    for n in range(1, steps + 1):
        ob.set_blktime(0.5 * n); gb.set_blktime(0.5 * n)
        ob.step(); gb.step()
    ob.calculate_macro_quantities()   # main.f90:107
    return ob, gb


@pytest.mark.parametrize("offset", [0, 2])
def test_flow_file_bytes(oracle, F, tmp_path, offset):
    ob, gb = _pair(oracle, F)
    Tref, time, ID = 25.0, 3.0, 7
    path = F.flow_io.write_flow(gb, time, Tref, ID=ID, offsetOutput=offset, outputtype=1, root=str(tmp_path))
    assert os.path.basename(path) == "Flow0000012000_b007"            # nint(3/25*1e5) = 12000
    o = offset
    sl = (slice(o, ob.xDim - o), slice(o, ob.yDim - o), slice(o, ob.zDim - o))
    invUref = 1.0 / ob.flow.Uref
    exp = np.array([ob.xDim - 2 * o, ob.yDim - 2 * o, ob.zDim - 2 * o, ID], dtype=np.int32).tobytes()
    exp += np.array([ob.xmin + o * ob.dh, ob.ymin + o * ob.dh, ob.zmin + o * ob.dh, ob.dh], dtype=np.float64).tobytes()
    exp += ((1.0 / 3.0) * (ob.den[sl] - ob.flow.denIn)).astype(np.float32).tobytes()               # FluidDomain.f90:1658
    for k in range(3):
        exp += (ob.uuu[(k,) + sl] * invUref).astype(np.float32).tobytes()                          # :1659-1661
    assert open(path, "rb").read() == exp
    dims, geo, arr = F.flow_io.read_flow(path)
    assert tuple(dims) == (ob.xDim - 2 * o, ob.yDim - 2 * o, ob.zDim - 2 * o, ID) and arr.shape[0] == 4
    gb.close()


def test_continue_round_trip_and_regrid(oracle, F, tmp_path):
    """write_continue -> check_is_continue on the same grid is exact (coefficients 0/1), and a run split by a restart
    equals the uninterrupted run bit for bit; a shifted, coarser block is re-gridded trilinearly."""
    from tests.common import rel_err
    ob, gb = _pair(oracle, F, steps=5)
    path = F.flow_io.write_continue_blocks([gb], step=5, time_over_Tref=0.1, root=str(tmp_path))
    assert os.path.basename(path) == "continue0000010000"
    raw = open(path, "rb").read()
    assert len(raw) == 16 + 32 + 12 + 19 * ob.fIn[0].size * 8
    assert np.frombuffer(raw, np.int32, 2).tolist() == [1, 5] and np.frombuffer(raw, np.float64, 1, 8)[0] == 0.1
    assert np.array_equal(np.frombuffer(raw, np.float64, ob.fIn.size, 60).reshape(ob.fIn.shape), ob.fIn)
    os.rename(path, os.path.join(str(tmp_path), "DatContinue", "continue"))
    # restart into a fresh block of the same geometry
    g2 = F.LBMBlock(ob.xDim, ob.yDim, ob.zDim, dh=ob.dh, xmin=ob.xmin, ymin=ob.ymin, zmin=ob.zmin, BndConds=ob.BndConds,
                    flow=F.FlowCondType(nu=0.03, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, volumeForceIn=(1e-6, 0.0, 0.0)))
    g2.initialise(0.0)
    step, t = F.flow_io.check_is_continue([g2], 1, root=str(tmp_path))
    assert (step, t) == (5, 0.1)
    assert np.array_equal(g2.download_fIn(), ob.fIn)
    g2.update_volume_force(); g2.set_boundary_conditions()      # main.f90:62-63 after the restart
    ob.update_volume_force(); ob.set_boundary_conditions()
    # half-way walls: the first boundary call of a run only allocates the stash (:660-661) -- a restarted oracle run does the same
    o2 = oracle.LBMBlock(ob.xDim, ob.yDim, ob.zDim, dh=ob.dh, xmin=ob.xmin, ymin=ob.ymin, zmin=ob.zmin, BndConds=ob.BndConds, flow=ob.flow)
    o2.initialise(0.0); o2.fIn[...] = ob.fIn
    o2.update_volume_force(); o2.set_boundary_conditions(); o2.calculate_macro_quantities()
    for n in range(6, 10):
        o2.set_blktime(0.5 * n); g2.set_blktime(0.5 * n)
        o2.step(); g2.step()
    assert np.array_equal(g2.download_fIn(), o2.fIn)
    # re-grid onto a finer block inside the saved one: compare with a direct trilinear evaluation
    g3 = F.LBMBlock(9, 7, 11, dh=0.25, xmin=0.0, ymin=1.0, zmin=3.0, BndConds=(0,) * 6, flow=g2.flow)
    g3.initialise(0.0)
    F.flow_io.check_is_continue([g3], 1, root=str(tmp_path))
    f3 = g3.download_fIn()
    x, y, z = 4, 3, 6
    P = np.array([0.0 + x * 0.25, 1.0 + y * 0.25, 3.0 + z * 0.25])
    co = (P - np.array([ob.xmin, ob.ymin, ob.zmin])) / ob.dh
    i0 = np.floor(co).astype(int); c = co - i0
    want = np.zeros(19)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (c[0] if dx else 1 - c[0]) * (c[1] if dy else 1 - c[1]) * (c[2] if dz else 1 - c[2])
                want += w * ob.fIn[:, i0[0] + dx, i0[1] + dy, i0[2] + dz]
    assert rel_err(f3[:, x, y, z], want) < 1e-14
    for b in (gb, g2, g3):
        b.close()


def test_turbulent_statistic_and_mean_flow_file(oracle, F, tmp_path):
    ob, gb = _pair(oracle, F, steps=2)
    X, Y, Z = ob.xDim, ob.yDim, ob.zDim
    ave = np.zeros((9, X, Y, Z))
    start = 3
    for step in range(3, 9):
        ob.set_blktime(0.5 * step); gb.set_blktime(0.5 * step)
        ob.step(); gb.step()
        ob.calculate_macro_quantities()
        gb.calculate_turbulent_statistic(step, start)
        invStep = float(np.float32(1) / np.float32(step - start + 1))        # real(4) quirk, FluidDomain.f90:1153
        u = ob.uuu
        for k in range(3):
            ave[k] = ave[k] * (1.0 - invStep) + invStep * u[k]
        for k in range(3):
            ave[3 + k] = ave[3 + k] * (1.0 - invStep) + invStep * (u[k] - ave[k]) * (u[k] - ave[k])
        for n, (a, b) in enumerate(((0, 1), (0, 2), (1, 2))):
            ave[6 + n] = ave[6 + n] * (1.0 - invStep) + invStep * (u[a] - ave[a]) * (u[b] - ave[b])
    F.flow_io.write_flow(gb, 1.0, 1.0, ID=1, offsetOutput=1, outputtype=3, root=str(tmp_path))
    raw = open(os.path.join(str(tmp_path), "DatFlow", "MeanFlow_b001"), "rb").read()
    sl = (slice(1, X - 1), slice(1, Y - 1), slice(1, Z - 1))
    n = (X - 2) * (Y - 2) * (Z - 2)
    got = np.frombuffer(raw, np.float32, 10 * n, 48).reshape(10, X - 2, Y - 2, Z - 2)
    invU = 1.0 / ob.flow.Uref
    assert np.array_equal(got[0], ((1.0 / 3.0) * (ob.den[sl] - 1.0)).astype(np.float32))
    for k in range(3):
        assert np.array_equal(got[1 + k], (ave[(k,) + sl] * invU).astype(np.float32))
    for k in range(3, 9):
        assert np.array_equal(got[1 + k], (ave[(k,) + sl] * (1.0 / ob.flow.Uref / ob.flow.Uref)).astype(np.float32))
    assert os.path.exists(os.path.join(str(tmp_path), "DatFlow", "Flow0000100000_b001"))
    gb.close()


def test_flux_probes_fieldstat(oracle, F, tmp_path):
    from tests.common import rel_err
    ob, gb = _pair(oracle, F)
    X, Y, Z, dh = ob.xDim, ob.yDim, ob.zDim, ob.dh
    # write_fluid_flux, FluidDomain.f90:2027-2054
    wy = np.ones(Y); wy[[0, -1]] = 0.5
    wz = np.ones(Z); wz[[0, -1]] = 0.5
    W = wy[:, None] * wz[None, :]
    raw = [float(np.sum(ob.uuu[0, x] * ob.den[x] * dh * dh * W)) for x in (0, (X + 1) // 2 - 1, X - 1)]
    Yref = dh * (Y - 1)                      # ymax - ymin (walls)
    Zref = dh * (Z - 1) + dh                 # periodic z: zmax carries the extra dh (FluidDomain.f90:103-105)
    want = np.array(raw) / (1.0 * ob.flow.Uref * Zref * Yref)
    got = F.flow_io.write_fluid_flux(gb, 3.0, 25.0, 1.0, ob.flow.Uref, root=str(tmp_path))
    assert rel_err(got, want) < 1e-12
    line = open(os.path.join(str(tmp_path), "DatInfo", "FluidFlux.dat")).read().splitlines()[0]
    assert len(line) == 80 and line.startswith("    0.1200000000E+00")
    # probes: grid_value_interpolation, Util.f90:123-157
    coords = np.array([[0.3, 1.7, 4.1], [ob.xmin, ob.ymin, ob.zmin], [ob.xmin + dh * (X - 1), ob.ymin + dh * (Y - 1), ob.zmin + dh * (Z - 1)]])
    vel = F.flow_io.write_fluid_information(gb, 3.0, 25.0, ob.flow.Uref, coords, root=str(tmp_path))
    for P, v in zip(coords, vel):
        co = (P - np.array([ob.xmin, ob.ymin, ob.zmin])) / dh
        i0 = np.minimum(np.floor(co).astype(int), [X - 1, Y - 1, Z - 1]); c = co - i0
        want = np.zeros(3)
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (c[0] if dx else 1 - c[0]) * (c[1] if dy else 1 - c[1]) * (c[2] if dz else 1 - c[2])
                    if w != 0.0:
                        want += w * ob.uuu[:, i0[0] + dx, i0[1] + dy, i0[2] + dz]
        assert rel_err(v, want) < 1e-12
    assert os.path.exists(os.path.join(str(tmp_path), "DatInfo", "FluidProbes_0003.dat"))
    with pytest.raises(F.FsilbmError):
        F.flow_io.write_fluid_information(gb, 3.0, 25.0, ob.flow.Uref, [[99.0, 0.0, 0.0]], root=str(tmp_path))
    # FIELDSTAT lines
    txt = F.flow_io.fieldstat_lines(gb)
    st = ob.ComputeFieldStat()
    assert txt.splitlines()[0] == f" FIELDSTAT L2 u {st[0]:18.12f}" and len(txt.splitlines()) == 6
    gb.close()


def test_async_macro_readback(oracle, F):
    """fsilbm_block_download_macro_async: den/uuu of the state at the call, copied out while later steps run."""
    import torch
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (24, 20, 28), BndConds=(101, 104, 203, 203, 301, 301), nu=0.05, uvwIn=(0.04, 0.0, 0.0), Uref=0.04,
                       volumeForceIn=(1e-6, 0.0, 0.0))
    for n in range(1, 6):
        ob.set_blktime(float(n)); ob.step([])
        F.tree_collision_streaming_IBM_FEM(gb, [], time=float(n))
    ob.calculate_macro_quantities()
    den = torch.empty(gb.shape, dtype=torch.float64, pin_memory=True)
    uuu = torch.empty((3,) + gb.shape, dtype=torch.float64, pin_memory=True)
    den.fill_(-1.0); uuu.fill_(-1.0)
    gb.download_macro_async(den.numpy(), uuu.numpy())
    for n in range(6, 10):                       # the update goes on while the copy is in flight
        F.tree_collision_streaming_IBM_FEM(gb, [], time=float(n))
    gb.download_wait()
    assert np.array_equal(den.numpy(), ob.den) and np.array_equal(uuu.numpy(), ob.uuu)
    # a second read-back waits for the first; the synchronous call still works in between
    gb.download_macro_async(den.numpy(), uuu.numpy())
    d2, u2 = gb.download_macro()
    gb.download_wait()
    assert np.array_equal(den.numpy(), d2) and np.array_equal(uuu.numpy(), u2)
    for n in range(6, 10):
        ob.set_blktime(float(n)); ob.step([])
    ob.calculate_macro_quantities()
    assert np.array_equal(d2, ob.den) and np.array_equal(u2, ob.uuu)
    gb.close()


def test_async_flow_window_readback(oracle, F):
    """fsilbm_block_write_flow_window_async: OUTtmp of write_flow_ (real(4) p,u,v,w over the output window) of the state at the call,
    copied out while later steps run; identical to the synchronous call's bytes."""
    import torch
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (24, 20, 28), BndConds=(101, 104, 203, 203, 301, 301), nu=0.05, uvwIn=(0.04, 0.0, 0.0), Uref=0.04,
                       volumeForceIn=(1e-6, 0.0, 0.0))
    for n in range(1, 6):
        F.tree_collision_streaming_IBM_FEM(gb, [], time=float(n))
    off = 2
    _, nx, ny, nz = F.flow_io.flow_window(gb, off)
    want = np.empty((4, nx, ny, nz), dtype=np.float32)
    F._lib.check(F.lib().fsilbm_block_write_flow_window(gb._h, off, 1, want.ctypes.data))
    out = torch.empty((4, nx, ny, nz), dtype=torch.float32, pin_memory=True)
    out.fill_(-7.0)
    gb.write_flow_window_async(out.numpy(), off, 1)
    for n in range(6, 10):                       # the update goes on while the copy is in flight
        F.tree_collision_streaming_IBM_FEM(gb, [], time=float(n))
    gb.download_wait()
    assert np.array_equal(out.numpy(), want)
    # ... and it is what the reference stores: den/uuu of the oracle at that step, scaled and cast as FluidDomain.f90:1650-1660 does
    for n in range(1, 6):
        ob.set_blktime(float(n)); ob.step([])
    ob.calculate_macro_quantities()
    sl = (slice(off, 24 - off), slice(off, 20 - off), slice(off, 28 - off))
    invUref = 1.0 / ob.flow.Uref
    assert np.array_equal(out.numpy()[1], (ob.uuu[0][sl] * invUref).astype(np.float32))
    # a second asynchronous read-back waits for the first (one staging buffer)
    gb.write_flow_window_async(out.numpy(), off, 1)
    gb.write_flow_window_async(out.numpy(), off, 1)
    gb.download_wait()
    gb.close()
