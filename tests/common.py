"""Shared helpers of the parity tests: build the same case on the oracle and on the CUDA path."""
import numpy as np

SEED = 20261017


def perturbed_state(shape_xyz, flow, rng_amp=1e-6, wave_amp=1e-3, seed=SEED):
    """f_eq(denIn, U + wave) plus a seeded uniform perturbation on every population."""
    X, Y, Z = shape_xyz
    ee = np.array([[0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0],
                   [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1],
                   [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]], dtype=float)
    wt = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)
    x = np.arange(X)[:, None, None]; y = np.arange(Y)[None, :, None]; z = np.arange(Z)[None, None, :]
    u = np.zeros((3, X, Y, Z))
    u[0] = flow.uvwIn[0] + wave_amp * np.sin(2 * np.pi * x / X) * np.sin(2 * np.pi * y / Y) * np.cos(2 * np.pi * z / Z)
    u[1] = flow.uvwIn[1] + 0.5 * wave_amp * np.cos(2 * np.pi * x / X) * np.sin(2 * np.pi * z / Z) + 0 * y
    u[2] = flow.uvwIn[2] + 0.25 * wave_amp * np.sin(2 * np.pi * y / Y) * np.cos(2 * np.pi * x / X) + 0 * z
    usq = (u ** 2).sum(0)
    f = np.empty((19, X, Y, Z))
    for q in range(19):
        eu = ee[0, q] * u[0] + ee[1, q] * u[1] + ee[2, q] * u[2]
        f[q] = wt[q] * flow.denIn * (1 + 3 * eu + 4.5 * eu * eu - 1.5 * usq)
    rng = np.random.default_rng(seed)
    f += rng.uniform(-rng_amp, rng_amp, size=f.shape)
    return np.ascontiguousarray(f)


def make_pair(O, F, dims, BndConds=(301,) * 6, model=1, params=(0.0,) * 10, dh=1.0, mins=(0.0, 0.0, 0.0), perturb=True, **flowkw):
    """Same block on both backends, after the start-up sequence of main.f90:50,62-64."""
    of = O.Flow(**flowkw)
    gf = F.FlowCondType(**flowkw)
    X, Y, Z = dims
    ob = O.LBMBlock(X, Y, Z, dh=dh, xmin=mins[0], ymin=mins[1], zmin=mins[2], BndConds=BndConds, iCollidModel=model, params=params, flow=of)
    gb = F.LBMBlock(X, Y, Z, dh=dh, xmin=mins[0], ymin=mins[1], zmin=mins[2], BndConds=BndConds, iCollidModel=model, params=params, flow=gf)
    ob.initialise(0.0)
    gb.initialise(0.0)
    if perturb:
        f0 = perturbed_state(dims, of)
        ob.fIn[...] = f0
        gb.upload_fIn(f0)
    ob.update_volume_force(); ob.set_boundary_conditions(); ob.calculate_macro_quantities()
    gb.update_volume_force(); gb.set_boundary_conditions()
    return ob, gb


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
