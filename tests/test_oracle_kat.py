"""Known-answer tests that pin the CPU oracle (the reference ships no tests or vectors).
Each is an analytic property of the reference's algorithm, not a recorded number."""
import numpy as np
import pytest


def start(b):
    b.initialise(0.0)
    b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()


def test_d3q19_tables(oracle):
    # lattice isotropy: sum w = 1, sum w e = 0, sum w e e = Cs2 I  (ConstParams.f90:11-25,39)
    ee = np.array([[0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0],
                   [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1],
                   [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]], dtype=float)
    b = oracle.LBMBlock(4, 4, 4, flow=oracle.Flow(uvwIn=(0.03, -0.02, 0.01)))
    b.initialise(0.0)
    f = b.fIn[:, 0, 0, 0]
    assert abs(f.sum() - 1.0) < 1e-15
    np.testing.assert_allclose(ee @ f, [0.03, -0.02, 0.01], atol=1e-16)


@pytest.mark.parametrize("model", [1, 2, 3])
def test_uniform_equilibrium_is_a_fixed_point(oracle, model):
    b = oracle.LBMBlock(6, 5, 7, iCollidModel=model, params=(0.25,) + (0.0,) * 9, flow=oracle.Flow(uvwIn=(0.05, 0.02, -0.01)))
    start(b)
    f0 = b.fIn.copy()
    for _ in range(5):
        b.step()
    assert np.max(np.abs(b.fIn - f0)) < 5e-16


@pytest.mark.parametrize("model", [1, 2, 3])
def test_uniform_force_momentum(oracle, model):
    # macro u at the start of step n+1 equals (n+1/2) F dh / rho  (SURVEY 8c pin 2)
    F = 1e-6
    b = oracle.LBMBlock(6, 4, 5, iCollidModel=model, params=(3 / 16,) + (0.0,) * 9, flow=oracle.Flow(volumeForceIn=(F, 0, 0)))
    start(b)
    n = 25
    for _ in range(n):
        b.step()
    b.calculate_macro_quantities()
    np.testing.assert_allclose(b.uuu[0], (n + 0.5) * F, rtol=1e-10)
    np.testing.assert_allclose(b.den, 1.0, rtol=1e-13)


# Exactly conservative rules only: periodic wrap and the half-way wall (which returns precisely the
# populations that left, FluidDomain.f90:571-577,665).  The full-way wall 201 and the mirror 302 copy
# post-stream neighbours (:655,697) and conserve mass only in the steady state.
@pytest.mark.parametrize("bc", [(301,) * 6, (301, 301, 203, 203, 301, 301), (203,) * 6, (203, 203, 301, 301, 203, 203)])
def test_mass_is_conserved(oracle, bc):
    from tests.common import perturbed_state
    fl = oracle.Flow(nu=0.05)
    b = oracle.LBMBlock(8, 7, 6, BndConds=bc, flow=fl)
    b.initialise(0.0)
    b.fIn[...] = perturbed_state((8, 7, 6), fl)
    b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    b.step()  # the first step after the skipped half-way start-up call settles the stash
    m0 = b.fIn.sum()
    for _ in range(20):
        b.step()
    assert abs(b.fIn.sum() - m0) / m0 < 1e-13


def test_streaming_returns_after_a_period(oracle):
    b = oracle.LBMBlock(5, 4, 3)
    b.initialise(0.0)
    rng = np.random.default_rng(1)
    b.fIn[...] = rng.uniform(0, 1, b.fIn.shape)
    f0 = b.fIn.copy()
    for _ in range(5 * 4 * 3):
        b.streaming()
    assert np.array_equal(b.fIn, f0)
    b.streaming()
    # population 1 moved one plane in +x: f_new(x) = f(x-1)
    assert np.array_equal(b.fIn[1], np.roll(f0[1], 1, axis=0))
    assert np.array_equal(b.fIn[18], np.roll(np.roll(f0[18], -1, axis=1), -1, axis=2))


def test_streaming_independent_of_thread_partition(oracle):
    rng = np.random.default_rng(2)
    f0 = rng.uniform(0, 1, (19, 9, 4, 5))
    outs = []
    for nps in (1, 2, 4):
        b = oracle.LBMBlock(9, 4, 5, npsize=nps)
        b.initialise(0.0)
        b.fIn[...] = f0
        b.streaming()
        outs.append(b.fIn.copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_poiseuille_trt_magic_parameter(oracle):
    # force-driven flow between half-way walls; TRT with Lambda = 3/16 puts the wall exactly half a
    # cell outside the boundary nodes and the steady profile is the exact parabola (SURVEY 8c pin 4)
    F, nu, H = 1e-6, 0.1, 9
    b = oracle.LBMBlock(3, H, 3, BndConds=(301, 301, 203, 203, 301, 301), iCollidModel=2, params=(3 / 16,) + (0.0,) * 9,
                        flow=oracle.Flow(nu=nu, volumeForceIn=(F, 0, 0)))
    start(b)
    for _ in range(4000):
        b.step()
    b.calculate_macro_quantities()
    y = np.arange(H) + 0.5
    ua = F / (2 * nu) * y * (H - y)
    assert np.max(np.abs(b.uuu[0, 1, :, 1] - ua)) / ua.max() < 1e-8


def test_delta_function_partition_of_unity(oracle):
    # sum_{-1..2} Phi(r - delta) = 1 for any delta (Solidbody.f90:822-833); first moment vanishes
    for d in np.linspace(0, 0.999, 37):
        w = np.array([oracle.Phi(k - d) for k in (-1, 0, 1, 2)])
        assert abs(w.sum() - 1.0) < 1e-15
        assert abs((w * (np.array([-1, 0, 1, 2]) - d)).sum()) < 1e-15
        assert abs(np.float32(w).astype(float).sum() - 1.0) < 2e-7   # as stored in real(4), Solidbody.f90:802-804


def _plate_body(oracle, n=6, m=5, origin=(5.3, 6.1, 4.7), h=1.0):
    body = oracle.VirtualBody(n * m, v_move=0, iBodyModel=1)
    k = 0
    for i in range(n):
        for j in range(m):
            body.v_Exyz[k] = (origin[0] + 0.5 * h * i, origin[1] + 0.1 * i, origin[2] + h * j)
            body.v_Evel[k] = (0.0, 0.0, 0.0)
            body.v_Ea[k] = -2.0 * 1.0 * h * h * 0.5
            k += 1
    return body


def test_ibm_force_conservation_and_no_slip(oracle):
    # sum_cells force dh^3 = - sum_markers v_Eforce (Solidbody.f90:968-976), and the penalty iteration
    # drives the interpolated velocity towards the marker velocity (Solidbody.f90:895-906)
    fl = oracle.Flow(nu=0.05, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, ntolLBM=4, dtolLBM=1e-30)
    b = oracle.LBMBlock(16, 16, 14, flow=fl)
    start(b)
    body = _plate_body(oracle)
    b.update_volume_force(); b.calculate_macro_quantities(); b.ResetVolumeForce()
    it = b.calculate_interaction_force([body])
    assert it == 4
    assert np.all(body.v_Ei >= 1) and np.all(body.v_Ei[:, 0:4] <= 16)
    np.testing.assert_allclose(b.force.sum(axis=(1, 2, 3)) * b.dh ** 3, -body.v_Eforce.sum(axis=0), rtol=1e-6, atol=1e-18)
    # the fluid pushes the body downstream: marker force on the fluid is -x, v_Eforce is the force on the body (+x)
    assert body.v_Eforce[:, 0].sum() > 0


def test_ibm_stencil_out_of_domain_stops(oracle):
    fl = oracle.Flow()
    b = oracle.LBMBlock(8, 8, 8, BndConds=(101, 104, 301, 301, 301, 301), flow=fl)
    start(b)
    body = oracle.VirtualBody(1)
    body.v_Exyz[0] = (0.2, 4.0, 4.0)   # stencil reaches x index -1 on a non-periodic face: reference stops
    body.v_Ea[0] = -1.0
    with pytest.raises(ValueError):
        b.calculate_interaction_force([body])


def test_ibm_stencil_folds_at_walls(oracle):
    # trimedindex: symmetric / full-way wall mirrors 0 -> 2, half-way wall 0 -> 1 (Solidbody.f90:845-848)
    for code, expect in ((201, 2), (302, 2), (203, 1)):
        b = oracle.LBMBlock(8, 8, 8, BndConds=(code, code, 301, 301, 301, 301))
        start(b)
        body = oracle.VirtualBody(1)
        body.v_Exyz[0] = (0.4, 4.0, 4.0)
        body.v_Ea[0] = -1.0
        b.calculate_interaction_force([body])
        assert list(body.v_Ei[0, 0:4]) == [expect, 1, 2, 3]
