"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances are north_star's: <= 1e-12 relative on density and velocity after N steps, <= 1e-10 on
body forces.  Where no atomics are involved the comparison is also reported bit-for-bit: the
library is built with -fmad=false and follows the reference's evaluation order, so fluid-only
cases are expected to agree exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FLUID = 1e-12
TOL_FORCE = 1e-10


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fsilbm3d_b200 as F
    return F


def compare_fluid(ob, gb, exact=True):
    ob.calculate_macro_quantities()
    den, uuu = gb.download_macro()
    f = gb.download_fIn()
    from tests.common import rel_err
    e_den, e_u, e_f = rel_err(den, ob.den), rel_err(uuu, ob.uuu), rel_err(f, ob.fIn)
    assert e_den <= TOL_FLUID and e_u <= TOL_FLUID and e_f <= TOL_FLUID, (e_den, e_u, e_f)
    if exact:
        assert np.array_equal(f, ob.fIn), f"not bit-exact: max |df| = {np.max(np.abs(f - ob.fIn))}"
    return e_den, e_u, e_f


@pytest.mark.parametrize("model,params", [(1, (0.0,) * 10), (2, (3 / 16,) + (0.0,) * 9), (3, (0.0,) * 10)])
def test_periodic_body_force_channel(oracle, F, model, params):
    """configs[1] (periodic body-force channel) at a size the oracle finishes in seconds."""
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (40, 24, 36), model=model, params=params, nu=0.1, volumeForceIn=(1e-6, 2e-7, -3e-7))
    for n in range(1, 101):
        ob.set_blktime(float(n)); gb.set_blktime(float(n))
        ob.step(); gb.step()
    compare_fluid(ob, gb)


def test_oscillating_volume_force(oracle, F):
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (16, 12, 20), nu=0.05, volumeForceIn=(1e-6, 0, 0), volumeForceAmp=5e-7,
                       volumeForceFreq=0.01, volumeForcePhi=30.0)
    for n in range(1, 41):
        ob.set_blktime(float(n)); gb.set_blktime(float(n))
        ob.step(); gb.step()
    assert np.allclose(gb.volumeForce, ob.volumeForce, rtol=0, atol=0)
    compare_fluid(ob, gb)


BC_CASES = {
    "halfway_channel": dict(bc=(301, 301, 203, 203, 301, 301), flow=dict(nu=0.05, volumeForceIn=(1e-6, 0, 0))),
    "fullway_walls": dict(bc=(301, 301, 201, 201, 201, 201), flow=dict(nu=0.05, volumeForceIn=(1e-6, 0, 0))),
    "moving_walls_shear": dict(bc=(301, 301, 202, 202, 301, 301), flow=dict(nu=0.05, uvwIn=(0.02, 0.0, 0.0), shearRateIn=(0.0, 1e-3, 0.0))),
    "halfway_moving_walls": dict(bc=(301, 301, 204, 204, 301, 301), flow=dict(nu=0.05, uvwIn=(0.02, 0.0, 0.0), shearRateIn=(0.0, 1e-3, 0.0))),
    "inlet_eq_outlet_o2": dict(bc=(101, 104, 301, 301, 301, 301), flow=dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0))),
    "inlet_neq_outlet_o1": dict(bc=(102, 103, 301, 301, 301, 301), flow=dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0))),
    "symmetric_sides": dict(bc=(101, 104, 302, 302, 302, 302), flow=dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0))),
    "all_faces_mixed": dict(bc=(102, 104, 202, 204, 203, 201), flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), shearRateIn=(0.0, 5e-4, 2e-4))),
    "neq_on_every_axis": dict(bc=(102, 102, 102, 102, 102, 102), flow=dict(nu=0.05, uvwIn=(0.03, 0.01, -0.01))),
    "extrapolate_everywhere": dict(bc=(103, 104, 104, 103, 103, 104), flow=dict(nu=0.05, uvwIn=(0.03, 0.01, -0.01))),
    "oscillatory_inflow": dict(bc=(101, 104, 301, 301, 301, 301), flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), velocityKind=2, shearRateIn=(0.01, 0.02, 45.0))),
}


@pytest.mark.parametrize("name", sorted(BC_CASES))
@pytest.mark.parametrize("model", [1, 2])
def test_boundary_conditions(oracle, F, name, model):
    """Every boundary code of ConstParams.f90:31-34 incl. face-order precedence on shared edges and the
    first-call skip of the half-way codes (FluidDomain.f90:660-661)."""
    from tests.common import make_pair
    case = BC_CASES[name]
    ob, gb = make_pair(oracle, F, (14, 18, 22), BndConds=case["bc"], model=model, params=(0.25,) + (0.0,) * 9, dh=0.5,
                       mins=(-1.0, 0.5, 2.0), **case["flow"])
    for n in range(1, 31):
        ob.set_blktime(0.5 * n); gb.set_blktime(0.5 * n)
        ob.step(); gb.step()
    compare_fluid(ob, gb)


def test_ragged_sizes(oracle, F):
    """Extents that are not multiples of the warp/block shape, down to the smallest the BCs allow."""
    from tests.common import make_pair
    for dims in [(3, 3, 3), (5, 7, 33), (4, 130, 31), (9, 2, 129), (2, 5, 257)]:
        ob, gb = make_pair(oracle, F, dims, BndConds=(301, 301, 301, 301, 301, 301), nu=0.07, volumeForceIn=(1e-6, 1e-6, 1e-6))
        for n in range(1, 8):
            ob.step(); gb.step()
        compare_fluid(ob, gb)
        gb.close()


def test_ghost_plane_streaming_matches_wrap(oracle, F):
    """The multi-rank streaming path (push into ghost planes, fold back) on one GPU."""
    from tests.common import make_pair
    from fsilbm3d_b200._lib import lib, check
    check(lib().fsilbm_set_option(b"force_ghost", 1))
    try:
        ob, gb = make_pair(oracle, F, (12, 10, 40), BndConds=(301, 301, 203, 203, 301, 301), nu=0.1, volumeForceIn=(1e-6, 0, 0))
        for n in range(20):
            ob.step(); gb.step()
        compare_fluid(ob, gb)
    finally:
        check(lib().fsilbm_set_option(b"force_ghost", 0))


def test_unknown_option_is_an_error(F):
    """The kernel sweep arms of round 1 ("variant": streaming stores, pull) are gone: one kernel form, one path."""
    from fsilbm3d_b200._lib import lib
    assert lib().fsilbm_set_option(b"variant", 1) != 0


@pytest.mark.parametrize("model", [1, 2, 3])
def test_unfused_passes_match_reference_procedures(oracle, F, model):
    """Procedure by procedure, in the order of LBMBlockComm.f90:283-303."""
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (10, 12, 34), BndConds=(301, 301, 203, 203, 301, 301), model=model, params=(0.2,) + (0.0,) * 9,
                       nu=0.08, volumeForceIn=(1e-6, -1e-6, 5e-7))
    for n in range(5):
        ob.update_volume_force(); gb.update_volume_force()
        ob.calculate_macro_quantities(); gb.calculate_macro_quantities()
        ob.ResetVolumeForce(); gb.ResetVolumeForce()
        ob.add_volume_force(); gb.add_volume_force()
        den, uuu, force = gb.download_fields()
        assert np.array_equal(den, ob.den) and np.array_equal(uuu, ob.uuu) and np.array_equal(force, ob.force)
        ob.collision(); gb.collision()
        assert np.array_equal(gb.download_fIn(), ob.fIn)
        ob.halfwayBCset(); gb.halfwayBCset()
        ob.streaming(); gb.streaming()
        assert np.array_equal(gb.download_fIn(), ob.fIn)
        ob.set_boundary_conditions(); gb.set_boundary_conditions()
        assert np.array_equal(gb.download_fIn(), ob.fIn)


def test_field_stat(oracle, F):
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (12, 10, 40), nu=0.1, uvwIn=(0.05, 0.01, 0.0), volumeForceIn=(1e-6, 0, 0), Uref=0.05)
    for n in range(5):
        ob.step(); gb.step()
    ob.calculate_macro_quantities()
    np.testing.assert_allclose(gb.ComputeFieldStat(), ob.ComputeFieldStat(), rtol=1e-12)


# ---- IBM --------------------------------------------------------------------------------------------
def plates_pair(oracle, F, denIn, moving=False, origin=(10.3, 9.2, 6.4), nEL=8, Nspan=10):
    kw = dict(origin=origin, nEL=nEL, len1=1.0, Nspan=Nspan, spanlen=float(Nspan), Lspan=0.0, chord_dir=(1.0, 0.35, 0.0),
              span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha=1.0, denIn=denIn)
    if moving:
        kw.update(XYZAmpl=(0.0, 1.5, 0.0), Freq=0.01, XYZPhi=(0.0, 0.3, 0.0))
    pg = F.RigidPlate(**kw)
    po = F.RigidPlate(**kw)
    ovb = oracle.VirtualBody(pg.body.v_nelmts, v_move=pg.body.v_move, iBodyModel=1)
    return pg, po, ovb


def sync_oracle_body(ovb, plate):
    ovb.v_Exyz[...] = plate.body.v_Exyz
    ovb.v_Evel[...] = plate.body.v_Evel
    ovb.v_Ea[...] = plate.body.v_Ea


@pytest.mark.parametrize("moving", [False, True])
@pytest.mark.parametrize("bc", [(301,) * 6, (101, 104, 202, 202, 301, 301)])
def test_rigid_plate_ibm(oracle, F, moving, bc):
    """configs[2] reduced: rigid plate, uniform/shear inflow, fixed iteration count (ntolLBM=4, dtolLBM tiny)."""
    from tests.common import make_pair, rel_err
    flow = dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, 2e-4, 0.0), Uref=0.05, ntolLBM=4, dtolLBM=1e-30)
    ob, gb = make_pair(oracle, F, (32, 28, 24), BndConds=bc, **flow)
    pg, po, ovb = plates_pair(oracle, F, 1.0, moving=moving)
    for n in range(1, 41):
        t = float(n)
        ob.set_blktime(t)
        po.UpdatePosVelArea(); sync_oracle_body(ovb, po)
        it_o = ob.step([ovb])
        po.structure(t, 1, ob.dh, ob.dh)
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg], time=t)
        assert it_o == it_g == 4
        if n in (1, 40):
            Ei, Ew = gb.download_stencil(0, pg.body.v_nelmts)
            assert np.array_equal(Ei, ovb.v_Ei) and np.array_equal(Ew, ovb.v_Ew)
        e = rel_err(pg.body.v_Eforce, ovb.v_Eforce)
        assert e <= TOL_FORCE, (n, e)
    compare_fluid(ob, gb, exact=False)
    assert abs(pg.body.v_Eforce[:, 0].sum()) > 1e-8   # the plate does feel the flow


def test_two_plates_gauss_seidel_and_convergence_exit(oracle, F):
    """Two bodies whose stencils overlap (body 2 sees body 1's correction, Solidbody.f90:898-903) and
    a tolerance that stops the loop before ntolLBM."""
    from tests.common import make_pair, rel_err
    flow = dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, ntolLBM=20, dtolLBM=0.05)
    ob, gb = make_pair(oracle, F, (36, 24, 20), **flow)
    pg1, po1, ov1 = plates_pair(oracle, F, 1.0, origin=(8.2, 9.1, 5.3), nEL=6, Nspan=8)
    pg2, po2, ov2 = plates_pair(oracle, F, 1.0, origin=(12.6, 10.4, 5.9), nEL=6, Nspan=8)
    pg3, po3, ov3 = plates_pair(oracle, F, 1.0, origin=(25.1, 4.3, 6.2), nEL=4, Nspan=6)   # far away: its own box
    its = []
    for n in range(1, 21):
        t = float(n)
        ob.set_blktime(t)
        for po, ov in ((po1, ov1), (po2, ov2), (po3, ov3)):
            po.UpdatePosVelArea(); sync_oracle_body(ov, po)
        it_o = ob.step([ov1, ov2, ov3])
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg1, pg2, pg3], time=t)
        assert it_o == it_g, (n, it_o, it_g)
        its.append(it_g)
        for pg, ov in ((pg1, ov1), (pg2, ov2), (pg3, ov3)):
            assert rel_err(pg.body.v_Eforce, ov.v_Eforce) <= TOL_FORCE
    assert 1 <= min(its) < 20, its   # the loop really stopped on the tolerance
    compare_fluid(ob, gb, exact=False)


def test_ibm_near_walls_and_periodic_wrap(oracle, F):
    """Stencil folding at a full-way wall and a plate that straddles the periodic z boundary."""
    from tests.common import make_pair, rel_err
    flow = dict(nu=0.05, uvwIn=(0.0, 0.0, 0.0), volumeForceIn=(2e-6, 0, 0), Uref=0.01, ntolLBM=3, dtolLBM=1e-30)
    ob, gb = make_pair(oracle, F, (20, 16, 18), BndConds=(301, 301, 201, 203, 301, 301), **flow)
    pg, po, ovb = plates_pair(oracle, F, 1.0, origin=(17.2, 0.45, 14.2), nEL=6, Nspan=8)   # wraps in x and z, touches y-min wall
    pg.dirc = po.dirc = np.array([0.0, 0.0, 1.0])
    pg.node_ref[:, 1] = 0.45; po.node_ref[:, 1] = 0.45   # keep the whole plate within the first cell row
    pg.structure(0.0, 1, 0.0, 0.0); po.structure(0.0, 1, 0.0, 0.0)
    pg.PlateUpdatePosVelArea(); po.PlateUpdatePosVelArea()
    for n in range(1, 16):
        sync_oracle_body(ovb, po)
        it_o = ob.step([ovb])
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg], solver=False)
        assert it_o == it_g == 3
        assert rel_err(pg.body.v_Eforce, ovb.v_Eforce) <= TOL_FORCE
    Ei, Ew = gb.download_stencil(0, pg.body.v_nelmts)
    assert np.array_equal(Ei, ovb.v_Ei) and np.array_equal(Ew, ovb.v_Ew)
    assert (Ei[:, 4:8] == 2).any()   # folded y index (0 -> 2) really occurred
    compare_fluid(ob, gb, exact=False)


def test_ibm_stencil_out_of_domain_is_an_error(oracle, F):
    from tests.common import make_pair
    ob, gb = make_pair(oracle, F, (10, 10, 10), BndConds=(101, 104, 301, 301, 301, 301), nu=0.05)
    body = F.VirtualBody(1)
    body.v_Exyz[0] = (0.2, 4.0, 4.0)
    body.v_Ea[0] = -1.0
    with pytest.raises(F.FsilbmError) as ei:
        gb.calculate_interaction_force([body])
    assert ei.value.code == 4


# ---- full-size properties (BASELINE configs[1] grid), no oracle needed -----------------------------------
def test_full_size_uniform_force_and_mass(F):
    """256^3 periodic channel: u = (n+1/2) F dh / rho at the start of step n+1 and mass is conserved."""
    Fx = 1e-6
    gb = F.LBMBlock(256, 256, 256, flow=F.FlowCondType(nu=0.1, volumeForceIn=(Fx, 0.0, 0.0)))
    gb.initialise(0.0)
    gb.update_volume_force(); gb.set_boundary_conditions()
    n = 20
    for _ in range(n):
        gb.step()
    den, uuu = gb.download_macro()
    np.testing.assert_allclose(uuu[0], (n + 0.5) * Fx, rtol=1e-10)
    assert np.max(np.abs(uuu[1:])) < 1e-15   # round-off only: no force, no gradient in y, z
    np.testing.assert_allclose(den, 1.0, rtol=1e-13)
    gb.close()


def test_full_size_streaming_period(F):
    """One-hot populations return after X (Y, Z) steps when the collision leaves them alone: here via the
    un-fused streaming pass on a 64x48x256 block (checks every direction's wrap at full z extent)."""
    X, Y, Z = 8, 6, 256
    gb = F.LBMBlock(X, Y, Z)
    gb.initialise(0.0)
    rng = np.random.default_rng(3)
    f0 = rng.uniform(0, 1, (19, X, Y, Z))
    gb.upload_fIn(f0)
    gb.streaming()
    f1 = gb.download_fIn()
    ee = [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0),
          (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1), (0, 1, 1), (0, -1, 1), (0, 1, -1), (0, -1, -1)]
    for q, e in enumerate(ee):
        assert np.array_equal(f1[q], np.roll(f0[q], e, axis=(0, 1, 2))), q
    gb.close()


# ---- LES collision models (FluidDomain.f90:1239-1258, closures :1265-1507) ------------------------------------
@pytest.mark.parametrize("model", [11, 14, 15])
@pytest.mark.parametrize("bc", [(301,) * 6, (101, 104, 203, 203, 201, 302)])
def test_les_models(oracle, F, model, bc):
    """Smagorinsky (local), WALE and Vreman (neighbour differences of this step's velocity, one-sided on the outer
    planes; the reference's `invdh = dh` quirk kept), fused path and pass-by-pass path."""
    from tests.common import make_pair, perturbed_state, rel_err
    kw = dict(nu=0.002, uvwIn=(0.05, 0.0, 0.0), volumeForceIn=(1e-6, 0.0, 0.0))
    dims = (12, 10, 36)
    ob, gb = make_pair(oracle, F, dims, BndConds=bc, model=model, perturb=False, **kw)
    f0 = perturbed_state(dims, ob.flow, wave_amp=2e-2)
    ob.fIn[...] = f0; gb.upload_fIn(f0)
    ob.set_boundary_conditions(); gb.set_boundary_conditions()
    for n in range(1, 16):
        ob.set_blktime(float(n)); gb.set_blktime(float(n))
        ob.step(); gb.step()
    exact = model != 14   # WALE uses pow(): last-bit differences between CUDA's and libm's pow
    compare_fluid(ob, gb, exact=exact)
    tg = gb.download_tau_all()
    assert rel_err(tg, ob.tau_all) <= TOL_FLUID
    assert ob.tau_all.max() > ob.tau * (1 + 1e-6)
    # pass by pass on top of the same state
    for n in range(3):
        for b in (ob, gb):
            b.update_volume_force(); b.calculate_macro_quantities(); b.ResetVolumeForce(); b.add_volume_force(); b.collision()
            b.halfwayBCset(); b.streaming(); b.set_boundary_conditions()
    compare_fluid(ob, gb, exact=exact)
    gb.close()


def test_les_with_plate(oracle, F):
    """WALE with an immersed plate: the velocity differences must see the IBM-corrected velocity."""
    from tests.common import make_pair, rel_err
    flow = dict(nu=0.002, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, ntolLBM=3, dtolLBM=1e-30)
    ob, gb = make_pair(oracle, F, (28, 24, 20), BndConds=(101, 104, 301, 301, 301, 301), model=14, **flow)
    pg, po, ovb = plates_pair(oracle, F, 1.0, origin=(9.3, 8.2, 5.4), nEL=6, Nspan=8)
    for n in range(1, 13):
        sync_oracle_body(ovb, po)
        ob.set_blktime(float(n))
        it_o = ob.step([ovb])
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg], time=float(n), solver=False)
        assert it_o == it_g == 3
        assert rel_err(pg.body.v_Eforce, ovb.v_Eforce) <= TOL_FORCE
    compare_fluid(ob, gb, exact=False)
    assert rel_err(gb.download_tau_all(), ob.tau_all) <= TOL_FLUID


def test_nine_mrt_blocks_share_or_refuse_matrix_entries(oracle, F):
    """The MRT matrix table (constant memory) holds 8 entries keyed by content: blocks of equal relaxation time share one, a
    ninth DISTINCT set of matrices is refused instead of overwriting block 0's (mrt_slot = slot % 8)."""
    from tests.common import make_pair
    pairs = []
    for k in range(9):        # nine live MRT blocks, two distinct viscosities -> two entries
        ob, gb = make_pair(oracle, F, (10, 8, 6), model=3, nu=0.05 if k % 2 == 0 else 0.08, volumeForceIn=(1e-6, 0.0, 0.0))
        pairs.append((ob, gb))
    for n in range(1, 6):
        for ob, gb in pairs:
            ob.set_blktime(float(n)); gb.set_blktime(float(n))
            ob.step(); gb.step()
    for ob, gb in pairs:
        assert np.array_equal(gb.download_fIn(), ob.fIn)
    # seven more distinct viscosities fill the table (2 + 6 = 8); the next one must be refused loudly
    extra = []
    with pytest.raises(F.FsilbmError, match="MRT"):
        for k in range(7):
            gb = F.LBMBlock(6, 6, 6, iCollidModel=3, flow=F.FlowCondType(nu=0.1 + 0.01 * k))
            extra.append(gb)
            gb.initialise(0.0)
    assert len(extra) == 7
    for gb in extra:
        gb.close()
    gb = F.LBMBlock(6, 6, 6, iCollidModel=3, flow=F.FlowCondType(nu=0.2))   # entries freed by close() are reusable
    gb.initialise(0.0)
    gb.close()
    for ob, gb in pairs:
        gb.close()


def test_more_separate_bodies_than_box_table_entries(oracle, F):
    """20 small plates far apart in one block: more stencil boxes than the 16-entry box table.  The nearest boxes are joined (a box
    is only storage), and the result stays bit-identical to the oracle."""
    from tests.common import make_pair
    flow = dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, ntolLBM=3, dtolLBM=1e-30)
    ob, gb = make_pair(oracle, F, (72, 60, 20), BndConds=(101, 104, 301, 301, 301, 301), **flow)
    pgs, ovs = [], []
    for i in range(5):
        for j in range(4):
            kw = dict(origin=(8.3 + 13.0 * i, 7.2 + 14.0 * j, 8.4), nEL=2, len1=1.0, Nspan=3, spanlen=3.0, Lspan=0.0, chord_dir=(1.0, 0.2, 0.0),
                      span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha=1.0, denIn=1.0)
            pg = F.RigidPlate(**kw)
            ov = oracle.VirtualBody(pg.body.v_nelmts, v_move=0, iBodyModel=1)
            ov.v_Exyz[...] = pg.body.v_Exyz; ov.v_Evel[...] = pg.body.v_Evel; ov.v_Ea[...] = pg.body.v_Ea
            pgs.append(pg); ovs.append(ov)
    for n in range(1, 9):
        ob.set_blktime(float(n))
        it_o = ob.step(ovs)
        it_g = F.tree_collision_streaming_IBM_FEM(gb, pgs, time=float(n))
        assert it_o == it_g == 3
    for pg, ov in zip(pgs, ovs):
        assert np.array_equal(pg.body.v_Eforce, ov.v_Eforce)
    compare_fluid(ob, gb)
    gb.close()
