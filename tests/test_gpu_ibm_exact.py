"""The ordered IBM mode (default): interpolation walks a marker's 64 nodes in the order of the reference's nested
loops (Solidbody.f90:1009-1015) and spreading is a per-cell gather over entries sorted in the order the reference's
serial marker loops reach the cell (:1034-1048, :938-978).  On one GPU the marker forces, the corrected velocity and
the fluid state are therefore BIT-IDENTICAL to the oracle -- asserted here with array_equal -- and identical run to
run.  The atomic mode (fsilbm_set_option("ibm_ordered", 0)) is kept as the comparison arm and agrees to round-off."""
import numpy as np
import pytest

from tests.test_gpu_parity import plates_pair, sync_oracle_body

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fsilbm3d_b200 as F
    F.lib()
    return F


def set_ordered(F, on):
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_ordered", 1 if on else 0))


def fluid_equal(ob, gb):
    ob.calculate_macro_quantities()
    den, uuu = gb.download_macro()
    return np.array_equal(gb.download_fIn(), ob.fIn) and np.array_equal(den, ob.den) and np.array_equal(uuu, ob.uuu)


@pytest.mark.parametrize("single_launch", [1, 0], ids=["cooperative_kernel", "kernel_per_phase"])
@pytest.mark.parametrize("moving", [False, True])
@pytest.mark.parametrize("bc", [(301,) * 6, (101, 104, 202, 202, 301, 301)])
def test_rigid_plate_bit_exact(oracle, F, moving, bc, single_launch):
    from tests.common import make_pair
    F._lib.check(F.lib().fsilbm_set_option(b"ibm_single_launch", single_launch))
    try:
        flow = dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, 2e-4, 0.0), Uref=0.05, ntolLBM=4, dtolLBM=1e-30)
        ob, gb = make_pair(oracle, F, (32, 28, 24), BndConds=bc, **flow)
        pg, po, ovb = plates_pair(oracle, F, 1.0, moving=moving)
        for n in range(1, 31):
            t = float(n)
            ob.set_blktime(t)
            po.UpdatePosVelArea(); sync_oracle_body(ovb, po)
            it_o = ob.step([ovb])
            po.structure(t, 1, ob.dh, ob.dh)
            it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg], time=t)
            assert it_o == it_g == 4
            assert np.array_equal(pg.body.v_Eforce, ovb.v_Eforce), n
        assert fluid_equal(ob, gb)
        gb.close()
    finally:
        F._lib.check(F.lib().fsilbm_set_option(b"ibm_single_launch", 1))


@pytest.mark.parametrize("bc", [(301,) * 6, (101, 104, 202, 202, 301, 301)])
def test_early_ibm_bit_exact(oracle, F, bc):
    """Early IBM: collide_stream updates the planes around the body first and the next interaction-force call runs beside the
    rest of the update on its own stream.  A moving plate well inside the block (so the overlap is actually taken, checked
    through fsilbm_ibm_early_count) must stay bit-identical to the oracle, and to the run with the overlap turned off."""
    from tests.common import make_pair
    flow = dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, 2e-4, 0.0), Uref=0.05, ntolLBM=4, dtolLBM=1e-30)
    runs = {}
    # (early, split): split = the planes around the body on the high-priority body stream beside the rest (the default) or
    # queued before the rest on the compute stream
    for early, split in ((1, 1), (1, 0), (0, 1)):
        F._lib.check(F.lib().fsilbm_set_option(b"ibm_early", early))
        F._lib.check(F.lib().fsilbm_set_option(b"update_split", split))
        try:
            ob, gb = make_pair(oracle, F, (48, 36, 32), BndConds=bc, **flow)
            pg, po, ovb = plates_pair(oracle, F, 1.0, moving=True, origin=(18.3, 14.2, 10.4))
            c0 = F.lib().fsilbm_ibm_early_count()
            for n in range(1, 26):
                t = float(n)
                ob.set_blktime(t)
                po.UpdatePosVelArea(); sync_oracle_body(ovb, po)
                it_o = ob.step([ovb])
                po.structure(t, 1, ob.dh, ob.dh)
                it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg], time=t)
                assert it_o == it_g == 4
                assert np.array_equal(pg.body.v_Eforce, ovb.v_Eforce), (early, split, n)
            taken = F.lib().fsilbm_ibm_early_count() - c0
            assert taken == (24 if early else 0), taken      # every call after the first update
            assert fluid_equal(ob, gb)
            runs[(early, split)] = gb.download_fIn()
            gb.close()
        finally:
            F._lib.check(F.lib().fsilbm_set_option(b"ibm_early", 1))
            F._lib.check(F.lib().fsilbm_set_option(b"update_split", 1))
    assert np.array_equal(runs[(0, 1)], runs[(1, 1)]) and np.array_equal(runs[(1, 0)], runs[(1, 1)])


def test_overlapping_bodies_and_wrap_bit_exact(oracle, F):
    """Three bodies, two of them sharing cells (Gauss-Seidel order, Solidbody.f90:898-903), early exit on the
    tolerance; then a plate that wraps periodically and folds at a wall (several nodes of one marker on one cell)."""
    from tests.common import make_pair
    flow = dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0), Uref=0.04, ntolLBM=20, dtolLBM=0.05)
    ob, gb = make_pair(oracle, F, (36, 24, 20), **flow)
    trio = [plates_pair(oracle, F, 1.0, origin=o, nEL=e, Nspan=s) for o, e, s in (((8.2, 9.1, 5.3), 6, 8), ((12.6, 10.4, 5.9), 6, 8), ((25.1, 4.3, 6.2), 4, 6))]
    its = []
    for n in range(1, 16):
        ob.set_blktime(float(n))
        for pg, po, ov in trio:
            po.UpdatePosVelArea(); sync_oracle_body(ov, po)
        it_o = ob.step([ov for _, _, ov in trio])
        it_g = F.tree_collision_streaming_IBM_FEM(gb, [pg for pg, _, _ in trio], time=float(n))
        assert it_o == it_g
        its.append(it_g)
        for pg, _, ov in trio:
            assert np.array_equal(pg.body.v_Eforce, ov.v_Eforce)
    assert 1 <= min(its) < 20
    assert fluid_equal(ob, gb)
    gb.close()

    flow = dict(nu=0.05, uvwIn=(0.0, 0.0, 0.0), volumeForceIn=(2e-6, 0, 0), Uref=0.01, ntolLBM=3, dtolLBM=1e-30)
    ob, gb = make_pair(oracle, F, (20, 16, 18), BndConds=(301, 301, 201, 203, 301, 301), **flow)
    pg, po, ovb = plates_pair(oracle, F, 1.0, origin=(17.2, 0.45, 14.2), nEL=6, Nspan=8)
    pg.dirc = po.dirc = np.array([0.0, 0.0, 1.0])
    pg.node_ref[:, 1] = 0.45; po.node_ref[:, 1] = 0.45
    pg.structure(0.0, 1, 0.0, 0.0); po.structure(0.0, 1, 0.0, 0.0)
    pg.PlateUpdatePosVelArea(); po.PlateUpdatePosVelArea()
    for n in range(1, 13):
        sync_oracle_body(ovb, po)
        assert ob.step([ovb]) == F.tree_collision_streaming_IBM_FEM(gb, [pg], solver=False) == 3
        assert np.array_equal(pg.body.v_Eforce, ovb.v_Eforce)
    assert fluid_equal(ob, gb)
    gb.close()


def test_atomic_mode_agrees_and_ordered_mode_is_reproducible(oracle, F):
    from tests.common import make_pair, rel_err
    flow = dict(nu=0.05, uvwIn=(0.05, 0.0, 0.0), Uref=0.05, ntolLBM=4, dtolLBM=1e-30)
    runs = {}
    try:
        for mode in ("ordered", "ordered_again", "atomic"):
            set_ordered(F, mode != "atomic")
            ob, gb = make_pair(oracle, F, (32, 28, 24), BndConds=(101, 104, 301, 301, 301, 301), **flow)
            pg, _, _ = plates_pair(oracle, F, 1.0, moving=True)
            for n in range(1, 21):
                F.tree_collision_streaming_IBM_FEM(gb, [pg], time=float(n))
            runs[mode] = (gb.download_fIn(), pg.body.v_Eforce.copy())
            gb.close()
    finally:
        set_ordered(F, True)
    assert np.array_equal(runs["ordered"][0], runs["ordered_again"][0]) and np.array_equal(runs["ordered"][1], runs["ordered_again"][1])
    assert rel_err(runs["atomic"][1], runs["ordered"][1]) <= 1e-10
    assert rel_err(runs["atomic"][0], runs["ordered"][0]) <= 1e-12
