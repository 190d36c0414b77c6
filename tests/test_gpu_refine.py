"""Grid-refinement parity (LBMBlockComm.f90:279-318, :340-979): father + son blocks with 2:1 sub-cycling on the CUDA path
against the C oracle; fluid-only, so bit-exact."""
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


CASES = {
    # name: scheme, father bc, father dims, son bc, son dims, son mins, models (father, son)
    "linear_all_faces": (1, (102, 104, 301, 301, 203, 203), (20, 16, 14), (0,) * 6, (17, 13, 11), (5.0, 4.0, 3.0), (1, 1)),
    "cubic_all_faces": (2, (102, 104, 301, 301, 203, 203), (20, 16, 14), (0,) * 6, (17, 13, 11), (5.0, 4.0, 3.0), (1, 1)),
    "linear_son_periodic_y": (1, (101, 104, 301, 301, 301, 301), (20, 16, 14), (0, 0, 301, 301, 0, 0), (17, 32, 11), (5.0, 0.0, 3.0), (1, 2)),
    "cubic_son_periodic_y_z": (2, (101, 104, 301, 301, 301, 301), (20, 16, 14), (0, 0, 301, 301, 301, 301), (17, 32, 28), (5.0, 0.0, 0.0), (2, 1)),
    "linear_son_on_walls": (1, (102, 104, 301, 301, 203, 203), (20, 16, 14), (0, 0, 0, 0, 203, 203), (17, 13, 27), (5.0, 4.0, 0.0), (3, 3)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_father_son_parity(oracle, F, name):
    from tests.common import perturbed_state
    O = oracle
    scheme, fbc, fdims, sbc, sdims, smins, models = CASES[name]
    kw = dict(nu=0.02, uvwIn=(0.03, 0.005, 0.0), Uref=0.03, volumeForceIn=(1e-6, 0.0, 0.0))
    params = (0.25,) + (0.0,) * 9
    of, gf = O.Flow(**kw), F.FlowCondType(**kw)
    oF = O.LBMBlock(*fdims, dh=1.0, BndConds=fbc, iCollidModel=models[0], params=params, flow=of)
    oS = O.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, iCollidModel=models[1], params=params, flow=of)
    gF = F.LBMBlock(*fdims, dh=1.0, BndConds=fbc, iCollidModel=models[0], params=params, flow=gf)
    gS = F.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, iCollidModel=models[1], params=params, flow=gf)
    for b in (oF, oS, gF, gS):
        b.initialise(0.0)
    f0F, f0S = perturbed_state(fdims, of, seed=1), perturbed_state(sdims, of, seed=2)
    oF.fIn[...] = f0F; oS.fIn[...] = f0S
    gF.upload_fIn(f0F); gS.upload_fIn(f0S)
    oroot = O.TreeNode(oF); oroot.add_son(O.TreeNode(oS), scheme)
    groot = F.build_block_tree([gS, gF], interpolateScheme=scheme)     # order on purpose: the tree finds the root itself
    assert groot.block is gF and len(groot.sons) == 1 and groot.sons[0].block is gS
    po, pg = oroot.comm[0], groot.comm[0]
    assert (po.sds, po.s, po.f, po.si, po.fi, po.dimS, po.dimF) == (pg.sds, pg.s, pg.f, pg.si, pg.fi, pg.dimS, pg.dimF)
    for b in (oF, oS):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for b in (gF, gS):
        b.update_volume_force(); b.set_boundary_conditions()
    for n in range(1, 11):
        O.set_blktime_all(oroot, float(n)); F.set_blktime_all(groot, float(n))
        O.tree_collision_streaming_IBM_FEM(oroot)
        F.tree_collision_streaming_IBM_FEM(groot)
        if n in (1, 2, 10):
            fF, fS = gF.download_fIn(), gS.download_fIn()
            assert np.array_equal(fF, oF.fIn), (n, "father", np.abs(fF - oF.fIn).max())
            assert np.array_equal(fS, oS.fIn), (n, "son", np.abs(fS - oS.fIn).max())
    for p in groot.comm:
        p.close()
    gF.close(); gS.close()


def test_mismatched_grids_are_rejected(F):
    gf = F.FlowCondType(nu=0.02)
    gF = F.LBMBlock(20, 16, 14, dh=1.0, flow=gf)
    gS = F.LBMBlock(16, 13, 11, dh=0.5, xmin=5.0, ymin=4.0, zmin=3.0, BndConds=(0,) * 6, flow=gf)   # even x extent, not periodic
    with pytest.raises(F.FsilbmError):
        F.CommPair(gF, gS)
    gF.close(); gS.close()


def test_plate_carried_by_son_block(oracle, F):
    """A rigid plate inside the refined son block (bodies go to the finest block containing their first marker,
    FluidDomain.f90:1974-2017); stencil folding uses the ROOT block's boundary codes (Solidbody.f90:337)."""
    from tests.common import perturbed_state, rel_err
    O = oracle
    kw = dict(nu=0.02, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, ntolLBM=3, dtolLBM=1e-30)
    fbc, fdims, sbc, sdims, smins = (101, 104, 301, 301, 301, 301), (24, 16, 16), (0,) * 6, (25, 17, 17), (6.0, 4.0, 4.0)
    of, gf = O.Flow(**kw), F.FlowCondType(**kw)
    oF = O.LBMBlock(*fdims, dh=1.0, BndConds=fbc, flow=of)
    oS = O.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, flow=of)
    gF = F.LBMBlock(*fdims, dh=1.0, BndConds=fbc, flow=gf)
    gS = F.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, flow=gf)
    for b in (oF, oS, gF, gS):
        b.initialise(0.0)
    kwp = dict(origin=(9.3, 7.2, 6.1), nEL=6, len1=0.5, Nspan=8, spanlen=4.0, Lspan=0.0, chord_dir=(1.0, 0.3, 0.0), denIn=1.0)
    pg = F.RigidPlate(**kwp)
    assert F.find_carrier_fluidblock([gF, gS], pg.body.v_Exyz[0]) == 1
    ov = O.VirtualBody(pg.body.v_nelmts)
    ov.v_Exyz[...] = pg.body.v_Exyz; ov.v_Evel[...] = pg.body.v_Evel; ov.v_Ea[...] = pg.body.v_Ea
    oroot = O.TreeNode(oF); oroot.add_son(O.TreeNode(oS, [ov]))
    groot = F.blockTreeNode(gF); groot.add_son(F.blockTreeNode(gS, [pg]))
    for b in (oF, oS):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for b in (gF, gS):
        b.update_volume_force(); b.set_boundary_conditions()
    for n in range(1, 9):
        O.set_blktime_all(oroot, float(n)); F.set_blktime_all(groot, float(n))
        io, ig = [], []
        O.tree_collision_streaming_IBM_FEM(oroot, iters=io)
        F.tree_collision_streaming_IBM_FEM(groot, solver=False, iters=ig)
        assert io == ig == [0, 3, 3]
        assert rel_err(pg.body.v_Eforce, ov.v_Eforce) <= 1e-10
    oF.calculate_macro_quantities(); oS.calculate_macro_quantities()
    for gb, ob in ((gF, oF), (gS, oS)):
        den, uuu = gb.download_macro()
        assert rel_err(den, ob.den) <= 1e-12 and rel_err(uuu, ob.uuu) <= 1e-12
    assert abs(pg.body.v_Eforce[:, 0].sum()) > 1e-9


def test_plate_in_son_early_ibm(oracle, F):
    """Early IBM on a son block: the father rewrites only the son's outermost planes (interpolation_father_to_son), so the
    son's next interaction-force call may start behind the planes around the plate -- the first of a root step even while the
    father is being updated.  A heaving plate well inside the son; bit-identical to the oracle tree, overlap asserted taken."""
    from tests.common import perturbed_state
    from tests.test_gpu_parity import sync_oracle_body
    O = oracle
    kw = dict(nu=0.02, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, ntolLBM=3, dtolLBM=1e-30)
    fbc, fdims, sbc, sdims, smins = (101, 104, 301, 301, 301, 301), (40, 24, 24), (0,) * 6, (41, 33, 33), (8.0, 4.0, 4.0)
    of, gf = O.Flow(**kw), F.FlowCondType(**kw)
    oF = O.LBMBlock(*fdims, dh=1.0, BndConds=fbc, flow=of)
    oS = O.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, flow=of)
    gF = F.LBMBlock(*fdims, dh=1.0, BndConds=fbc, flow=gf)
    gS = F.LBMBlock(*sdims, dh=0.5, xmin=smins[0], ymin=smins[1], zmin=smins[2], BndConds=sbc, flow=gf)
    for b in (oF, oS, gF, gS):
        b.initialise(0.0)
    f0F, f0S = perturbed_state(fdims, of, seed=1), perturbed_state(sdims, of, seed=2)
    oF.fIn[...] = f0F; oS.fIn[...] = f0S
    gF.upload_fIn(f0F); gS.upload_fIn(f0S)
    kwp = dict(origin=(15.3, 11.2, 9.6), nEL=8, len1=0.5, Nspan=10, spanlen=5.0, Lspan=0.0, chord_dir=(1.0, 0.3, 0.0), denIn=1.0,
               XYZAmpl=(0.0, 0.6, 0.0), Freq=0.02, XYZPhi=(0.0, 0.3, 0.0))
    pg, po = F.RigidPlate(**kwp), F.RigidPlate(**kwp)
    ov = O.VirtualBody(pg.body.v_nelmts, v_move=1, iBodyModel=1)
    oroot = O.TreeNode(oF)
    oson = oroot.add_son(O.TreeNode(oS, [ov]))
    pair = oroot.comm[0]
    groot = F.blockTreeNode(gF); groot.add_son(F.blockTreeNode(gS, [pg]))
    for b in (oF, oS):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for b in (gF, gS):
        b.update_volume_force(); b.set_boundary_conditions()
    c0 = F.lib().fsilbm_ibm_early_count()
    steps = 8
    for n in range(1, steps + 1):
        O.set_blktime_all(oroot, float(n)); F.set_blktime_all(groot, float(n))
        ig = []
        F.tree_collision_streaming_IBM_FEM(groot, solver=True, iters=ig)
        # the oracle tree by hand in the order of LBMBlockComm.f90:279-318, moving the plate before each son sub-step as the
        # reference's IBM_FEM does (extract(1) reads fIn, which the father's pre-collision work does not touch)
        pair.extract_interpolate_layer(1)
        O.tree_collision_streaming_IBM_FEM(O.TreeNode(oF))
        pair.extract_interpolate_layer(2)
        for k in range(2):
            oS.set_blktime(oS.blktime + float(k) * oS.dh)
            po.UpdatePosVelArea(); sync_oracle_body(ov, po)
            io = []
            O.tree_collision_streaming_IBM_FEM(oson, rootBC=oF.BndConds, iters=io)
            po.structure(oS.blktime, 1, oS.dh, oS.dh)
            pair.interpolation_father_to_son(k)
            assert io == [3]
        pair.deliver_son_to_father()
        assert ig == [0, 3, 3]
        assert np.array_equal(pg.body.v_Eforce, ov.v_Eforce), n
    assert F.lib().fsilbm_ibm_early_count() - c0 == 2 * steps - 1     # every son call after its first update
    assert np.array_equal(gF.download_fIn(), oF.fIn) and np.array_equal(gS.download_fIn(), oS.fIn)
    assert np.abs(pg.body.v_Exyz - po.body.v_Exyz).max() == 0.0 and abs(pg.body.v_Eforce[:, 0].sum()) > 1e-9
    for p in groot.comm:
        p.close()
    gF.close(); gS.close()
