"""The C oracle against the committed golden vectors (tests/golden/, produced by the independent numpy
restatement oracle/numpy_restatement.py) and against that restatement run live.  Fluid cases must agree
bit for bit: both follow the reference's evaluation order and neither uses fused multiply-adds."""
import numpy as np
import pytest

from tests.common import golden_names, load_golden, rel_err, run_golden_case


def _oracle_backend(O):
    def make_block(c):
        b = O.LBMBlock(*c["dims"], dh=c["dh"], xmin=c["mins"][0], ymin=c["mins"][1], zmin=c["mins"][2], BndConds=c["bc"],
                       iCollidModel=c["model"], params=c["params"], flow=O.Flow(**c["flow"]))
        b.initialise(0.0)
        return b

    def step(b, bodies, t):
        b.set_blktime(t)
        return b.step(bodies)
    return make_block, (lambda n: O.VirtualBody(n)), step


def test_golden_files_present():
    assert len(golden_names()) >= 8


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_reproduces_golden(oracle, name):
    case, g = load_golden(name)
    mk, mb, st = _oracle_backend(oracle)
    blk, bodies, its = run_golden_case(case, g, mk, mb, st)
    blk.calculate_macro_quantities()
    assert np.array_equal(its, g["iters"])
    assert np.array_equal(blk.fIn, g["fIn"]), f"max |df| {np.abs(blk.fIn - g['fIn']).max():.3e}"
    assert np.array_equal(blk.den, g["den"]) and np.array_equal(blk.uuu, g["uuu"])
    if "tau_all" in g.files:
        assert np.array_equal(blk.tau_all, g["tau_all"])
        assert g["tau_all"].max() > blk.tau * (1 + 1e-6)   # the closure really raised the relaxation time somewhere
    if bodies:
        assert np.array_equal(bodies[0].v_Ei, g["Ei"]) and np.array_equal(bodies[0].v_Ew, g["Ew"])
        assert np.array_equal(bodies[0].v_Eforce, g["Eforce"])


def test_numpy_restatement_live_matches_c_oracle_unfused_passes(oracle):
    """Pass by pass (LBMBlockComm.f90:283-303 order) on a case that is not among the golden files."""
    from oracle import numpy_restatement as NR
    from tests.common import perturbed_state
    dims, bc = (6, 7, 9), (102, 103, 201, 302, 204, 203)
    kw = dict(nu=0.06, uvwIn=(0.02, 0.01, -0.01), shearRateIn=(1e-4, 2e-4, 3e-4), volumeForceIn=(1e-6, -2e-6, 5e-7))
    for model, params in ((1, (0.0,) * 10), (2, (0.2,) + (0.0,) * 9), (3, (0.0,) * 10)):
        ob = oracle.LBMBlock(*dims, dh=0.5, xmin=-1.0, ymin=0.25, zmin=3.0, BndConds=bc, iCollidModel=model, params=params, flow=oracle.Flow(**kw))
        nb = NR.Block(*dims, dh=0.5, xmin=-1.0, ymin=0.25, zmin=3.0, BndConds=bc, iCollidModel=model, params=params, flow=NR.Flow(**kw))
        ob.initialise(0.0); nb.initialise(0.0)
        assert np.array_equal(ob.fIn, nb.f)
        if model == 3:
            assert np.array_equal(ob.M_COLLID, nb.M_COLLID) and np.array_equal(ob.M_FORCE, nb.M_FORCE)
        f0 = perturbed_state(dims, ob.flow)
        ob.fIn[...] = f0; nb.f[...] = f0
        ob.update_volume_force(); nb.update_volume_force()
        ob.set_boundary_conditions(); nb.set_boundary_conditions()
        assert np.array_equal(ob.fIn, nb.f)
        for n in range(1, 6):
            ob.set_blktime(0.5 * n); nb.blktime = 0.5 * n
            ob.update_volume_force(); nb.update_volume_force()
            ob.calculate_macro_quantities(); nb.calculate_macro_quantities()
            assert np.array_equal(ob.den, nb.den) and np.array_equal(ob.uuu, nb.uuu)
            ob.ResetVolumeForce(); nb.ResetVolumeForce()
            ob.add_volume_force(); nb.add_volume_force()
            assert np.array_equal(ob.force, nb.force)
            ob.collision(); nb.collision()
            assert np.array_equal(ob.fIn, nb.f), (model, n, "collision")
            ob.halfwayBCset(); nb.halfwayBCset()
            ob.streaming(); nb.streaming()
            assert np.array_equal(ob.fIn, nb.f), (model, n, "streaming")
            ob.set_boundary_conditions(); nb.set_boundary_conditions()
            assert np.array_equal(ob.fIn, nb.f), (model, n, "bc")
        assert abs(ob.tau - nb.tau) == 0 and abs(ob.Omega2 - nb.Omega2) == 0


@pytest.mark.parametrize("name", golden_names(refine=True))
def test_c_oracle_reproduces_refinement_golden(oracle, name):
    from tests.common import REFINE_FLOW, REFINE_PARAMS, run_golden_refine
    O = oracle
    case, g = load_golden(name)

    def make_block(dims, dh, mins, bc, model):
        b = O.LBMBlock(*dims, dh=dh, xmin=mins[0], ymin=mins[1], zmin=mins[2], BndConds=bc, iCollidModel=model, params=REFINE_PARAMS, flow=O.Flow(**REFINE_FLOW))
        b.initialise(0.0)
        return b

    def make_tree(Fb, Sb, scheme):
        root = O.TreeNode(Fb); root.add_son(O.TreeNode(Sb), scheme)
        return root
    Fb, Sb, p = run_golden_refine(case, g, make_block, make_tree, O.tree_collision_streaming_IBM_FEM, O.set_blktime_all)
    assert list(g["pair"]) == p.sds + p.s + p.f + p.si + p.fi + p.dimS + p.dimF
    assert np.array_equal(Fb.fIn, g["fF"]) and np.array_equal(Sb.fIn, g["fS"])


def test_refinement_keeps_uniform_flow_uniform(oracle):
    """Known answer: a uniform equilibrium stays uniform through father->son interpolation, the non-equilibrium
    rescale (zero non-equilibrium) and son->father restriction."""
    O = oracle
    fl = O.Flow(nu=0.02, uvwIn=(0.03, 0.01, -0.02), Uref=0.03)
    Fb = O.LBMBlock(16, 12, 12, dh=1.0, BndConds=(301,) * 6, flow=fl)
    Sb = O.LBMBlock(13, 9, 9, dh=0.5, xmin=5.0, ymin=4.0, zmin=4.0, BndConds=(0,) * 6, flow=fl)
    Fb.initialise(0.0); Sb.initialise(0.0)
    root = O.TreeNode(Fb); root.add_son(O.TreeNode(Sb), 2)
    for b in (Fb, Sb):
        b.update_volume_force(); b.set_boundary_conditions(); b.calculate_macro_quantities()
    for n in range(1, 13):
        O.set_blktime_all(root, float(n))
        O.tree_collision_streaming_IBM_FEM(root)
    for b in (Fb, Sb):
        b.calculate_macro_quantities()
        assert np.abs(b.den - 1.0).max() < 1e-14
        for k, v in enumerate(fl.uvwIn):
            assert np.abs(b.uuu[k] - v).max() < 1e-15
