"""Parity at the sizes and step counts BASELINE.json names (SURVEY 8(d) "Parity tolerance"): the CUDA path through the C
ABI against the CPU oracle on the same seeded state, compared bit for bit on the populations (array_equal), with
north_star's tolerances (1e-12 relative on density and velocity, 1e-10 on marker forces) asserted as well and the
IBM iteration counts required to be equal every step.

  channel256   configs[1]: 256^3 periodic body-force channel, SRT, 200 steps
  plate512     configs[2]: 512x256x256, shear inflow, moving walls, rigid plate of 8 192 markers, 50 steps
  two plates   32 768 markers in two bodies apart in x, so the ordered plane list of collide_stream holds two groups of
               planes around bodies and three groups of remaining planes (before / between / after)

The oracle runs on all host cores (about 0.3 s per step per 16.8 M cells); the whole module takes a few minutes."""
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
    F.lib()
    return F


def all_cores(oracle):
    import os
    oracle.lib().orc_omp_set_threads(len(os.sched_getaffinity(0)))


def check_fluid(ob, gb):
    from tests.common import rel_err
    ob.calculate_macro_quantities()
    den, uuu = gb.download_macro()
    e_den, e_u = rel_err(den, ob.den), rel_err(uuu, ob.uuu)
    assert e_den <= TOL_FLUID and e_u <= TOL_FLUID, (e_den, e_u)
    del den, uuu
    f = gb.download_fIn()
    exact = np.array_equal(f, ob.fIn)
    assert exact, f"populations not bit-exact: max |df| = {np.max(np.abs(f - ob.fIn))}"


def test_channel256_200_steps(oracle, F):
    """configs[1] at full size (main.f90:93-108 loop; collision FluidDomain.f90:1208-1263, streaming :1514-1625)."""
    from tests.common import make_pair
    all_cores(oracle)
    ob, gb = make_pair(oracle, F, (256, 256, 256), model=1, nu=0.1, volumeForceIn=(1e-6, 0.0, 0.0))
    for n in range(1, 201):
        ob.set_blktime(float(n)); gb.set_blktime(float(n))
        ob.step(); gb.step()
    check_fluid(ob, gb)
    gb.close()


def run_plates(oracle, F, dims, bc, flow, dh, plates_kw, steps, ntol):
    from tests.common import make_pair
    all_cores(oracle)
    ob, gb = make_pair(oracle, F, dims, BndConds=bc, dh=dh, **flow)
    pgs = [F.RigidPlate(**kw) for kw in plates_kw]
    ovs = []
    for pg in pgs:
        ov = oracle.VirtualBody(pg.body.v_nelmts, v_move=pg.body.v_move, iBodyModel=1)
        ov.v_Exyz[...] = pg.body.v_Exyz; ov.v_Evel[...] = pg.body.v_Evel; ov.v_Ea[...] = pg.body.v_Ea
        ovs.append(ov)
    e0 = F.lib().fsilbm_ibm_early_count()
    for n in range(1, steps + 1):
        t = n * dh
        ob.set_blktime(t)
        it_o = ob.step(ovs)
        it_g = F.tree_collision_streaming_IBM_FEM(gb, pgs, time=t)
        assert it_o == it_g == ntol, (n, it_o, it_g)
        for pg, ov in zip(pgs, ovs):
            assert np.array_equal(pg.body.v_Eforce, ov.v_Eforce), f"marker forces differ at step {n}"
    for k, (pg, ov) in enumerate(zip(pgs, ovs)):
        Ei, Ew = gb.download_stencil(k, pg.body.v_nelmts)
        assert np.array_equal(Ei, ov.v_Ei) and np.array_equal(Ew, ov.v_Ew)
        assert abs(pg.body.v_Eforce[:, 0].sum()) > 1e-10
    overlapped = F.lib().fsilbm_ibm_early_count() - e0
    check_fluid(ob, gb)
    gb.close()
    return overlapped


def test_plate512_50_steps(oracle, F):
    """configs[2] at full size: the plate, flow and boundary codes of bench.py's default workload (LBMBlockComm.f90:320-338,
    Solidbody.f90:869-918 on 8 192 markers, five IBM iterations)."""
    dh = 1.0 / 64.0
    gamma = 0.02 / ((256 - 1) * dh)
    flow = dict(nu=5e-4, uvwIn=(0.05, 0.0, 0.0), shearRateIn=(0.0, gamma, 0.0), Uref=0.05, ntolLBM=5, dtolLBM=1e-30)
    plate = dict(origin=(3.0, 2.0 - 0.013, 1.0 + 0.003), nEL=64, len1=dh, Nspan=128, spanlen=2.0, Lspan=0.0,
                 chord_dir=(1.0, 0.0, 0.0), span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha=1.0, denIn=1.0)
    overlapped = run_plates(oracle, F, (512, 256, 256), (101, 104, 202, 202, 301, 301), flow, dh, [plate], 50, 5)
    assert overlapped >= 45   # the interaction-force calls ran beside the update of the planes away from the body


def test_32768_markers_two_bodies_three_plane_ranges(oracle, F):
    """Two plates of 16 384 markers each (64 x 256), apart in x and inclined, in a 224x96x288 block: box merging and the
    per-cell ordered gather at 32 768 markers, int16 stencil indices up to 287, and a plane list with two groups of planes
    around bodies and three groups of remaining planes."""
    flow = dict(nu=0.02, uvwIn=(0.04, 0.0, 0.0), shearRateIn=(0.0, 1e-4, 0.0), Uref=0.04, ntolLBM=3, dtolLBM=1e-30)
    common = dict(nEL=64, len1=1.0, Nspan=256, spanlen=256.0, Lspan=0.0, span_dir=(0.0, 0.0, 1.0), IBPenaltyAlpha=1.0, denIn=1.0)
    plates = [dict(common, origin=(28.3, 40.2, 15.4), chord_dir=(1.0, 0.25, 0.0)),
              dict(common, origin=(130.6, 52.7, 16.1), chord_dir=(1.0, -0.3, 0.0))]
    overlapped = run_plates(oracle, F, (224, 96, 288), (101, 104, 202, 202, 301, 301), flow, 1.0, plates, 30, 3)
    assert overlapped >= 25
