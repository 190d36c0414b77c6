"""N>1 host logic on CPU: the slab decomposition plan replayed over gloo with world_size 2 and 3."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29612)])
def test_slab_plan_over_gloo(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gloo_slab_case.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("bit-exact True") == 3      # two fluid cases + IBM on slabs


def test_halo_plan_neighbours():
    from fsilbm3d_b200.block_comm import halo_plan
    assert halo_plan(0, 1, True)[:2] == (0, 0)
    assert halo_plan(0, 1, False)[:2] == (-1, -1)
    assert halo_plan(0, 2, True)[:2] == (1, 1)
    assert halo_plan(0, 8, False)[:2] == (-1, 1) and halo_plan(7, 8, False)[:2] == (6, -1)
    assert halo_plan(7, 8, True)[:2] == (6, 0)
    _, _, up, dn = halo_plan(3, 8, True)
    ex = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
    assert [q for q in range(19) if ex[q] == 1] == list(up) and [q for q in range(19) if ex[q] == -1] == list(dn)


def test_ibm_box_participants():
    """Slab runs: which ranks iterate a body and what they exchange (the rule fsilbm_ibm_interaction_force applies)."""
    import fsilbm3d_b200 as F
    slabs = [F.slab_range(40, r, 4) for r in range(4)]                      # 4 x 10 planes
    assert F.ibm_box_participants(12, 6, slabs, 40) == ([(1, 0, 6)], 1)         # inside slab 1: no communication
    assert F.ibm_box_participants(17, 8, slabs, 40) == ([(1, 0, 3), (2, 3, 8)], 1)   # across the 1|2 interface, led by rank 1
    assert F.ibm_box_participants(36, 9, slabs, 40) == ([(3, 0, 4), (0, 4, 9)], 3)   # periodic wrap: ranks 3 and 0
    runs, lead = F.ibm_box_participants(8, 24, slabs, 40)                       # a long body over four slabs
    assert [r for r, _, _ in runs] == [0, 1, 2, 3] and lead == 0 and runs[-1] == (3, 22, 24)
    two = [(0, 7), (7, 5)]                                                    # two ranks, periodic: the box re-enters rank 0
    assert F.ibm_box_participants(5, 10, two, 12) == ([(0, 0, 2), (1, 2, 7), (0, 7, 10)], 0)
    with pytest.raises(ValueError):
        F.ibm_box_participants(0, 3, [(0, 5)], 8)
