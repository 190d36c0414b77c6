"""Host orchestration of one step (block_comm.tree_collision_streaming_IBM_FEM, LBMBlockComm.f90:279-338) on CPU with a
recording stand-in for the block: what is issued, in which order, and when the postponed structural work actually runs."""
import tempfile

import numpy as np
import pytest

import fsilbm3d_b200 as F
from fsilbm3d_b200.block_comm import blockTreeNode, tree_collision_streaming_IBM_FEM
from tests.beam_cases import BOX, chain


class FakeBlock:
    """Duck-typed LBMBlock: records the calls and returns marker forces like the device would."""

    def __init__(self, log, name="root", dh=1.0, numsubstep=2):
        self.log, self.name, self.dh, self.blktime = log, name, dh, 0.0
        self.BndConds = (301,) * 6
        self.flow = F.FlowCondType(numsubstep=numsubstep)

    def set_blktime(self, t):
        self.blktime = t

    def update_volume_force(self):
        self.log.append((self.name, "update_volume_force"))

    def calculate_interaction_force(self, bodies, rootBC=None, collective=False):
        self.log.append((self.name, "ibm", len(bodies), collective))
        for b in bodies:
            b.v_Eforce[...] = 1e-4
        return 3

    def collide_stream(self):
        self.log.append((self.name, "collide_stream"))


class SplitFakeBlock(FakeBlock):
    """A block that offers the two-halved interaction force (fsilbm_ibm_interaction_force_begin / _wait)."""

    def calculate_interaction_force_begin(self, bodies, rootBC=None, collective=False):
        self.log.append((self.name, "ibm_begin", len(bodies), collective))
        self._bodies = bodies

    def calculate_interaction_force_wait(self):
        self.log.append((self.name, "ibm_wait"))
        for b in self._bodies:
            b.v_Eforce[...] = 1e-4
        return 4


def open_bodies(wd, n=2):
    from fsilbm3d_b200 import solid_solver as S
    import os
    S.write_plate_dat(os.path.join(wd, "plate.dat"), chain(5), 0.05, 0.05, (0.0, 1.0, 0.0), Nspan=2, material=(1.0e4, 4.0e3, 0.01, 1.0, 0.0, 1e-5, 8e-6, 2e-5))
    groups = [dict(fishNum=1, mesh="plate.dat", iBodyModel=2, iBodyType=1, isMotionGiven=(1,) * 6, firstXYZ=(0.3 + k, 0.2, 0.1)) for k in range(n)]
    with open(os.path.join(wd, "inFlow.dat"), "w") as f:
        f.write(S.inflow_text(UrefType=9, Uref=1.0, LrefType=1, Lref=1.0, isKB=2, blocks=[BOX], groups=groups, numsubstep=2, dtolFEM=1e-12))
    return S.SolidBodies("inFlow.dat", (301,) * 6, cwd=wd)


def test_structural_work_is_issued_behind_the_launch_and_runs_when_needed():
    sb = open_bodies(tempfile.mkdtemp(prefix="order_"))
    log = []
    calls = []
    real_advance = sb.advance
    sb.advance = lambda *a: (calls.append(len(log)), real_advance(*a))[1]      # remembers how far the log was when the beams were solved
    blk = FakeBlock(log)
    pos0 = sb.VBodies[0].pos.copy()
    it = tree_collision_streaming_IBM_FEM(blk, sb.plates, time=1.0)
    assert it == 3
    assert [e[1] for e in log] == ["update_volume_force", "ibm", "collide_stream"]
    assert calls == [] and len(sb._pending) == 1              # nothing solved yet: it waits for someone to need the result
    tree_collision_streaming_IBM_FEM(blk, sb.plates, time=2.0)
    # the postponed work of step 1 ran at the top of step 2 (before its interaction-force call), after step 1's launch
    assert calls == [4] and log[3] == ("root", "update_volume_force") and log[4][1] == "ibm"
    assert not np.array_equal(sb.VBodies[0].pos, pos0)        # reading the state flushes step 2's work too
    assert len(calls) == 2 and sb._pending == []
    sb.close()


def test_collective_call_without_bodies_and_without_solver():
    log = []
    blk = FakeBlock(log)
    assert tree_collision_streaming_IBM_FEM(blk, [], time=1.0) == 0
    assert [e[1] for e in log] == ["update_volume_force", "collide_stream"]          # no body, not collective: IBM skipped (Solidbody.f90:891)
    blk.ibm_collective = True
    log.clear()
    tree_collision_streaming_IBM_FEM(blk, [], time=2.0)
    assert log[1] == ("root", "ibm", 0, True)                                        # slab run with per-rank lists: every rank calls
    sb = open_bodies(tempfile.mkdtemp(prefix="order2_"), n=1)
    log.clear(); blk.ibm_collective = False
    lod0 = sb.VBodies[0].lodFlow.copy()
    tree_collision_streaming_IBM_FEM(blk, sb.plates, time=3.0, solver=False)
    assert sb._pending == [] and not np.array_equal(sb.VBodies[0].lodFlow, lod0)     # solver=False: loads only, at once
    sb.close()


def test_tree_order_with_a_son():
    """Root without bodies, one son carrying them: LBMBlockComm.f90:279-318 -- extract(1), root update, extract(2), then twice
    (son step, father->son), then son->father; the son's sub-step time advances by its own dh."""
    log = []

    class FakePair:
        def extract_interpolate_layer(self, t): log.append(("pair", "extract", t))
        def interpolation_father_to_son(self, n): log.append(("pair", "f2s", n))
        def deliver_son_to_father(self): log.append(("pair", "s2f"))

    root, son = FakeBlock(log, "root", dh=1.0), FakeBlock(log, "son", dh=0.5)
    node = blockTreeNode(root)
    node.sons.append(blockTreeNode(son)); node.comm.append(FakePair())
    F.set_blktime_all(node, 5.0)
    tree_collision_streaming_IBM_FEM(node)
    assert log == [("root", "update_volume_force"), ("pair", "extract", 1), ("root", "collide_stream"), ("pair", "extract", 2),
                   ("son", "update_volume_force"), ("son", "collide_stream"), ("pair", "f2s", 0),
                   ("son", "update_volume_force"), ("son", "collide_stream"), ("pair", "f2s", 1), ("pair", "s2f")]
    assert son.blktime == 5.5 and root.blktime == 5.0


def test_update_is_enqueued_before_the_host_waits_for_the_forces():
    """With the two-halved call the collide-stream launch is issued between _begin and _wait: the host never stands between
    the interaction force and the update that consumes it (the structural work still follows the forces)."""
    sb = open_bodies(tempfile.mkdtemp(prefix="order3_"), n=1)
    log = []
    blk = SplitFakeBlock(log)
    its = []
    assert tree_collision_streaming_IBM_FEM(blk, sb.plates, time=1.0, iters=its) == 4 and its == [4]
    assert [e[1] for e in log] == ["update_volume_force", "ibm_begin", "collide_stream", "ibm_wait"]
    assert len(sb._pending) == 1                      # loads + sub-steps postponed behind the wait, as before
    sb.close()
