"""Step orchestration of LBMBlockComm.f90:279-338 for a root block without sons, plus the x-slab
decomposition plumbing (one process per GPU; torch.distributed only hands the NCCL id around)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

from ._lib import check, ensure_init, lib


def slab_range(xDim: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Contiguous x-planes [offset, offset+count) of `rank`; remainders go to the last ranks, as
    OMPPrePartition does for threads (FluidDomain.f90:411-430)."""
    psize, residual = divmod(xDim, nranks)
    counts = [psize + (1 if p >= nranks - residual else 0) for p in range(nranks)]
    return sum(counts[:rank]), counts[rank]


# populations that cross an x-face of a slab: eₓ=+1 go to the right neighbour, eₓ=-1 to the left (ConstParams.f90:13)
UP_POPULATIONS = (1, 7, 9, 11, 13)
DOWN_POPULATIONS = (2, 8, 10, 12, 14)


def halo_plan(rank: int, nranks: int, periodic_x: bool):
    """Neighbours of `rank` in the x-slab decomposition and what is exchanged with each per step.
    Returns (left, right, UP_POPULATIONS, DOWN_POPULATIONS); a neighbour is -1 at a non-periodic domain end.
    After the local push, ghost plane X+1 holds the UP populations for `right` (they become its plane 1) and ghost
    plane 0 holds the DOWN populations for `left` (they become its plane X).  The same plan is hard-wired in
    csrc/fsilbm_api.cu (halo_exchange / halo_setup); tests/test_multi_cpu.py replays it with gloo on CPU slabs."""
    right = rank + 1 if rank + 1 < nranks else (0 if periodic_x else -1)
    left = rank - 1 if rank > 0 else (nranks - 1 if periodic_x else -1)
    return left, right, UP_POPULATIONS, DOWN_POPULATIONS


def ibm_box_participants(x0: int, length: int, slabs: Sequence[Tuple[int, int]], xDim: int):
    """Which ranks iterate a body whose stencil box covers global x-planes [x0, x0+length) (modulo xDim when it wraps), and
    what they send one another -- the host logic of fsilbm_ibm_interaction_force on slab runs (csrc/fsilbm_api.cu, "slab runs").
    `slabs[r] = (offset, count)` of rank r.  Returns (runs, leader): runs = [(rank, dx0, dx1), ...] are the maximal stretches of
    box planes dx in [dx0, dx1) owned by one rank, in box order; the ranks named there are the participants (one rank: no
    communication; several: every participant sends its stretches to every other participant and then iterates the whole
    box); leader = owner of the box's first plane, the rank that reports the body's residual to the loop control and its
    forces to the force exchange."""
    owner = [-1] * xDim
    for r, (off, cnt) in enumerate(slabs):
        for x in range(off, off + cnt):
            owner[x] = r
    if any(o < 0 for o in owner):
        raise ValueError("the slabs do not cover every x-plane")
    runs = []
    for dx in range(length):
        r = owner[(x0 + dx) % xDim]
        if runs and runs[-1][0] == r:
            runs[-1] = (r, runs[-1][1], dx + 1)
        else:
            runs.append((r, dx, dx + 1))
    return runs, runs[0][0]


def son_slab_plan(father_xmin: float, father_dh: float, slabs: Sequence[Tuple[int, int]], son_xmin: float, son_xDim: int, son_dh: float,
                  son_periodic_x: bool = False):
    """Where a son block goes on a slab run (include/fsilbm.h "Slab runs", fsilbm_pair_create / fsilbm_pair_create_remote).
    `slabs[r] = (offset, count)` of rank r's father slab.  Returns (owner, remote): `owner` creates the son block and the pair
    (the rank holding most of the footprint, the lower rank on a tie); `remote` lists the ranks -- directly left / right of the owner
    -- that must call fsilbm_pair_create_remote(father, owner) because the footprint reaches into their slabs.  The footprint is
    the father planes f(1)..f(2) of build_blocks_comunication (LBMBlockComm.f90:58-66).  Raises ValueError when it does not fit
    into the owner's slab and its two neighbours (a son that wide has to be split by the user)."""
    sD = son_xDim - (1 if son_periodic_x else 0)
    ratio = int(math.floor(father_dh / son_dh + 0.5))
    f0 = int(math.floor((son_xmin - father_xmin) / father_dh + 1.5)) - 1           # 0-based first father plane
    f1 = f0 + (sD - 1) // ratio
    owner_of = {}
    for r, (off, cnt) in enumerate(slabs):
        for x in range(off, off + cnt):
            owner_of[x] = r
    if f0 not in owner_of or f1 not in owner_of:
        raise ValueError(f"son block spans father planes {f0 + 1}..{f1 + 1}: not inside its father")
    counts = {}
    for x in range(f0, f1 + 1):
        counts[owner_of[x]] = counts.get(owner_of[x], 0) + 1
    owner = max(sorted(counts), key=lambda r: counts[r])
    remote = sorted(r for r in counts if r != owner)
    if any(abs(r - owner) != 1 for r in remote):
        raise ValueError(f"son block spans father planes {f0 + 1}..{f1 + 1} on ranks {sorted(counts)}: more than the owner's slab and its two neighbours")
    return owner, remote


def init_process_group(rank: int, nranks: int, device: int, broadcast_bytes) -> None:
    """Bind this process to `device` and join the library's NCCL communicator.
    broadcast_bytes(b: bytes|None) -> bytes must return rank 0's 128-byte id on every rank."""
    ensure_init(device)
    if nranks == 1:
        check(lib().fsilbm_comm_init(0, 1, None))
        return
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib().fsilbm_comm_unique_id(buf))
    raw = broadcast_bytes(buf.raw if rank == 0 else None)
    check(lib().fsilbm_comm_init(rank, nranks, raw))


class CommPair:
    """type CommPair (LBMBlockComm.f90:11-18) over the C ABI (fsilbm_pair_*)."""

    def __init__(self, father, son, interpolateScheme: int = 1):
        h = C.c_int(-1)
        check(lib().fsilbm_pair_create(father._h, son._h, interpolateScheme, C.byref(h)))
        self._h = h.value
        self.father, self.son = father, son
        out = (C.c_int * 36)()
        check(lib().fsilbm_pair_info(self._h, out))
        v = list(out)
        self.sds, self.s, self.f, self.si, self.fi = v[0:6], v[6:12], v[12:18], v[18:24], v[24:30]
        self.dimS, self.dimF = v[30:33], v[33:36]

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().fsilbm_pair_destroy(self._h)
            self._h = None

    def extract_interpolate_layer(self, time: int):
        check(lib().fsilbm_pair_extract_layer(self._h, time))

    def interpolation_father_to_son(self, n_timeStep: int):
        check(lib().fsilbm_pair_father_to_son(self._h, n_timeStep))

    def deliver_son_to_father(self):
        check(lib().fsilbm_pair_son_to_father(self._h))


class RemoteSon:
    """The neighbour's half of a son across a slab interface (fsilbm_pair_create_remote): rank `owner_rank`, directly left or
    right of this one, holds a son of `father` whose footprint reaches into this rank's slab."""

    def __init__(self, father, owner_rank: int):
        h = C.c_int(-1)
        check(lib().fsilbm_pair_create_remote(father._h, owner_rank, C.byref(h)))
        self._h = h.value
        self.father, self.owner_rank = father, owner_rank

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().fsilbm_pair_destroy(self._h)
            self._h = None


class blockTreeNode:
    """blockTreeNode (LBMBlockComm.f90:19-25): a block, the plates it carries (carriedBodies), its sons and CommPairs."""

    def __init__(self, block, plates: Sequence = ()):
        self.block, self.plates, self.sons, self.comm = block, list(plates), [], []

    def add_son(self, node: "blockTreeNode", interpolateScheme: int = 1) -> "blockTreeNode":
        self.sons.append(node)
        self.comm.append(CommPair(self.block, node.block, interpolateScheme))
        return node

    def walk(self):
        yield self
        for s in self.sons:
            yield from s.walk()


MachineTolerace = 1.0e-12   # ConstParams.f90:36 (name as spelt there)


def _extent(b):
    """xmin..zmax of a block incl. the periodic extension of FluidDomain.f90:97-105."""
    dims = (b.xDim, b.yDim, b.zDim)
    mins = (b.xmin, b.ymin, b.zmin)
    out = []
    for k in range(3):
        mx = mins[k] + b.dh * (dims[k] - 1)
        if b.BndConds[2 * k] == 301 and b.BndConds[2 * k + 1] == 301:
            mx = mx + b.dh
        out += [mins[k], mx]
    return out


def CompareBlocks(bi, bj) -> int:
    """FluidDomain.f90:1845-1972: 1 if block i contains j, -1 if i is inside j, 0 if separate or partially overlapping."""
    vi, vj = _extent(bi), _extent(bj)
    for (a, va, b, vb) in ((bi, vi, bj, vj), (bj, vj, bi, vi)):          # :1868-1902
        pos = [c == 0 for c in a.BndConds]
        if sum(pos) == 1:
            for p in range(6):
                if not pos[p]:
                    continue
                if p % 2 == 0:
                    if b.BndConds[p + 1] == 1 and abs(vb[p + 1] - va[p] - b.dh) < MachineTolerace:
                        va[p + 1] = va[p]
                else:
                    if b.BndConds[p - 1] == 1 and abs(va[p] - vb[p - 1] - b.dh) < MachineTolerace:
                        va[p - 1] = va[p]
    cnt = align = 0
    for k in range(3):                                                  # :1903-1953
        lo, hi = 2 * k, 2 * k + 1
        d1 = vi[lo] < vj[lo] or abs(vi[lo] - vj[lo]) < MachineTolerace
        d2 = vj[hi] < vi[hi] or abs(vj[hi] - vi[hi]) < MachineTolerace
        d3 = vj[lo] < vi[lo] or abs(vj[lo] - vi[lo]) < MachineTolerace
        d4 = vi[hi] < vj[hi] or abs(vi[hi] - vj[hi]) < MachineTolerace
        d5, d6 = vi[hi] < vj[lo], vj[hi] < vi[lo]
        if d1 and d2 and not (d3 and d4):
            cnt += 1
        elif d3 and d4 and not (d1 and d2):
            cnt -= 1
        elif d1 and d2 and d3 and d4:
            align += 1
        elif d5 or d6:
            return 0
    if align > 0:
        if cnt < 0:
            cnt -= align
        if cnt > 0:
            cnt += align
    return 1 if cnt == 3 else (-1 if cnt == -3 else 0)


def build_block_tree(blocks: Sequence, interpolateScheme: int = 1) -> blockTreeNode:
    """build_block_tree (LBMBlockComm.f90:195-211): unique root by containment (findremove_blockTreeRoot :98-133),
    the rest nested by array_to_tree (:135-193), one CommPair per father/son (build_blocks_comunication :32-96)."""
    n = len(blocks)
    nodes = [blockTreeNode(b) for b in blocks]

    def nest(ids, root):
        if not ids:
            return
        fa = {i: None for i in ids}
        for a in range(len(ids) - 1):
            for b in range(a + 1, len(ids)):
                c = CompareBlocks(blocks[ids[a]], blocks[ids[b]])
                if c == 1:
                    fa[ids[b]] = ids[a]
                elif c == -1:
                    fa[ids[a]] = ids[b]
        changed = True
        while changed:                       # lift every block to its outermost container within this level (:154-165)
            changed = False
            for i in ids:
                if fa[i] is not None and fa[fa[i]] is not None and fa[i] != fa[fa[i]]:
                    fa[i] = fa[fa[i]]
                    changed = True
        roots = [i for i in ids if fa[i] is None]
        for r in roots:
            nodes[root].add_son(nodes[r], interpolateScheme)
            nest([j for j in ids if fa[j] == r], r)

    ids = list(range(n))
    fa = {i: None for i in ids}
    for a in range(n - 1):
        for b in range(a + 1, n):
            c = CompareBlocks(blocks[a], blocks[b])
            if c == 1:
                fa[b] = a
            elif c == -1:
                fa[a] = b
    roots = [i for i in ids if fa[i] is None]
    if len(roots) != 1:
        raise ValueError("Error: there exist more than one block tree root")   # LBMBlockComm.f90:129-132
    nest([i for i in ids if i != roots[0]], roots[0])
    return nodes[roots[0]]


def find_carrier_fluidblock(blocks: Sequence, x) -> int:
    """FluidDomain.f90:1974-1996: index of the finest block containing point x."""
    best, dh = -1, 1e10
    for i, b in enumerate(blocks):
        e = _extent(b)
        if all(e[2 * k] <= x[k] <= e[2 * k + 1] for k in range(3)) and b.dh < dh:
            best, dh = i, b.dh
    if best < 0:
        raise ValueError(f"Error: carrier fluid block not found {x}")
    return best


def set_blktime_all(node: blockTreeNode, time: float) -> None:
    """LBMblks(:)%blktime = time, main.f90:97."""
    for nd in node.walk():
        nd.block.set_blktime(time)


def tree_collision_streaming_IBM_FEM(node, plates: Sequence = (), time: Optional[float] = None,
                                     rootBC=None, solver: bool = True, iters: Optional[list] = None) -> int:
    """tree_collision_streaming_IBM_FEM (LBMBlockComm.f90:279-318).  `node` is a blockTreeNode, or a bare LBMBlock
    (then `plates` are the bodies it carries and there are no sons).  `plates` are RigidPlate-like objects
    (UpdatePosVelArea(), structure(), .body).  Returns iterLBM of this node's block."""
    if not isinstance(node, blockTreeNode):
        node = blockTreeNode(node, plates)
    block, plates = node.block, node.plates
    if time is not None:
        block.set_blktime(time)
    rootBC = block.BndConds if rootBC is None else rootBC
    block.update_volume_force()                                             # :283
    it = 0
    collective = getattr(block, "ibm_collective", False)   # slab runs with per-rank body lists: every rank calls, even with no body
    ibm = len(plates) or collective
    split = ibm and hasattr(block, "calculate_interaction_force_begin")
    if ibm:                                                                 # IBM_FEM, :287 -> :320-338
        for p in plates:
            p.UpdatePosVelArea()                                            # Solidbody.f90:597-600
        if split:   # enqueue only: the update below is issued before the host waits for the marker forces
            block.calculate_interaction_force_begin([p.body for p in plates], rootBC, collective=collective)   # :601
        else:
            it = block.calculate_interaction_force([p.body for p in plates], rootBC, collective=collective)
    for pair in node.comm:
        pair.extract_interpolate_layer(1)                                   # :290
    block.collide_stream()                                                  # :285-303 fused; asynchronous launch (follows the IBM on the device)
    if split:
        it = block.calculate_interaction_force_wait()                       # marker forces + iterLBM back on the host
    if iters is not None:
        iters.append(it)
    # The host halves of FSInteraction_force -- nodal loads (Solidbody.f90:911,945-967) and Solver (:333-335) -- only read
    # the marker forces just returned and touch no fluid state.  The reference runs them before the collision; here they
    # are issued after the launch so that they overlap the device's collide-stream of the same step.
    owner = getattr(plates[0], "owner", None) if len(plates) else None
    if len(plates) and solver and owner is not None and all(getattr(p, "owner", None) is owner for p in plates):
        # ... and postponed until the markers are next needed (the top of these bodies' next step), so that device work queued in
        # between -- a son's transfers, the father's next collide-stream -- also runs while the beams are solved
        owner.advance_later([p.body.index for p in plates], block.blktime, block.flow.numsubstep, block.dh)
    elif len(plates):
        for p in plates:
            if hasattr(p, "FluidVolumeForce"):
                p.FluidVolumeForce()
        if solver:
            nsub = block.flow.numsubstep
            dt_solid = block.dh / float(nsub)
            for isub in range(1, nsub + 1):
                for p in plates:
                    p.structure(block.blktime, isub, block.dh, dt_solid)
    for pair in node.comm:
        pair.extract_interpolate_layer(2)                                   # :305
    for son, pair in zip(node.sons, node.comm):                             # :307-317
        for n_timeStep in range(2):
            son.block.set_blktime(son.block.blktime + float(n_timeStep) * son.block.dh)   # :311
            tree_collision_streaming_IBM_FEM(son, rootBC=rootBC, solver=solver, iters=iters)
            pair.interpolation_father_to_son(n_timeStep)
        pair.deliver_son_to_father()
    return it
