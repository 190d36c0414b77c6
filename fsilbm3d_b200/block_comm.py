"""Step orchestration of LBMBlockComm.f90:279-338 for a root block without sons, plus the x-slab
decomposition plumbing (one process per GPU; torch.distributed only hands the NCCL id around)."""
from __future__ import annotations

import ctypes as C
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


def tree_collision_streaming_IBM_FEM(block, plates: Sequence = (), time: Optional[float] = None,
                                     rootBC=None, solver: bool = True) -> int:
    """One pass of tree_collision_streaming_IBM_FEM (LBMBlockComm.f90:279-305) on a block without sons.
    `plates` are RigidPlate-like objects (UpdatePosVelArea(), structure(), .body).  Returns iterLBM."""
    if time is not None:
        block.set_blktime(time)
    block.update_volume_force()                                             # :283
    it = 0
    if len(plates):
        for p in plates:
            p.UpdatePosVelArea()                                            # Solidbody.f90:597-600
        it = block.calculate_interaction_force([p.body for p in plates], rootBC)   # :601
        if solver:                                                          # IBM_FEM, LBMBlockComm.f90:333-335
            nsub = block.flow.numsubstep
            dt_solid = block.dh / float(nsub)
            for isub in range(1, nsub + 1):
                for p in plates:
                    p.structure(block.blktime, isub, block.dh, dt_solid)
    block.collide_stream()                                                  # :285-303 fused
    return it
