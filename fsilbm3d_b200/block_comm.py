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
