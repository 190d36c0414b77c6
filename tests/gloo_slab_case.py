"""Run under torchrun with the gloo backend (CPU, world_size >= 2): replays the product's x-slab decomposition plan
(fsilbm3d_b200.slab_range + halo_plan: which populations cross which face, who the neighbours are, ghost-plane
layout) on CPU slabs stepped by the oracle, exchanging the halo planes through torch.distributed, and compares
the assembled field on rank 0 with the single-block oracle.  Exit code 0 = bit-exact."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from fsilbm3d_b200.block_comm import halo_plan, slab_range
    from oracle import oracle as O
    from tests.common import perturbed_state

    ok = True
    for X, Y, Z, bc, model in [(4 * world + 3, 6, 10, (301, 301, 203, 203, 301, 301), 1),
                               (3 * world + 1, 5, 7, (301, 301, 301, 301, 201, 302), 2)]:
        flow = O.Flow(nu=0.05, volumeForceIn=(1e-6, 0.0, 2e-7))
        off, cnt = slab_range(X, rank, world)
        left, right, UP, DN = halo_plan(rank, world, periodic_x=True)
        # local slab with two ghost planes, stepped by the oracle (its own x-wrap only ever touches the ghosts'
        # outward populations, which the exchange overwrites)
        sb = O.LBMBlock(cnt + 2, Y, Z, BndConds=bc, iCollidModel=model, params=(0.2,) + (0.0,) * 9, flow=flow, npsize=1)
        sb.initialise(0.0)
        f0 = perturbed_state((X, Y, Z), flow)
        sb.fIn[:, 1:cnt + 1] = f0[:, off:off + cnt]
        sb.fIn[:, 0] = f0[:, (off - 1) % X]
        sb.fIn[:, cnt + 1] = f0[:, (off + cnt) % X]
        sb.update_volume_force(); sb.set_boundary_conditions()
        if rank == 0:
            ob = O.LBMBlock(X, Y, Z, BndConds=bc, iCollidModel=model, params=(0.2,) + (0.0,) * 9, flow=flow, npsize=1)
            ob.initialise(0.0)
            ob.fIn[...] = f0
            ob.update_volume_force(); ob.set_boundary_conditions()
        for n in range(1, 9):
            sb.set_blktime(float(n))
            sb.step()
            # ghost plane cnt+1 holds what left through the right face; ghost plane 0 what left through the left face
            send_r = torch.from_numpy(np.ascontiguousarray(sb.fIn[list(UP), cnt + 1]))
            send_l = torch.from_numpy(np.ascontiguousarray(sb.fIn[list(DN), 0]))
            recv_l, recv_r = torch.empty_like(send_r), torch.empty_like(send_l)
            reqs = [dist.isend(send_r, right, tag=1), dist.isend(send_l, left, tag=2),
                    dist.irecv(recv_l, left, tag=1), dist.irecv(recv_r, right, tag=2)]
            for r in reqs:
                r.wait()
            sb.fIn[list(UP), 1] = recv_l.numpy()
            sb.fIn[list(DN), cnt] = recv_r.numpy()
            # the face rules of the y/z faces act on the streamed field, i.e. after the halo has arrived
            sb.set_boundary_conditions()
            if rank == 0:
                ob.set_blktime(float(n))
                ob.step()
        parts = [None] * world
        dist.gather_object((off, cnt, np.array(sb.fIn[:, 1:cnt + 1])), parts if rank == 0 else None, dst=0)
        if rank == 0:
            FF = np.concatenate([p[2] for p in parts], axis=1)
            exact = bool(np.array_equal(FF, ob.fIn))
            print(f"[gloo x{world}] {X}x{Y}x{Z} bc={bc} model={model}: bit-exact {exact}", flush=True)
            ok &= exact
    ok &= ibm_on_slabs(dist, rank, world)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def ibm_on_slabs(dist, rank, world):
    """The product's IBM protocol on slab runs (csrc/fsilbm_api.cu "slab runs"; block_comm.ibm_box_participants), replayed on
    CPU slabs: a plate across the last slab interface is iterated only by the ranks that own planes of its stencil box; they
    send one another the box velocities of the planes they own and then run the whole penalty iteration redundantly; each keeps
    its own planes of the corrected velocity and of the force; the leader reports the marker forces.  Bit-exact against the
    single-block oracle on rank 0."""
    import ctypes as C
    import fsilbm3d_b200 as F
    from fsilbm3d_b200.block_comm import halo_plan, ibm_box_participants, slab_range
    from oracle import oracle as O
    from tests.common import perturbed_state
    X, Y, Z = 8 * world + 2, 14, 16
    bc = (301,) * 6
    flow = O.Flow(nu=0.05, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, ntolLBM=4, dtolLBM=1e-30)
    slabs = [slab_range(X, r, world) for r in range(world)]
    off, cnt = slabs[rank]
    left, right, UP, DN = halo_plan(rank, world, periodic_x=True)
    x_if = slabs[world - 1][0]                               # first plane of the last slab
    plate = F.RigidPlate(origin=(x_if - 2.7, 6.3, 4.2), nEL=5, len1=1.0, Nspan=6, spanlen=6.0, Lspan=0.0, chord_dir=(1.0, 0.2, 0.0), denIn=1.0)

    def oracle_body():
        ov = O.VirtualBody(plate.body.v_nelmts, v_move=0, iBodyModel=1)
        ov.v_Exyz[...] = plate.body.v_Exyz; ov.v_Evel[...] = plate.body.v_Evel; ov.v_Ea[...] = plate.body.v_Ea
        return ov
    ov = oracle_body()
    # the box of the stencils in x: base index i = floor(x/dh) (0-based), stencil i-1..i+2, one guard plane each side
    ix = np.floor(plate.body.v_Exyz[:, 0]).astype(int)
    x0, length = int(ix.min()) - 2, int(ix.max() + 3) - (int(ix.min()) - 2) + 1
    runs, leader = ibm_box_participants(x0 % X, length, slabs, X)
    parts = sorted({r for r, _, _ in runs})
    assert len(parts) == 2 and (world < 3 or 0 not in parts)     # across one interface; with three ranks the first holds no body

    sb = O.LBMBlock(cnt + 2, Y, Z, BndConds=bc, flow=flow, npsize=1)
    sb.initialise(0.0)
    f0 = perturbed_state((X, Y, Z), flow)
    sb.fIn[:, 1:cnt + 1] = f0[:, off:off + cnt]
    sb.fIn[:, 0] = f0[:, (off - 1) % X]
    sb.fIn[:, cnt + 1] = f0[:, (off + cnt) % X]
    sb.update_volume_force(); sb.set_boundary_conditions()
    if rank == 0:
        ob = O.LBMBlock(X, Y, Z, BndConds=bc, flow=flow, npsize=1)
        ob.initialise(0.0)
        ob.fIn[...] = f0
        ob.update_volume_force(); ob.set_boundary_conditions()
        ovb = oracle_body()
    U, Fo = np.zeros((3, X, Y, Z)), np.zeros((3, X, Y, Z))       # scratch with the block's own index space (same stencil arithmetic)
    dp = C.POINTER(C.c_double)
    ok = True
    for n in range(1, 7):
        sb.set_blktime(float(n))
        sb.update_volume_force(); sb.calculate_macro_quantities(); sb.ResetVolumeForce()      # LBMBlockComm.f90:283-286
        it = -1
        if rank in parts:
            U[...] = 0.0; Fo[...] = 0.0
            reqs, keep = [], []
            for (r, d0, d1) in runs:
                gx = [(x0 + d) % X for d in range(d0, d1)]
                if r == rank:
                    U[:, gx] = sb.uuu[:, [g - off + 1 for g in gx]]
                    t = torch.from_numpy(np.ascontiguousarray(U[:, gx]))
                    reqs += [dist.isend(t, p, tag=10 + d0) for p in parts if p != rank]
                else:
                    t = torch.empty((3, len(gx), Y, Z), dtype=torch.float64)
                    reqs.append(dist.irecv(t, r, tag=10 + d0)); keep.append((gx, t))
            for q in reqs:
                q.wait()
            for gx, t in keep:
                U[:, gx] = t.numpy()
            arr = (C.c_void_p * 1)(ov._h)
            it = O.lib().orc_calculate_interaction_force(arr, 1, 1.0, 1.0, 0.0, 0.0, 0.0, X, Y, Z, U.ctypes.data_as(dp), Fo.ctypes.data_as(dp),
                                                         (C.c_int * 6)(*bc), flow.denIn, flow.Uref, flow.ntolLBM, flow.dtolLBM)
            for (r, d0, d1) in runs:                         # each participant keeps its own planes of the result
                if r == rank:
                    gx = [(x0 + d) % X for d in range(d0, d1)]
                    loc = [g - off + 1 for g in gx]
                    sb.uuu[:, loc] = U[:, gx]
                    sb.force[:, loc] = Fo[:, gx]
        sb.add_volume_force(); sb.collision(); sb.halfwayBCset(); sb.streaming()               # :288-299
        send_r = torch.from_numpy(np.ascontiguousarray(sb.fIn[list(UP), cnt + 1]))
        send_l = torch.from_numpy(np.ascontiguousarray(sb.fIn[list(DN), 0]))
        recv_l, recv_r = torch.empty_like(send_r), torch.empty_like(send_l)
        reqs = [dist.isend(send_r, right, tag=1), dist.isend(send_l, left, tag=2), dist.irecv(recv_l, left, tag=1), dist.irecv(recv_r, right, tag=2)]
        for q in reqs:
            q.wait()
        sb.fIn[list(UP), 1] = recv_l.numpy()
        sb.fIn[list(DN), cnt] = recv_r.numpy()
        sb.set_boundary_conditions()                                                             # :303
        forces = [None] * world
        dist.gather_object((it, np.array(ov.v_Eforce)) if rank == leader else None, forces if rank == 0 else None, dst=0)
        if rank == 0:
            ob.set_blktime(float(n))
            it_o = ob.step([ovb])
            it_l, f_l = forces[leader]
            ok &= it_o == it_l and bool(np.array_equal(f_l, ovb.v_Eforce))
    gathered = [None] * world
    dist.gather_object(np.array(sb.fIn[:, 1:cnt + 1]), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        exact = bool(np.array_equal(np.concatenate(gathered, axis=1), ob.fIn)) and ok and float(np.abs(ovb.v_Eforce).max()) > 0.0
        print(f"[gloo x{world}] IBM on slabs, plate across the interface of ranks {parts} (leader {leader}): bit-exact {exact}", flush=True)
        return exact
    return True


if __name__ == "__main__":
    main()
