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
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
