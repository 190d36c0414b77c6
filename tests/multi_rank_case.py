"""Run under torchrun (one process per GPU): x-slab decomposition of a block with boundary conditions and an
immersed plate that straddles slab interfaces, compared on rank 0 with the single-block CPU oracle.
Exit code 0 = parity within north_star's tolerances."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    torch.cuda.set_device(local)
    import fsilbm3d_b200 as F
    from tests.common import perturbed_state, rel_err

    def bcast(b):
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.tensor(list(b), dtype=torch.uint8)
        dist.broadcast(t, src=0)
        return bytes(t.tolist())
    F.init_process_group(rank, world, local, bcast)

    cases = [
        dict(name="periodic_x", dims=(8 * world + 3, 14, 40), bc=(301, 301, 203, 203, 301, 301), plate_origin=None,
             flow=dict(nu=0.05, volumeForceIn=(1e-6, 0.0, 2e-7))),
        dict(name="inlet_outlet_plate", dims=(10 * world + 1, 20, 24), bc=(102, 104, 202, 202, 301, 301), plate_origin="mid",
             flow=dict(nu=0.05, uvwIn=(0.04, 0.0, 0.0), shearRateIn=(0.0, 3e-4, 0.0), Uref=0.04, ntolLBM=3, dtolLBM=1e-30)),
        dict(name="periodic_plate_wrap", dims=(9 * world, 18, 20), bc=(301,) * 6, plate_origin="wrap",
             flow=dict(nu=0.05, uvwIn=(0.03, 0.0, 0.0), Uref=0.03, ntolLBM=4, dtolLBM=1e-30)),
    ]
    F._lib.check(F.lib().fsilbm_set_option(b"halo_timeout_s", 30))
    ok = True
    for halo_mode, case in [(m, c) for m in (1, 0) for c in cases]:
        F._lib.check(F.lib().fsilbm_set_option(b"halo", halo_mode))
        X, Y, Z = case["dims"]
        off, cnt = F.slab_range(X, rank, world)
        flow = F.FlowCondType(**case["flow"])
        gb = F.LBMBlock(X, Y, Z, BndConds=case["bc"], flow=flow, xOffset=off, xLocal=cnt, device=local)
        gb.initialise(0.0)
        f0 = perturbed_state((X, Y, Z), flow)
        gb.upload_fIn(np.ascontiguousarray(f0[:, off:off + cnt]))
        gb.update_volume_force(); gb.set_boundary_conditions()
        plates = []
        if case["plate_origin"]:
            ox = X / 2.0 - 4.2 if case["plate_origin"] == "mid" else X - 3.3
            plates = [F.RigidPlate(origin=(ox, 8.3, 5.2), nEL=8, len1=1.0, Nspan=8, spanlen=8.0, Lspan=0.0, chord_dir=(1.0, 0.2, 0.0), denIn=1.0)]
        if rank == 0:
            from oracle import oracle as O
            of = O.Flow(**case["flow"])
            ob = O.LBMBlock(X, Y, Z, BndConds=case["bc"], flow=of)
            ob.initialise(0.0)
            ob.fIn[...] = f0
            ob.update_volume_force(); ob.set_boundary_conditions(); ob.calculate_macro_quantities()
            ovs = []
            for p in plates:
                ov = O.VirtualBody(p.body.v_nelmts, v_move=0, iBodyModel=1)
                ov.v_Exyz[...] = p.body.v_Exyz; ov.v_Evel[...] = p.body.v_Evel; ov.v_Ea[...] = p.body.v_Ea
                ovs.append(ov)
        nsteps = 25
        eF = 0.0
        for n in range(1, nsteps + 1):
            it_g = F.tree_collision_streaming_IBM_FEM(gb, plates, time=float(n), solver=False)
            if rank == 0:
                ob.set_blktime(float(n))
                it_o = ob.step(ovs)
                if plates:
                    ok &= (it_o == it_g)
                    eF = max(eF, rel_err(plates[0].body.v_Eforce, ovs[0].v_Eforce))
        den, uuu = gb.download_macro()
        floc = gb.download_fIn()
        # gather slabs on rank 0
        parts = [None] * world
        dist.gather_object((off, cnt, den, uuu, floc), parts if rank == 0 else None, dst=0)
        if rank == 0:
            ob.calculate_macro_quantities()
            DEN = np.concatenate([p[2] for p in parts], axis=0)
            UUU = np.concatenate([p[3] for p in parts], axis=1)
            FF = np.concatenate([p[4] for p in parts], axis=1)
            e_den, e_u, e_f = rel_err(DEN, ob.den), rel_err(UUU, ob.uuu), rel_err(FF, ob.fIn)
            exact = bool(np.array_equal(FF, ob.fIn))
            good = e_den <= 1e-12 and e_u <= 1e-12 and e_f <= 1e-12 and eF <= 1e-10
            ok &= good
            print(f"[multi x{world}] halo={gb.halo_transport!r} {case['name']}: rel err den {e_den:.2e} u {e_u:.2e} f {e_f:.2e} force {eF:.2e} bit-exact {exact} -> {'OK' if good else 'FAIL'}", flush=True)
        gb.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
