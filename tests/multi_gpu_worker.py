"""Multi-rank parity worker (run under torchrun, one process per GPU):
every rank owns one z-slab of the same image; the distributed result must equal
the single-GPU result (integers bit-exact, tau to 1e-9 relative) and the oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openimpala_b200 import capi, synth  # noqa: E402
from openimpala_b200.tortuosity import tau_from_fluxes  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm = capi.Comm(rank, world, idt.cpu().numpy().tobytes(), device=local)
    ok = True
    cases = [((96, 64, 80), 2, 6), ((96, 64, 80), 0, 6), ((64, 100, 36), 1, 5), ((128, 128, 128), 2, 8)]
    # every case through the default halo path (peer memory when the ranks can map
    # each other, which is required on an NVLink box), the first two also through
    # NCCL send/recv: both must give the single-GPU answer
    runs = [(c, capi.OI_HALO_AUTO) for c in cases] + [(c, capi.OI_HALO_NCCL) for c in cases[:2]]
    want_peer = os.environ.get("OI_HALO_MODE", "") != "nccl"
    for (shape, direction, radius), halo_mode in runs:
        full = synth.sphere_packing_slab(shape, seed=11, radius=radius, solid_target=0.5)
        z0, nzl = capi.slab_partition(shape[0], world)[rank]
        slab = np.ascontiguousarray(full[z0:z0 + nzl])
        s = capi.Solver(shape, direction, 1, -1.0, 1.0, device=local, z_begin=z0, nz_local=nzl, comm=comm,
                        halo_mode=halo_mode)
        mode_used, _ = s.halo_info()
        if halo_mode == capi.OI_HALO_AUTO and want_peer and mode_used != capi.OI_HALO_PEER:
            print(f"rank {rank}: peer halo path not active (mode {mode_used})", flush=True)
            ok = False
        s.set_phase(slab)
        pc, tc = s.volume_fraction()
        n_active = s.build_mask()
        mask = s.mask()
        chk = s.check_matrix_properties()
        info = s.solve()
        fin, fout, ni, no = s.fluxes()
        tau, _, _ = tau_from_fluxes(fin, fout, n_active / full.size, float(shape[2 - direction]),
                                    float(full.size / shape[2 - direction]), -1.0, 1.0)
        n_peer = s.halo_info()[1]
        s.close()
        if rank == 0:
            r = capi.Solver(shape, direction, 1, -1.0, 1.0, device=local)
            r.set_phase(full)
            pc1, tc1 = r.volume_fraction()
            n1 = r.build_mask()
            mask1 = r.mask()
            info1 = r.solve()
            fin1, fout1, ni1, no1 = r.fluxes()
            tau1, _, _ = tau_from_fluxes(fin1, fout1, n1 / full.size, float(shape[2 - direction]),
                                         float(full.size / shape[2 - direction]), -1.0, 1.0)
            r.close()
            good = ((pc, tc, n_active, ni, no) == (pc1, tc1, n1, ni1, no1) and chk and
                    np.array_equal(mask, mask1[z0:z0 + nzl]) and info.converged and
                    abs(tau - tau1) <= 1e-8 * abs(tau1))
            print(f"case {shape} dir {direction} halo={mode_used} peer_exchanges={n_peer}: ranks={world} n_active={n_active}/{n1} iters={info.iterations}/"
                  f"{info1.iterations} tau={tau:.10f}/{tau1:.10f} {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
        else:
            full_mask_ok = True   # each rank checks its own slab against the oracle-free single-GPU answer on rank 0 only
            ok = ok and chk and full_mask_ok
    # Repeated solves must be bit-identical: every reduction is ordered (block partials summed in block order,
    # one all-reduce per scalar), so any ghost plane read before its neighbour stored it -- or overwritten while
    # still being read -- shows up as a changed bit.  mg_degree = 1 exchanges the same level-0 field in
    # consecutive sweeps (the schedule the peer halo's write-after-read argument is weakest for).
    import hashlib
    for shape, direction, radius, deg, reps in [((128, 128, 128), 2, 8, 0, 10), ((128, 96, 64), 2, 6, 1, 6),
                                                ((256, 192, 128), 2, 8, 2, 4)]:
        full = synth.sphere_packing_slab(shape, seed=17, radius=radius, solid_target=0.5)
        z0, nzl = capi.slab_partition(shape[0], world)[rank]
        s = capi.Solver(shape, direction, 1, -1.0, 1.0, device=local, z_begin=z0, nz_local=nzl, comm=comm,
                        mg_degree=deg)
        s.set_phase(np.ascontiguousarray(full[z0:z0 + nzl]))
        sigs = []
        for rep in range(reps):
            s.build_mask()
            info = s.solve()
            fin, fout, _, _ = s.fluxes()
            sigs.append((info.iterations, info.rel_residual, fin, fout, hashlib.sha256(s.solution().tobytes()).hexdigest()))
        s.close()
        same = all(sg == sigs[0] for sg in sigs) and bool(info.converged)
        if rank == 0:
            print(f"repeat {shape} mg_degree={deg or 4}: {reps} solves, iters={sigs[0][0]} relres={sigs[0][1]:.3e} "
                  f"{'BIT-IDENTICAL' if same else 'DIFFER: ' + str([sg[:4] for sg in sigs])}", flush=True)
        ok = ok and same
    # homogenisation cell problem on z-slabs of a periodic box: same tensor as one GPU
    from openimpala_b200.effdiff import calculate_Deff_tensor_homogenization
    for shape, radius, halo_mode in [((96, 64, 80), 6, capi.OI_HALO_AUTO), ((64, 48, 36), 5, capi.OI_HALO_NCCL),
                                     ((128, 128, 128), 8, capi.OI_HALO_AUTO)]:
        full = synth.sphere_packing_slab(shape, seed=13, radius=radius, solid_target=0.5)
        z0, nzl = capi.slab_partition(shape[0], world)[rank]
        slab = np.ascontiguousarray(full[z0:z0 + nzl])
        D, conv, infos = calculate_Deff_tensor_homogenization(slab, 1, global_shape=shape, z_begin=z0, nz_local=nzl,
                                                              comm=comm, device=local, halo_mode=halo_mode)
        if rank == 0:
            D1, conv1, infos1 = calculate_Deff_tensor_homogenization(full, 1, device=local)
            good = conv and conv1 and float(np.abs(D - D1).max()) <= 1e-8
            print(f"cell problem {shape} halo={halo_mode}: iters={[i['iterations'] for i in infos]}/"
                  f"{[i['iterations'] for i in infos1]} Dxx={D[0][0]:.10f}/{D1[0][0]:.10f} "
                  f"max|dD|={float(np.abs(D - D1).max()):.2e} {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
        else:
            ok = ok and conv
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY " + ("PASS" if int(flag) == 1 else "FAIL"), flush=True)
    return 0 if int(flag) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
