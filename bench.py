#!/usr/bin/env python
"""bench.py -- tau solve on the BASELINE.json workload (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full pass of the hot path for one direction:
    phase field (uint8) -> volume-fraction count -> percolation mask (CCL) ->
    connectivity bytes -> MG-PCG solve to 1e-9 -> boundary fluxes -> tau.
`value`  : DOF*iter/s, phase field already resident in HBM when the timer starts.
`e2e`    : same metric through the public class (TortuosityHypre + value()) with
           the phase field in pinned HOST memory; H2D/D2H inside the timed region.
Workload : synthetic overlapping-sphere packing S^3 (default 1024^3), tau in Z;
           N GPUs split the SAME box into z-slabs (strong scaling).
--impl reference : the CPU restatement (oracle/oi_oracle.c, OpenMP on all host
           cores) on a bounded sample; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tau_solve_dof_iter_per_s"
UNIT = "DOF*iter/s"
SEED, RADIUS, SOLID = 12345, 12, 0.60
CPU_SAMPLE_N = 128          # bounded CPU sample: 128^3 packing with the same generator


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=int(os.environ.get("OI_BENCH_SIZE", "1024")))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--direction", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mg-degree", type=int, default=0)
    return ap.parse_args()


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps: int, warmup: int, direction: int):
    """The reference path restated on the host cores (oracle/oi_oracle.c)."""
    import numpy as np
    from openimpala_b200 import synth
    from oracle import oi_c
    ph = synth.sphere_packing(CPU_SAMPLE_N, SEED, RADIUS, SOLID).astype(np.int32)
    n = ph.size
    out = None
    for _ in range(max(0, min(warmup, 1))):
        out = oi_c.tortuosity(ph, 1, direction, -1.0, 1.0, eps=1e-9)
    t0 = time.perf_counter()
    its = 0
    for _ in range(steps):
        out = oi_c.tortuosity(ph, 1, direction, -1.0, 1.0, eps=1e-9)
        its += out["iters"]
    dt = time.perf_counter() - t0
    return dict(value=n * its / dt, seconds=dt, iters=out["iters"], tau=out["tau"], n=n,
                cores=oi_c.num_threads(),
                sample=f"{CPU_SAMPLE_N}^3 sphere packing (seed {SEED}, R {RADIUS}), tau in "
                       f"{'XYZ'[direction]}, full path mask+assemble+Jacobi-PCG to 1e-9, {steps} step(s)")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    direction = args.direction

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 3))
        r = cpu_reference_run(steps, args.warmup, direction)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"sphere-packing {args.size}^3 tau in Z (bounded CPU sample: {r['sample']})"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "iterations": r["iters"], "tau": r["tau"],
            "note": "reference (AMReX+HYPRE+MPI+gfortran) cannot be built in this image; this is the "
                    "repo's C/OpenMP restatement with a Jacobi-PCG solver, not HYPRE FlexGMRES+SMG",
        }
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    from openimpala_b200 import capi, synth
    from openimpala_b200.tortuosity import Direction, SolverType, TortuosityHypre

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = capi.Comm(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local_rank)

    n = args.size
    shape = (n, n, n)
    z_begin, nz_local = capi.slab_partition(n, world)[rank]
    t_gen = time.perf_counter()
    slab = synth.sphere_packing_slab(shape, SEED, RADIUS, SOLID, z_begin, nz_local)
    t_gen = time.perf_counter() - t_gen
    host_pinned = torch.from_numpy(slab).pin_memory()
    d_phase = host_pinned.to(dev, non_blocking=False)          # resident copy for `value`
    ncells = n * n * n

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM ----------------
    S = capi.Solver(shape, direction, 1, -1.0, 1.0, eps=1e-9, maxiter=200, device=local_rank,
                    z_begin=z_begin, nz_local=nz_local, comm=comm, mg_degree=args.mg_degree)

    def step_resident():
        S.set_phase_device(d_phase.data_ptr())
        pc, tc = S.volume_fraction()
        n_active = S.build_mask()
        info = S.solve()
        fin, fout, ni, no = S.fluxes()
        return pc, n_active, info, fin, fout

    for _ in range(args.warmup):
        res = step_resident()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = S.launch_count()
    S.timer_record(0)
    t0 = time.perf_counter()
    iters_total = 0
    solve_ms = setup_ms = 0.0
    for _ in range(args.steps):
        res = step_resident()
        iters_total += res[2].iterations
        solve_ms += res[2].solve_ms
        setup_ms += res[2].setup_ms
    S.timer_record(1)
    dev_ms = S.timer_elapsed_ms(0, 1)
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = S.launch_count() - l0
    halo_mode, halo_peer_exchanges = S.halo_info()
    tmax = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = (float(v) for v in tmax.cpu())
    pc, n_active, info, fin, fout = res
    active_vf = n_active / ncells
    from openimpala_b200.tortuosity import tau_from_fluxes
    tau, deff, _ = tau_from_fluxes(fin, fout, active_vf, float(n), float(n) * n, -1.0, 1.0)
    value = ncells * iters_total / (dev_ms * 1e-3)

    # ---------------- roofline: dominant kernel timed live ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
    kern = {}
    local_cells = n * n * nz_local
    # algorithmic bytes per cell; multigrid vectors are fp32
    # Dense-box counts first; then the bytes the kernels really touch: a 16-byte group (2 fp64 /
    # 4 fp32 cells) without an unknown is skipped by the vector kernels and not stored by the
    # stencil kernels (its loads are still staged), flags are always read.
    n_unk, pairs, quads = S.sparsity()
    fp, fq = 2.0 * pairs / local_cells, 4.0 * quads / local_cells     # fraction of cells in touched groups
    bytes_dense = {"apply": 17.0, "smooth": 13.0, "residual_restrict": 9.5, "axpy2_dot": 33.0,
                   "xpby": 37.0, "dot": 16.0}
    bytes_touched = {"apply": 9.0 + 8.0 * fp,              # p 8 + flags 1 staged for every cell; q stored per pair
                     "smooth": 9.0 + 4.0 * fq,             # z 4 + r 4 + flags 1 staged; z' stored per quad
                     "residual_restrict": 9.5,
                     "axpy2_dot": 1.0 + 32.0 * fp,         # flags; r (r/w), q, r32, z1 per pair (x += alpha p is deferred)
                     "xpby": 1.0 + 36.0 * fp,              # flags; x (r/w), p (r/w), z per pair
                     "dot": 16.0}
    bytes_per_cell = bytes_touched
    for name, bpc in bytes_per_cell.items():
        ms, _ = S.time_kernel(name, 10)
        kern[name] = {"ms": ms, "gbs_dense": bytes_dense[name] * local_cells / (ms * 1e-3) / 1e9,
                      "bytes_per_cell_dense": bytes_dense[name],
                      "gbs_touched_16B_groups": bpc * local_cells / (ms * 1e-3) / 1e9, "bytes_per_cell_touched": bpc}
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        # ncu dram bytes per launch of the APPLY kernel, captured at tj["apply_cells"] cells and
        # scaled to this launch's cell count (traffic is proportional to cells for this kernel)
        traffic = tj["apply_bytes_per_launch"] / tj["apply_cells"] * local_cells
    except Exception:
        pass
    # headline: SURVEY 8(d)'s algorithmic figure for K3 (17 B per cell of the dense box) x cells per launch
    roofline = {"bound": "hbm", "kernel": "l0_ring_kernel<double,APPLY,dot> (y = A p, p.Ap)",
                "achieved": kern["apply"]["gbs_dense"], "peak": peak, "unit": "GB/s",
                "frac": kern["apply"]["gbs_dense"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": 17.0,
                "touched_fraction": {"fp64_pairs": fp, "fp32_quads": fq, "unknown_cells": n_unk / local_cells},
                "cells_per_launch": local_cells,
                "kernels": kern}
    S.close()

    # ---------------- e2e: public class, host buffers ----------------
    e2e = None
    if not args.no_e2e:
        host_np = host_pinned.numpy()

        e2e_parts = {"construct_ms": 0.0, "value_ms": 0.0, "close_ms": 0.0, "solve_ms": 0.0}

        def step_e2e():
            ta = time.perf_counter()
            t = TortuosityHypre(None, None, None, host_np, 0.5, 1, Direction(direction), SolverType.FlexGMRES,
                                "", -1.0, 1.0, global_shape=shape, z_begin=z_begin, nz_local=nz_local,
                                comm=comm, device=local_rank, mg_degree=args.mg_degree)
            tb = time.perf_counter()
            v = t.value()
            tc = time.perf_counter()
            it = t.getSolverIterations()
            e2e_parts["solve_ms"] += t.last_info.solve_ms if t.last_info is not None else 0.0
            t.close()
            td = time.perf_counter()
            e2e_parts["construct_ms"] += 1e3 * (tb - ta)
            e2e_parts["value_ms"] += 1e3 * (tc - tb)
            e2e_parts["close_ms"] += 1e3 * (td - tc)
            return v, it
        for _ in range(min(args.warmup, 1)):
            step_e2e()
        sync_all()
        for k_ in e2e_parts:
            e2e_parts[k_] = 0.0
        t0 = time.perf_counter()
        its = 0
        for _ in range(args.steps):
            tau_e2e, it = step_e2e()
            its += it
        sync_all()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.cpu())
        per_step_iters = its / args.steps
        e2e = {"value": ncells * its / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(slab.nbytes),
               "d2h_bytes_per_step": int(8 * (per_step_iters + 4) + 16 + 24 + 8),
               "ms_per_step": 1e3 * dt / args.steps, "tau": tau_e2e,
               "rank0_breakdown_ms_per_step": {k_: v_ / args.steps for k_, v_ in e2e_parts.items()},
               "timed": "wall clock around TortuosityHypre(...).value(), max over ranks; includes handle "
                        "creation, cudaMalloc, H2D of the uint8 phase slab from pinned memory"}

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 0, direction)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "seconds": r["seconds"], "iterations": r["iters"],
               "note": "C/OpenMP restatement with Jacobi-PCG on the stored 7-coefficient matrix; the "
                       "reference's HYPRE FlexGMRES+SMG stack is not buildable in this image"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "precision_note": "operator apply, Krylov vectors, dots and fluxes in fp64; multigrid preconditioner vectors in fp32",
            "config": {"workload": f"sphere-packing {n}^3 uint8 (seed {SEED}, R {RADIUS}, solid {SOLID}), "
                                   f"tau in {'XYZ'[direction]}, phase 1, eps 1e-9, MG-PCG",
                       "parallelism": f"z-slabs x{world}",
                       "halo": {0: "none (single slab)", 1: "NCCL send/recv",
                                2: "peer-memory stores over NVLink (CUDA IPC) + stream wait on flag words"}[halo_mode],
                       "l2_policy": "inputs larger than L2 (every fp64 vector >= 1 GiB at 512^3+)",
                       "sparsity_note": "16-byte groups without an unknown (solid) are skipped by the vector kernels and "
                                        "not stored by the stencil kernels: gbs_dense uses SURVEY 8(d)'s dense-box bytes "
                                        "(can exceed the copy peak where sectors are skipped), gbs_touched counts only "
                                        "occupied 16-byte groups (a lower bound on DRAM traffic, which moves whole sectors)",
                       "porosity": float(slab.mean()) if world == 1 else None,
                       "generate_s": t_gen},
            "time_to_solution_s": dev_ms * 1e-3 / args.steps,
            "wall_ms_per_step": wall_ms / args.steps,
            "solve_ms_per_step": solve_ms / args.steps, "mg_setup_ms_per_step": setup_ms / args.steps,
            "iterations": info.iterations, "rel_residual": info.rel_residual,
            "converged": bool(info.converged), "tau": tau, "deff": deff, "active_cells": n_active,
            "phase_cells": pc, "flux_in": fin, "flux_out": fout,
            "dof_active_iter_per_s": n_active * iters_total / (dev_ms * 1e-3),
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e,
        }
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
