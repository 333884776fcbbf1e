"""Multi-rank coverage.  CPU: the host-side slab logic under a real
world_size-2 gloo group (partition, slab generation, reduction of per-slab
integer results).  GPU: the NCCL path end to end when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from openimpala_b200 import capi, synth
    from oracle import oi_numpy as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = (48, 40, 44)
    z0, nzl = capi.slab_partition(shape[0], world)[rank]
    slab = synth.sphere_packing_slab(shape, seed=5, radius=5, solid_target=0.5, z_begin=z0, nz_local=nzl)
    # what the C-ABI all-reduces: per-slab phase counts
    cnt = torch.tensor([int((slab == 1).sum()), slab.size], dtype=torch.int64)
    dist.all_reduce(cnt)
    # slabs must tile the box and agree with the single-process image
    parts = [None] * world
    dist.all_gather_object(parts, (z0, nzl, slab))
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        full = np.concatenate([p[2] for p in parts])
        ref = synth.sphere_packing_slab(shape, seed=5, radius=5, solid_target=0.5)
        q.put((np.array_equal(full, ref), int(cnt[0]) == o.volume_fraction_counts(ref, 1)[0],
               int(cnt[1]) == ref.size, [p[:2] for p in parts]))
    dist.destroy_process_group()


def test_slab_logic_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, count_ok, total_ok, parts = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert same and count_ok and total_ok
    assert parts[0][0] == 0 and parts[0][0] + parts[0][1] == parts[1][0] and parts[1][0] % 2 == 0


@pytest.mark.gpu
def test_two_sweep_passes_on_slabs():
    """OI_PAIR_SLAB=1: the pair kernel on z-slabs, fed by the boundary pre-sweep through the neighbours' vb planes
    (opt-in: no faster than single sweeps on thin slabs) -- the same parity worker must pass on 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500, env={**os.environ, "OI_PAIR_SLAB": "1"})
    assert r.returncode == 0 and "MULTI_GPU_PARITY PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    """Distributed result == single-GPU result (integers exact, tau 1e-8), peer halo and NCCL halo, the cell
    problem on periodic slabs, and bit-identical repeated solves, on every world size the box offers."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29611 + world),
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "MULTI_GPU_PARITY PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
