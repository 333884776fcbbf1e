#!/bin/bash
# Round-2 GPU call 8 (two B200s): the two-sweep pass on z-slabs (pair kernel fed by the boundary pre-sweep):
# parity worker (distributed == single GPU, bit-identical repeats) and A/B against single sweeps.
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_host_apps.py -q -m gpu -k "(multi_gpu_parity and 2) or (several_ranks and 2)" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -8 $O/tests.log | cut -c1-300
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-weak-extra"
$T --master-port 29541 $B > $O/n2_default.json 2> $O/n2_default.err; echo "default rc=$?"
OI_PAIR_SLAB=0 $T --master-port 29542 $B > $O/n2_nopairslab.json 2> $O/n2_nopairslab.err
OI_PROFILE=1 $T --master-port 29543 $B > $O/n2_prof.json 2> $O/n2_prof.err
ls $O | wc -l
