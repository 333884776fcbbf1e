#!/bin/bash
# Round-2 GPU call 4 (two B200s): multi-rank parity (worker: distributed == single GPU, NCCL and peer halo,
# bit-identical repeated solves), Diffusion b200.ranks=2, and the halo A/B at 1024^3 strong + 1280^3 weak.
O=gpurun_out/r2c4; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
python -m pytest tests/test_multi_rank.py tests/test_host_apps.py -q -m gpu -k "multi_gpu_parity or several_ranks" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -15 $O/tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e"
$T --master-port 29511 $B > $O/n2_default.json 2> $O/n2_default.err; echo "default rc=$?"
OI_HALO_INKERNEL=0 $T --master-port 29512 $B --no-weak-extra > $O/n2_streamwait.json 2> $O/n2_streamwait.err
OI_HALO_FUSE=0 $T --master-port 29513 $B --no-weak-extra > $O/n2_nofuse.json 2> $O/n2_nofuse.err
OI_TAIL=1 $T --master-port 29514 $B --no-weak-extra > $O/n2_tail.json 2> $O/n2_tail.err
OI_AGG_CELLS=2097152 $T --master-port 29515 $B --no-weak-extra > $O/n2_agg128.json 2> $O/n2_agg128.err
OI_PROFILE=1 $T --master-port 29516 $B --no-weak-extra > $O/n2_prof.json 2> $O/n2_prof.err
OI_PROFILE=1 OI_HALO_FUSE=0 $T --master-port 29517 $B --no-weak-extra > $O/n2_prof_nofuse.json 2> $O/n2_prof_nofuse.err
python tools/debug_polish.py > $O/debug_polish.log 2>&1
ls $O | wc -l
