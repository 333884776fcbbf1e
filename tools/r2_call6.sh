#!/bin/bash
# Round-2 GPU call 6 (eight B200s): multi-rank parity at 8 ranks, Diffusion on 4 ranks, strong 1024^3 at N=8 / 4,
# weak 2048^3 at N=8, fused vs explicit halo, phase profile.
O=gpurun_out/r2c6; mkdir -p $O
nvidia-smi -L | wc -l > $O/gpus.txt
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_host_apps.py -q -m gpu -k "(multi_gpu_parity and 8) or (several_ranks and 4)" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -6 $O/tests.log | cut -c1-300
T8="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
T4="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$T8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > $O/n8_default.json 2> $O/n8_default.err; echo "n8 rc=$?"
OI_HALO_FUSE=0 $T8 --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n8_nofuse.json 2> $O/n8_nofuse.err
OI_PROFILE=1 $T8 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-weak-extra > $O/n8_prof.json 2> $O/n8_prof.err
OI_TAIL=1 OI_AGG_CELLS=2097152 $T8 --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n8_tail_agg.json 2> $O/n8_tail_agg.err
$T4 --master-port 29535 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n4_default.json 2> $O/n4_default.err
$T8 --master-port 29536 bench.py --gpus 8 --steps 3 --warmup 2 --scaling weak > $O/n8_weak_e2e.json 2> $O/n8_weak_e2e.err
ls $O | wc -l
