#!/bin/bash
# Round-2 GPU call 5 (two B200s): re-validate the multi-rank path after the publish / coarse-halo / CCL changes,
# A/B fused vs explicit halo, CCL old vs new at N=1.
O=gpurun_out/r2c5; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_host_apps.py tests/test_gpu_parity.py -q -m gpu -k "multi_gpu_parity or several_ranks or mask_rows or flux_gate or sample_image" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -5 $O/tests.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-weak-extra"
$T --master-port 29521 $B > $O/n2_default.json 2> $O/n2_default.err; echo "default rc=$?"
OI_HALO_FUSE=0 $T --master-port 29522 $B > $O/n2_nofuse.json 2> $O/n2_nofuse.err
OI_PROFILE=1 $T --master-port 29523 $B > $O/n2_prof.json 2> $O/n2_prof.err
B1="timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B1 > $O/n1_default.json 2> $O/n1_default.err
OI_CCL=0 $B1 > $O/n1_ccl_old.json 2> $O/n1_ccl_old.err
ls $O | wc -l
