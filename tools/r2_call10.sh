#!/bin/bash
# Round-2 GPU call 10 (eight B200s): final multi-GPU numbers -- parity at 8 ranks, strong 1024^3 at N = 2, 4, 8,
# weak 2048^3 at N = 8, two-sweep passes on slabs A/B, phase profile with halo counters.
O=gpurun_out/r2c10; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_rank.py -q -m gpu -k "multi_gpu_parity and 8" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -6 $O/tests.log | cut -c1-300
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 8 29551 --steps 5 --warmup 3 --no-e2e > $O/n8_default.json 2> $O/n8_default.err; echo "n8 rc=$?"
OI_PAIR_SLAB=0 run 8 29552 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n8_nopairslab.json 2> $O/n8_nopairslab.err
OI_PROFILE=1 run 8 29553 --steps 3 --warmup 2 --no-e2e --no-weak-extra > $O/n8_prof.json 2> $O/n8_prof.err
run 4 29554 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n4_default.json 2> $O/n4_default.err
run 2 29555 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n2_default.json 2> $O/n2_default.err
OI_PAIR_SLAB=0 run 2 29556 --steps 5 --warmup 3 --no-e2e --no-weak-extra > $O/n2_nopairslab.json 2> $O/n2_nopairslab.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/n1_default.json 2> $O/n1_default.err
ls $O | wc -l
