#!/bin/bash
# Round-2 GPU call 15 (eight B200s, short): strong 1024^3 at N = 8, 4, 2 and weak 2048^3 at N = 8 on the final defaults.
O=gpurun_out/r2c15; mkdir -p $O
run() { n=$1; port=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 8 29561 --steps 5 --warmup 2 --no-e2e > $O/n8_default.json 2> $O/n8_default.err; echo "n8 rc=$?"
run 4 29562 --steps 5 --warmup 2 --no-e2e --no-weak-extra > $O/n4_default.json 2> $O/n4_default.err
run 2 29563 --steps 5 --warmup 2 --no-e2e --no-weak-extra > $O/n2_default.json 2> $O/n2_default.err
ls $O | wc -l
