#!/bin/bash
# Round-2 GPU call 3 (one B200): full GPU suite on the new defaults (pair variant 3, coarse degree 8, half
# coefficient copies), level-0 degree sweep at coarse degree 8 / 6, and the full bench line.
O=gpurun_out/r2c3; mkdir -p $O
python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > $O/d4_c8.json 2> $O/d4_c8.err
OI_COARSE_HALF=0 $B > $O/d4_c8_nohalf.json 2> $O/d4_c8_nohalf.err
for d in 5 6 7; do $B --mg-degree $d > $O/d${d}_c8.json 2> $O/d${d}_c8.err; done
for d in 5 6; do OI_MG_DEG_COARSE=6 $B --mg-degree $d > $O/d${d}_c6.json 2> $O/d${d}_c6.err; done
OI_BENCH_SIZE=512 $B --mg-degree 5 > $O/s512_d5_c8.json 2> $O/s512_d5_c8.err
OI_BENCH_SIZE=512 $B --mg-degree 6 > $O/s512_d6_c8.json 2> $O/s512_d6_c8.err
python bench.py --steps 3 --warmup 3 > $O/bench1024.json 2> $O/bench1024.err; echo "bench rc=$?"
ls $O | wc -l
