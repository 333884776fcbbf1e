#!/bin/bash
# Round-2 GPU call 12 (one B200): level-1 two-sweep pass -- parity against single sweeps and A/B; fourth-kind
# Chebyshev weights and a separate level-1 degree as experiments; full suite.
O=gpurun_out/r2c12; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log | cut -c1-300
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > $O/default.json 2> $O/default.err
OI_COARSE_PAIR=0 $B > $O/nocoarsepair.json 2> $O/nocoarsepair.err
OI_PROFILE=1 $B > $O/prof.json 2> $O/prof.err
OI_MG_CHEB4=1 $B > $O/cheb4.json 2> $O/cheb4.err
OI_MG_DEG_L1=6 $B > $O/l1deg6.json 2> $O/l1deg6.err
OI_MG_DEG_L1=10 $B > $O/l1deg10.json 2> $O/l1deg10.err
OI_BENCH_SIZE=512 $B > $O/s512.json 2> $O/s512.err
ls $O | wc -l
