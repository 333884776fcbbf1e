#!/bin/bash
# Round-2 GPU call 9 (one B200): full GPU suite on the current tree, labelling schedules A/B, Chebyshev interval
# combinations at the default degrees, profile of the mask build.
O=gpurun_out/r2c9; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log | cut -c1-300
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
OI_PROFILE=1 $B > $O/prof.json 2> $O/prof.err
OI_CCL=0 $B > $O/ccl_old.json 2> $O/ccl_old.err
for pr in "0.08 0.05" "0.05 0.05" "0.05 0.03" "0.03 0.03" "0.08 0.03" "0.10 0.06"; do set -- $pr; OI_MG_LO0=$1 OI_MG_LOC=$2 $B > $O/lo_$1_$2.json 2> $O/lo_$1_$2.err; done
OI_MG_LO0=0.05 OI_MG_LOC=0.05 OI_BENCH_SIZE=512 $B > $O/s512_lo_0.05_0.05.json 2> $O/s512_lo.err
OI_BENCH_SIZE=512 $B > $O/s512_default.json 2> $O/s512_default.err
ls $O | wc -l
