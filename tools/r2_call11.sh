#!/bin/bash
# Round-2 GPU call 11 (one B200): full GPU suite on the final tree, divide-and-conquer labelling A/B, degree /
# z-chunk re-check with the re-tuned Chebyshev intervals, final bench line and launch list.
O=gpurun_out/r2c11; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log | cut -c1-300
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
OI_PROFILE=1 $B > $O/prof.json 2> $O/prof.err
OI_CCL=0 $B > $O/ccl_old.json 2> $O/ccl_old.err
$B --mg-degree 4 > $O/d4.json 2> $O/d4.err
$B --mg-degree 6 > $O/d6.json 2> $O/d6.err
OI_MG_DEG_COARSE=6 $B > $O/d5_c6.json 2> $O/d5_c6.err
OI_ZCHUNK=32 $B > $O/zc32.json 2> $O/zc32.err
OI_ZCHUNK=128 $B > $O/zc128.json 2> $O/zc128.err
N="timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_1024.csv $N > $O/ncu_list.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/bench1024.json 2> $O/bench1024.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
ls $O | wc -l
