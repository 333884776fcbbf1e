#!/bin/bash
# Round-2 GPU call 14 (one B200): full GPU suite and bench lines on the final defaults (fourth-kind Chebyshev
# weights, degrees 5 / 4 / 8).
O=gpurun_out/r2c14; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
OI_PROFILE=1 $B > $O/prof.json 2> $O/prof.err
OI_MG_DEG_L1=3 $B > $O/l1deg3.json 2> $O/l1deg3.err
python bench.py --steps 5 --warmup 3 > $O/bench1024.json 2> $O/bench1024.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/ref_arm.json 2> $O/ref_arm.err
ls $O | wc -l
