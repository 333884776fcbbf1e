#!/bin/sh
# One gpurun call that measures everything round 1 left unmeasured (end of section 10).
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'sh tools/round2_first_call.sh'
# Everything lands in gpurun_out/r2_first/.
OUT=gpurun_out/r2_first
mkdir -p $OUT
timeout 240 python -m pytest tests -m gpu -q --maxfail=10 --durations=15 > $OUT/gpu_tests.log 2>&1; echo "rc=$?" >> $OUT/gpu_tests.log
timeout 180 python bench.py --steps 3 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
OI_TIMING_SIZES=64,128,256 timeout 60 python tools/graph_timing.py > $OUT/graph_tail_timing.json 2> $OUT/graph_tail_timing.err
OI_TAIL_SMEM=1 OI_TIMING_SIZES=64,128,256 timeout 60 python tools/graph_timing.py > $OUT/graph_tail_smem_timing.json 2> $OUT/graph_tail_smem_timing.err
for w in 1 3; do
    timeout 60 openimpala_b200/bin/Diffusion tests/inputs/diffusion_flow_through.inputs results_path=$OUT/dir_w$w/ b200.dir_workers=$w verbose=1 2>&1 | grep -E "workers|Total run time|Calculated Tortuosity" > $OUT/dir_workers_$w.log
done
for w in 1 4; do
    timeout 120 openimpala_b200/bin/Diffusion filename=SampleData_2Phase_stack_3d_1bit.tif data_path=tests/golden/ results_path=$OUT/rev_w$w/ \
        rev.do_study=1 rev.num_samples=4 "rev.sizes=32 48 64" calculation_method=skip_if_rev rev.verbose=0 verbose=1 b200.rev_workers=$w 2>&1 \
        | grep -E "workers|Total run time" > $OUT/rev_workers_$w.log
done
cmp $OUT/rev_w1/rev_study_Deff.csv $OUT/rev_w4/rev_study_Deff.csv > $OUT/rev_workers_cmp.log 2>&1; echo "cmp rc=$?" >> $OUT/rev_workers_cmp.log
tail -3 $OUT/gpu_tests.log; cat $OUT/dir_workers_1.log $OUT/dir_workers_3.log $OUT/rev_workers_1.log $OUT/rev_workers_4.log $OUT/rev_workers_cmp.log
