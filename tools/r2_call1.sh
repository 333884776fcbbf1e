#!/bin/bash
# Round-2 GPU call 1 (one B200): full GPU suite, the round's first bench line, multigrid knob sweeps at
# 1024^3, the phase profile, and the ncu launch list + DRAM bytes of the default 1024^3 command.
O=gpurun_out/r2c1; mkdir -p $O
python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
python bench.py --steps 3 --warmup 3 > $O/bench1024.json 2> $O/bench1024.err; echo "bench rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
OI_PROFILE=1 $B > $O/prof1024.json 2> $O/prof1024.err
for dc in 6 8; do OI_MG_DEG_COARSE=$dc $B > $O/sweep_degc$dc.json 2> $O/sweep_degc$dc.err; done
for wf in 2 3; do OI_MG_W_FROM=$wf $B > $O/sweep_w$wf.json 2> $O/sweep_w$wf.err; done
OI_MG_DEG_COARSE=6 OI_MG_W_FROM=2 $B > $O/sweep_degc6_w2.json 2> $O/sweep_degc6_w2.err
for d in 3 5; do $B --mg-degree $d > $O/sweep_deg$d.json 2> $O/sweep_deg$d.err; done
OI_BENCH_SIZE=512 $B > $O/bench512.json 2> $O/bench512.err
OI_BENCH_SIZE=512 OI_MG_W_FROM=2 $B > $O/bench512_w2.json 2> $O/bench512_w2.err
# ncu: launch list of the default command (one timed step), then DRAM bytes of the hot kernels
N="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_1024.csv $N > $O/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"l0_ring|l0_pair|axpy2|xpby|coarse_stencil|prolong" -s 200 -c 120 --csv --log-file $O/dram_1024.csv $N > $O/ncu_dram.log 2>&1
ls -la $O
tail -3 $O/tests.log
