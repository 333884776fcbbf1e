#!/bin/bash
# Round-2 GPU call 17 (one B200): ncu launch list of two consecutive PCG iterations of the default 1024^3 command on the
# final build (a whole step is 8615 launches at ~0.2 s each under ncu: not affordable; call 11's list stopped at 2793).
O=gpurun_out/r2c17; mkdir -p $O
N="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 1000 --csv \
  --log-file $O/launches_1024_iterations.csv $N > $O/ncu_list.log 2>&1
echo "ncu rc=$?"; grep -c gpu__time_duration $O/launches_1024_iterations.csv
