#!/bin/bash
# Round-2 GPU call 7 (one B200): evidence for the final defaults -- launch list and phase profile (with the mask
# build split up), Chebyshev interval sweep, one ncu --set full capture of the two top kernels, full bench line.
O=gpurun_out/r2c7; mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
OI_PROFILE=1 $B > $O/prof.json 2> $O/prof.err
for lo in 0.08 0.17 0.22; do OI_MG_LO0=$lo $B > $O/lo0_$lo.json 2> $O/lo0_$lo.err; done
for lo in 0.05 0.12 0.16; do OI_MG_LOC=$lo $B > $O/loc_$lo.json 2> $O/loc_$lo.err; done
N="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_1024.csv $N > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"l0_pair512u_kernel|l0_ring_kernel" -s 30 -c 4 -o $O/top_kernels_full $N > $O/ncu_full.log 2>&1
python bench.py --steps 3 --warmup 3 > $O/bench1024.json 2> $O/bench1024.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/ref_arm.json 2> $O/ref_arm.err
ls -la $O | head -40
