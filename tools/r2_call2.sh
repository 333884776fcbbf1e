#!/bin/bash
# Round-2 GPU call 2 (one B200): new kernel variants (TMA ring, unrolled pair, NC=2 vector kernels) against the
# defaults, and the (level-0 degree, coarse degree) sweep at 1024^3.
O=gpurun_out/r2c2; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -q -k "tma_ring or pair_kernel_matches or flux_gate or tau_matches or sphere_packing_golden or coarse_tail or symmetric" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -3 $O/tests.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > $O/base.json 2> $O/base.err
OI_COARSE_HALF=0 $B > $O/nohalf.json 2> $O/nohalf.err
OI_TMA=1 $B > $O/tma.json 2> $O/tma.err
OI_TMA=1 OI_PAIR=0 $B > $O/tma_nopair.json 2> $O/tma_nopair.err
OI_PAIR=0 $B > $O/nopair.json 2> $O/nopair.err
OI_PAIR=3 $B > $O/pair3.json 2> $O/pair3.err
OI_VEC_NC=2 $B > $O/vecnc2.json 2> $O/vecnc2.err
for dc in 8 10 12; do OI_MG_DEG_COARSE=$dc $B > $O/d4_c$dc.json 2> $O/d4_c$dc.err; done
for dc in 8 12; do OI_MG_DEG_COARSE=$dc $B --mg-degree 3 > $O/d3_c$dc.json 2> $O/d3_c$dc.err; done
OI_MG_DEG_COARSE=10 $B --mg-degree 5 > $O/d5_c10.json 2> $O/d5_c10.err
OI_BENCH_SIZE=512 OI_MG_DEG_COARSE=8 $B > $O/s512_d4_c8.json 2> $O/s512_d4_c8.err
ls $O
