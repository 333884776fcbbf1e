#!/bin/bash
# Round-2 GPU call 13 (one B200): fourth-kind Chebyshev weights with per-level degrees (level 0 / level 1 / below).
O=gpurun_out/r2c13; mkdir -p $O
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
export OI_MG_CHEB4=1
$B > $O/c4_d5_l8_c8.json 2> $O/c4_d5_l8_c8.err
OI_MG_DEG_L1=6 $B > $O/c4_d5_l6_c8.json 2> $O/c4_d5_l6_c8.err
OI_MG_DEG_L1=5 $B > $O/c4_d5_l5_c8.json 2> $O/c4_d5_l5_c8.err
OI_MG_DEG_L1=4 $B > $O/c4_d5_l4_c8.json 2> $O/c4_d5_l4_c8.err
OI_MG_DEG_L1=6 OI_MG_DEG_COARSE=6 $B > $O/c4_d5_l6_c6.json 2> $O/c4_d5_l6_c6.err
OI_MG_DEG_L1=6 $B --mg-degree 4 > $O/c4_d4_l6_c8.json 2> $O/c4_d4_l6_c8.err
OI_MG_DEG_L1=6 $B --mg-degree 6 > $O/c4_d6_l6_c8.json 2> $O/c4_d6_l6_c8.err
OI_MG_DEG_L1=6 OI_BENCH_SIZE=512 $B > $O/s512_c4_d5_l6_c8.json 2> $O/s512.err
OI_MG_DEG_L1=5 OI_BENCH_SIZE=512 $B > $O/s512_c4_d5_l5_c8.json 2> $O/s512b.err
ls $O | wc -l
