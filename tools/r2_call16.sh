#!/bin/bash
# Round-2 GPU call 16 (two B200s, short): the multi-GPU tests of the suite on the final build.
O=gpurun_out/r2c16; mkdir -p $O
timeout 400 python -m pytest tests/test_multi_rank.py tests/test_host_apps.py -m gpu -q -rs \
  -k "multi_gpu_parity or two_sweep_passes_on_slabs or several_ranks" > $O/tests_two_gpus.log 2>&1
echo "rc=$?"; tail -4 $O/tests_two_gpus.log
