#!/bin/bash
# Round-2 GPU call 19 (two B200s, short): the homogenisation method of the Diffusion app on two ranks (C++ class on
# z-slabs) and the flow-through rank test again (the stand-in's FillBoundary changed).
O=gpurun_out/r2c19; mkdir -p $O
timeout 170 python -m pytest tests/test_host_apps.py -m gpu -q -rs -x \
  -k "homogenization_on_two_ranks or several_ranks or default_method_is_homogenization" > $O/tests.log 2>&1
echo "rc=$?"; tail -25 $O/tests.log | cut -c1-400
