#!/bin/bash
# Round-2 GPU call 20 (one B200, short): the host-app GPU tests after the change to EffectiveDiffusivityHypre / the stand-in.
O=gpurun_out/r2c20; mkdir -p $O
timeout 120 python -m pytest tests/test_host_apps.py -m gpu -q -rs -x > $O/tests.log 2>&1
echo "rc=$?"; tail -6 $O/tests.log | cut -c1-300
