#!/bin/bash
# Round-2 GPU call 21 (one B200, what is left of the budget): the whole GPU suite at the final commit.
O=gpurun_out/r2c21; mkdir -p $O
timeout 75 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1
echo "rc=$?"; tail -3 $O/tests.log | cut -c1-300
