import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys, time, os
sys.path.insert(0, _ROOT)
import numpy as np
from openimpala_b200 import capi, synth
n = int(sys.argv[1])
ph = synth.sphere_packing(n, 12345, 12, 0.60)
s = capi.Solver(ph.shape, 2, 1, -1.0, 1.0)
s.set_phase(ph)
for rep in range(3):
    s.timer_record(0); na = s.build_mask(); s.timer_record(1)
    print(f"n={n} OI_CCL_ONE_PASS={os.environ.get('OI_CCL_ONE_PASS','0')} build_mask {s.timer_elapsed_ms(0,1):.2f} ms n_active {na}", flush=True)
s.close()
