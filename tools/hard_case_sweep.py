import os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
sys.path.insert(0, _ROOT)
import numpy as np
from tools.robustness_sweep import blobs, run
for n in (128, 256, 384, 512):
    run(f"blobs {n}^3 porosity 0.25 sigma 1.0", blobs((n, n, n), 3, 0.25, 1.0))
    run(f"blobs {n}^3 porosity 0.22 sigma 1.0", blobs((n, n, n), 3, 0.22, 1.0))
