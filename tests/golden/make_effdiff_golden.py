"""Golden D_eff tensors of the homogenisation path on the reference's sample image
(oracle/oi_effdiff.py: scipy CG to 1e-12 on the rows of effdiff_fillmtx).

    python tests/golden/make_effdiff_golden.py     # writes tests/golden/effdiff_golden.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oi_effdiff as oe  # noqa: E402
from oracle import oi_numpy as o  # noqa: E402

ph = o.threshold(o.read_tiff_raw(os.path.join(HERE, "SampleData_2Phase_stack_3d_1bit.tif")), 0.5)
out = {"image": "SampleData_2Phase_stack_3d_1bit.tif", "threshold": 0.5, "eps": 1e-12,
       "generator": "oracle/oi_effdiff.py deff_tensor (restatement, not HYPRE output)"}
for phase_id in (1, 0):
    D = oe.deff_tensor(ph, phase_id, eps=1e-12)
    out[f"phase{phase_id}"] = {"deff": D.tolist(), "n_active": int((ph == phase_id).sum())}
    print(phase_id, D)
json.dump(out, open(os.path.join(HERE, "effdiff_golden.json"), "w"), indent=1)
