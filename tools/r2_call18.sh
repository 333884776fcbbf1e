#!/bin/bash
# Round-2 GPU call 18 (two B200s, short): z-chunk cap of the ring kernels on 512-plane slabs (64 against 128), and the
# phase profile of the final defaults at N = 2.
O=gpurun_out/r2c18; mkdir -p $O
run() { tag=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 4 --warmup 2 --no-e2e --no-weak-extra > $O/$tag.json 2> $O/$tag.err; }
run default
OI_ZCHUNK=128 run zc128
OI_PROFILE=1 run prof
python - <<'PY'
import json
for t in ("default", "zc128", "prof"):
    try:
        d = json.load(open(f"gpurun_out/r2c18/{t}.json"))
        print(t, d["ms_per_step"], d["iterations"], d["parity"].get("counts_exact"), d["parity"].get("tau_rel_vs_golden"))
    except Exception as e:
        print(t, "failed", e)
PY
