#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/time_solve.py > gpurun_out/r2e_time_solve.txt 2>&1; cat gpurun_out/r2e_time_solve.txt
python - <<'PY' 2>&1 | tail -8
import time, sys
sys.path.insert(0, ".")
import torch, cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
d = I.two_asset_instance()
us = [cf.Swap(0, 2, t) for t in d["amounts"]]
args = (d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], us)
for rep in range(3):
    t0 = time.perf_counter(); rs = cf.solve_sweep(*args, tol=1e-9, batched=False); dt = time.perf_counter() - t0
    print("sequential sweep", rep, f"{dt:.3f}s", sum(r.iters for r in rs), sum(r.evals for r in rs))
PY
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15
