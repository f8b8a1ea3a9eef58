#!/bin/bash
# Round 2, call Y (1 GPU, short): GPU suite + solve timings + the bench line under the driver's flags
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2y_pytest_gpu.txt; cat gpurun_out/r2y_pytest_gpu.txt
timeout 100 python scripts/time_solve.py > gpurun_out/r2y_time_solve.txt 2>&1; cat gpurun_out/r2y_time_solve.txt
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2y_bench_n1.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","steps")}, (d.get("roofline") or {}).get("frac"), {k:(d.get("e2e") or {}).get(k) for k in ("wall_s","evals","hvps","status","native_loop")}, d.get("cpu_baseline"))
for c in d.get("configs") or []: print("  ", {k:c.get(k) for k in ("config","time_to_1e-6_gap_ms","status","evals")})
PY
