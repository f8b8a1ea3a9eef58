#!/bin/bash
# Round 2, last call (1 GPU, ~4 GPU-minutes left): the whole GPU suite including the loopback-rank tests of the
# pool-sharded protocols (4-at-a-time peer polling), the same loopback tests against a library built with the previous
# one-after-the-other polling (A/B: harness vs protocol), then the bench line under the driver's flags.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q --timeout 60 --durations=12 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2z4_pytest_gpu.txt; tail -25 gpurun_out/r2z4_pytest_gpu.txt
CFMM_LIB=build/variants/libcfmm_serialpoll.so timeout 60 python -m pytest tests/test_loopback_ranks.py -m gpu -q --timeout 40 -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2z4_loopback_serialpoll.txt; cat gpurun_out/r2z4_loopback_serialpoll.txt
timeout 90 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z4_bench_n1.json 2> gpurun_out/r2z4_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2z4_bench_n1.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","steps")}, (d.get("roofline") or {}).get("frac"), {k:(d.get("e2e") or {}).get(k) for k in ("wall_s","evals","hvps","status","native_loop")}, d.get("cpu_baseline"))
for c in d.get("configs") or []: print("  ", {k:c.get(k) for k in ("config","time_to_1e-6_gap_ms","status","evals")})
PY
