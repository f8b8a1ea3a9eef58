#!/bin/bash
# Round 2, call X (8 GPUs): multi-GPU tests at 2/4/8 ranks, bench at N=8 and N=4 under the driver's flags
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x -s 2>&1 | tail -12 > gpurun_out/r2x_multigpu_test.txt; cat gpurun_out/r2x_multigpu_test.txt
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2x_bench_n$n.json 2> gpurun_out/r2x_bench_n$n.err
  echo "n$n rc=$?"; tail -n 3 gpurun_out/r2x_bench_n$n.err
done
python - <<'PY'
import json
for f in ("n8","n4"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2x_bench_{f}.json") if l.startswith("{")][-1])
        keep={k:d.get(k) for k in ("value","ms_per_step","steps","scaling","eval_only_us","allreduce_us","weak","gpu_launches")}
        keep["frac"]=(d.get("roofline") or {}).get("frac"); keep["e2e"]={k:(d.get("e2e") or {}).get(k) for k in ("value","wall_s","wall_s_all","evals","hvps","status","native_loop")}
        print(f, json.dumps(keep))
    except Exception as e:
        print(f, "ERR", e)
PY
