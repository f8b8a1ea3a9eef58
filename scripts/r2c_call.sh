#!/bin/bash
# Round 2, call C (2 GPUs): new bench.py at N=1 (driver flags) and N=2 (torchrun), reference arm.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
echo "n2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err
echo "ref rc=$?"
tail -5 gpurun_out/r2c_bench_n1.err gpurun_out/r2c_bench_n2.err gpurun_out/r2c_bench_ref.err
python - <<'PY'
import json
for f in ("n1","n2","ref"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2c_bench_{f}.json") if l.startswith("{")][-1])
        keep={k:d.get(k) for k in ("value","ms_per_step","steps","scaling","eval_only_us","allreduce_us","weak","gpu_launches","clocks")}
        keep["frac"]=(d.get("roofline") or {}).get("frac"); keep["e2e"]={k:(d.get("e2e") or {}).get(k) for k in ("value","wall_s","wall_s_all","evals","hvps","status","native_loop")}
        keep["cpu"]=d.get("cpu_baseline")
        print(f, json.dumps(keep))
        for c in d.get("configs") or []:
            print("   ", {k:c.get(k) for k in ("config","time_to_1e-6_gap_ms","solver_ms","status","evals","value","obj_rel_diff_vs_oracle")}, c.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
