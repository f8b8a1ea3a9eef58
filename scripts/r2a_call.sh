#!/bin/bash
# Round 2, GPU call A (1 GPU): the suite after the stopping-rule fix, the tile variants, batch lanes, the configs, a bench line.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -rxXs --durations=15 > gpurun_out/r2a_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.txt
for tp in 1024 960 896 0; do
  echo "== TILE_POOLS=$tp" >> gpurun_out/r2a_tiles.txt
  TILE_POOLS=$tp timeout 300 python scripts/microbench.py 2>&1 | tail -5 >> gpurun_out/r2a_tiles.txt
done
echo "== TILE_POOLS=0 BLOCKED_CFG=3 (regs variant)" >> gpurun_out/r2a_tiles.txt
TILE_POOLS=0 BLOCKED_CFG=3 timeout 300 python scripts/microbench.py 2>&1 | tail -5 >> gpurun_out/r2a_tiles.txt
echo "== TILE_POOLS=1024 BLOCKED_CFG=3 (regs variant)" >> gpurun_out/r2a_tiles.txt
TILE_POOLS=1024 BLOCKED_CFG=3 timeout 300 python scripts/microbench.py 2>&1 | tail -5 >> gpurun_out/r2a_tiles.txt
for mp in 125000 250000 500000; do
  echo "== M_POOLS=$mp (strong-scaling shard sizes)" >> gpurun_out/r2a_tiles.txt
  M_POOLS=$mp timeout 300 python scripts/microbench.py 2>&1 | tail -3 >> gpurun_out/r2a_tiles.txt
done
for lanes in 1 32; do
  echo "== CFMM_BATCH_LANES=$lanes" >> gpurun_out/r2a_batch_lanes.txt
  CFMM_BATCH_LANES=$lanes timeout 300 python scripts/time_batch.py >> gpurun_out/r2a_batch_lanes.txt 2>&1
done
timeout 600 python scripts/run_configs.py > gpurun_out/r2a_configs.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_driver_flags.json 2> gpurun_out/r2a_bench_driver_flags.err
timeout 600 python bench.py --no-cpu > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
tail -25 gpurun_out/r2a_pytest_gpu.txt; cat gpurun_out/r2a_tiles.txt; cat gpurun_out/r2a_configs.txt
grep -E '"ms"|LANES' gpurun_out/r2a_batch_lanes.txt | head -20
cut -c1-700 gpurun_out/r2a_bench_driver_flags.json; cut -c1-400 gpurun_out/r2a_bench_default.json
