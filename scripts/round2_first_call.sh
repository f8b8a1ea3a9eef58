#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in ONE gpurun call.
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Results land in gpurun_out/r2_*.txt.  (1) the GPU suite, with the xfail-tolerant first-run tests reported verbosely;
# (2) the eval / hvp kernels at every tile size; (3) the batch solver at 1 and 32 lanes per problem.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxX 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.txt
for tp in 1024 960 896 0; do
  echo "== TILE_POOLS=$tp" >> gpurun_out/r2_tiles.txt
  TILE_POOLS=$tp timeout 300 python scripts/microbench.py 2>&1 | tail -5 >> gpurun_out/r2_tiles.txt
done
for lanes in 1 32; do
  echo "== CFMM_BATCH_LANES=$lanes" >> gpurun_out/r2_batch_lanes.txt
  CFMM_BATCH_LANES=$lanes timeout 300 python scripts/time_batch.py >> gpurun_out/r2_batch_lanes.txt 2>&1
done
for cfg in -1 400 1296; do          # the bench line itself with the default layout, planned tiles, tiles of 896
  echo "== CFMM_BLOCKED_CFG=$cfg" >> gpurun_out/r2_bench_cfgs.txt
  CFMM_BLOCKED_CFG=$cfg timeout 300 python bench.py --steps 16000 --no-cpu 2>/dev/null | cut -c1-600 >> gpurun_out/r2_bench_cfgs.txt
done
tail -5 gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_bench_cfgs.txt; cat gpurun_out/r2_tiles.txt; grep -E '"ms"|LANES|B[0-9]+"' gpurun_out/r2_batch_lanes.txt | head -40
