#!/bin/bash
# tile-size / CTAs-per-SM sweep of the blocked kernels (experiment libraries built with -DCFMM_TILE_P / -DCFMM_CTAS_PER_SM)
set -u
mkdir -p gpurun_out
for v in libcfmm_b200.so libcfmm_b200_p512.so libcfmm_b200_p576.so libcfmm_b200_p640.so libcfmm_b200_p704.so libcfmm_b200_p768.so; do
  echo "== $v" >> gpurun_out/r2s_tiles.txt
  CFMM_LIB=$v timeout 200 python scripts/microbench.py 2>&1 | tail -4 >> gpurun_out/r2s_tiles.txt
  CFMM_LIB=$v timeout 200 python scripts/time_solve.py 2>&1 | sed -n 2,3p >> gpurun_out/r2s_tiles.txt
done
cat gpurun_out/r2s_tiles.txt
