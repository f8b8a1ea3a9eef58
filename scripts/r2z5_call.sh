#!/bin/bash
# Round 2, diagnostic (1 GPU, ~1 GPU-minute): why the loopback test stalled against the serial-poll variant library, and a
# second pass of the loopback tests against the product library
set -u
mkdir -p gpurun_out
CFMM_LIB=build/variants/libcfmm_serialpoll.so timeout 28 python -X faulthandler -m pytest tests/test_loopback_ranks.py -k "2-persist" -m gpu -q -x --timeout 14 -p no:cacheprovider 2>&1 | tail -70 > gpurun_out/r2z5_variant.txt
timeout 25 python -m pytest tests/test_loopback_ranks.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2z5_product.txt; cat gpurun_out/r2z5_product.txt
head -c 6000 gpurun_out/r2z5_variant.txt
