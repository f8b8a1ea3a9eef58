PDL=1 python scripts/microbench.py 0 1 2>&1 | grep -v "store build" | tail -5
