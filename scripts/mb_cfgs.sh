python scripts/microbench.py 0 1 2>&1 | grep -v "hess" | tail -4
