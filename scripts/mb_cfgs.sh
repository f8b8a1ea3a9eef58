for c in 0 1; do BLOCKED_CFG=$c python scripts/microbench.py 0 1 8 2>&1 | grep -v "hess\|store build" | tail -4; done
