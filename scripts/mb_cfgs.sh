for c in 1 0 2; do BLOCKED_CFG=$c python scripts/microbench.py 0 1 2>&1 | grep -v "hess\|store build" | tail -3; done
