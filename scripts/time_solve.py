"""Times the 1M-pool solve (BASELINE configs[4]) through both native loops, and the stages of the e2e call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

m, n = int(os.environ.get("M_POOLS", 1_000_000)), 4096
s = I.synth_const_product(m, n, seed=3)
hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]).pin_memory()
util = cf.Arbitrage(s["prices"])
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); st = cf.PoolStore(hp, validate=False); torch.cuda.synchronize(); tb = time.perf_counter() - t0
print(f"PoolStore build (upload + layout) {1e3*tb:.2f} ms")
for impl in ("persist", "hostloop"):
    for tol in (1e-6, 1e-9):
        ws = []
        for rep in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = cf.solve_pools(hp, util, tol=tol, store=st, native=impl, want_trades=False)
            torch.cuda.synchronize(); ws.append(time.perf_counter() - t0)
        print(f"{impl:9s} tol {tol:.0e}: {r.status} iters {r.iters} evals {r.evals} hvps {r.hvps} gap {r.gap:+.1e} infeas {r.primal_infeas:.1e} "
              f"solve wall min {1e3*min(ws):.3f} ms median {1e3*sorted(ws)[2]:.3f} ms (solver-only {1e3*r.wall_s:.3f} ms) value {r.value:.9g}", flush=True)
ws = []
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = cf.solve_pools(hp, util, tol=1e-6, want_trades=False)
    torch.cuda.synchronize(); ws.append(time.perf_counter() - t0)
print(f"e2e solve_pools from pinned host arrays: min {1e3*min(ws):.3f} ms median {1e3*sorted(ws)[2]:.3f} ms")
