import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I, pools as PL
m, n = 1_000_000, 4096
s = I.synth_const_product(m, n, seed=3)
hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]).pin_memory()
def T(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:44s} {1e3*best:7.2f} ms", flush=True); return r
T("hp.validate()", hp.validate)
sp = T("split_buckets", lambda: PL.split_buckets(hp))
T("3 x H2D (pinned)", lambda: (torch.from_numpy(hp.reserves).to("cuda", non_blocking=True), torch.from_numpy(hp.tok_idx).to("cuda", non_blocking=True), torch.from_numpy(hp.gamma).to("cuda", non_blocking=True)))
st = T("PoolStore(hp) [validate=True]", lambda: cf.PoolStore(hp))
st = T("PoolStore(hp, validate=False)", lambda: cf.PoolStore(hp, validate=False))
T("solve_pools(store=st) native", lambda: cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False, store=st))
T("solve_pools(hp) full e2e", lambda: cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
